#!/usr/bin/env python
"""bench.py -- headline benchmark of the SpLLT numerical phase on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

A "step" is one numerical factorization (spllt_factor + spllt_wait) of BASELINE.json's headline
configuration, configs[2]: 3D Poisson 7-point 100^3 (n = 1 000 000), nb = 768, METIS nested
dissection, nemin = 32 -- the configuration the metric is quoted on at 1/2/4/8 GPUs; it fits one GPU
(7.8 GB of factor).  Protocol of the reference's own benchmark scripts
(aux/run_tests_pde_job.sh:77-90: analyse once, then time the factorization and the solve).

`value` = factor GFLOP/s with `val` already resident in HBM (flops = the reference's own count,
sum_nodes sum_j (m-n+j)^2, src/spllt_analyse_mod.F90:1013-1021); `e2e` = the same metric through
the reference-facing C ABI with HOST buffers (spllt_factor copies val H2D, spllt_wait, pivot flag
read back).  `roofline` = the DMMA tile kernels against the measured FP64 tensor peak
(ALGORITHMIC flops: 2 K per updated entry of the lower triangle; the issued/padded count is
reported beside it); `roofline_solve` = the triangular solves against the measured HBM bandwidth;
`extra` = BASELINE configs[1] (64^3 factor + solve) and configs[4] (80^3 solve, nrhs 1/16/64).
Prints ONE JSON line on rank 0.

--impl reference times the CPU restatement of the reference's OpenMP build (oracle/: OpenMP tasks +
sequential OpenBLAS; the Fortran reference cannot be compiled in this image) on all host cores,
same config / metric; a step is a bounded sample of the factorization (see run_reference).  That
arm never imports or loads the product (spllt_b200 / libspllt_b200.so).
"""
import argparse
import ctypes as C
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (generator, args, nb, description)
    "p2d200": ("poisson2d", (200,), 256, "2D Poisson 5-point 200x200 (n=40000), nb=256"),
    "p3d64": ("poisson3d", (64,), 512, "3D Poisson 7-point 64^3 (n=262144), nb=512"),
    "p3d80": ("poisson3d", (80,), 512, "3D Poisson 7-point 80^3 (n=512000), nb=512"),
    "p3d100": ("poisson3d", (100,), 768, "3D Poisson 7-point 100^3 (n=1000000), nb=768"),
    "el3d60": ("elasticity3d", (60,), 768, "3D elasticity 27-point 3-dof 60^3 (n=648000), nb=768"),
    "p3d32": ("poisson3d", (32,), 256, "3D Poisson 7-point 32^3 (n=32768), nb=256 [smoke size]"),
}
HEADLINE = "p3d100"
REF_BUDGET_S = 200.0     # the whole --impl reference run (all steps) aims at this much CPU wall time


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("SPLLT_BENCH_WORKLOAD", HEADLINE), choices=sorted(WORKLOADS))
    ap.add_argument("--nrhs", default="1", help="right-hand sides of the timed solve; a comma list times several "
                                                "(the first is the headline, the others go to `solve_more`)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the 64^3 / 80^3 side measurements")
    return ap.parse_args()


def matrices_module():
    """spllt_b200/matrices.py loaded by path (pure numpy): the CPU arm must not import the product package."""
    spec = importlib.util.spec_from_file_location("spllt_matrices", os.path.join(ROOT, "spllt_b200", "matrices.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def make_matrix(workload):
    M = matrices_module()
    gen, a, nb, desc = WORKLOADS[workload]
    return getattr(M, gen)(*a), nb, desc


L2_NOTE = "working set (factor) far above the 126 MB L2, no flush needed"


def workload_config(desc, n, nnz_lower, nnz_L, flops, nb):
    """`config` of the JSON line -- identical in both arms (same keys, same values)."""
    return {"workload": desc + ", METIS nested dissection, nemin=32", "n": int(n), "nnz_lower": int(nnz_lower),
            "nnz_L": int(nnz_L), "flops_per_step": int(flops), "nb": int(nb), "l2": L2_NOTE}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    return 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for nme, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------ CPU arm
class CpuReference:
    """The restated reference (oracle/) on the host cores.  Symbolic inputs come from the SSIDS
    stand-in compiled into oracle/libssids_standin.so -- nothing of the product is loaded."""

    def __init__(self, mat, nb, nthreads):
        from oracle import oracle as O
        n, ptr, row, val = mat
        self.val, self.nthreads = val, nthreads
        order, sptr, sparent, rptr, rlist = O.symbolic(n, ptr, row, nemin=32)
        # the reference prunes the tree for ncpu workers (src/spllt_analyse_mod.F90:297)
        self.o = O.Oracle(n, ptr, row, order, sptr, sparent, rptr, rlist, nb, ncpu=nthreads)
        w = self.o.weight()
        nn = len(sptr) - 1
        own = w[:-1].astype(np.float64).copy()
        par = np.asarray(sparent) - 1
        np.subtract.at(own, par[par < nn], w[:-1][par < nn].astype(np.float64))
        self.cum = np.cumsum(own)          # flops of nodes 1..k (postorder prefix)
        self.small = self.o.small()
        self.nn = nn
        self.total = float(self.cum[-1]) if nn else 0.0
        self.flops_int = int(w[-1])                      # akeep%weight(nnodes+1): the factorization's flop count
        ncol = np.diff(np.asarray(sptr)).astype(np.int64)
        mrow = np.diff(np.asarray(rptr)).astype(np.int64)
        self.nnz_L = int(np.sum(ncol * mrow - ncol * (ncol - 1) // 2))

    def prefix_for(self, frac):
        """last node of the smallest postorder prefix holding >= frac of the flops that does not
        cut a pruned subtree; (last_node, flops of the prefix)"""
        if frac >= 0.9 or self.nn == 0:
            return self.nn, self.total
        ok = np.nonzero((self.cum >= frac * self.total) & (self.small >= 0))[0]
        last = int(ok[0]) + 1 if len(ok) else self.nn
        return last, float(self.cum[last - 1])

    def run(self, last_node):
        t = time.perf_counter()
        self.o.factor_prefix(self.val, self.nthreads, last_node)
        return time.perf_counter() - t

    def sampled(self, nsteps_total, budget_s):
        """Chooses the sample: one calibration run on a 5 % prefix gives a rate; the sample is the
        largest prefix (whole subtrees + the upper-tree nodes above them, in postorder) such that
        nsteps_total steps fit the budget.  Returns (last_node, flops, description)."""
        last, fl = self.prefix_for(0.05)
        dt = self.run(last)            # also allocates the factor storage (not timed later)
        dt = min(dt, self.run(last))
        rate = fl / max(dt, 1e-9)
        frac = min(1.0, budget_s * rate / (max(nsteps_total, 1) * max(self.total, 1.0)))
        last, fl = self.prefix_for(frac)
        if last >= self.nn:
            return self.nn, self.total, "one full factorization of the workload per step"
        return last, fl, ("bounded sample per step: nodes 1..%d of %d in postorder (whole subtrees and the "
                          "upper-tree nodes above them) = %.1f%% of the factorization's flops, all of their "
                          "init / factorize / update tasks" % (last, self.nn, 100.0 * fl / self.total))


def run_reference(args, rank, world):
    if rank != 0:
        return
    mat, nb, desc = make_matrix(args.workload)
    cores = os.cpu_count() or 1
    ref = CpuReference(mat, nb, cores)
    last, fl, what = ref.sampled(args.warmup + args.steps, REF_BUDGET_S)
    for _ in range(args.warmup):
        ref.run(last)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.run(last)
    sec = (time.perf_counter() - t0) / max(args.steps, 1)
    val = fl / sec / 1e9
    sample = ("%s; OpenMP tasks on %d threads + sequential OpenBLAS (restated reference OMP build; the Fortran "
              "reference cannot be compiled in this image)" % (what, cores))
    out = {
        "impl": "reference", "metric": "factor_gflops", "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # same `config` as the GPU arm; the metric is a rate, the step of THIS arm is the bounded sample below
        "config": workload_config(desc, mat[0], mat[3].size, ref.nnz_L, ref.flops_int, nb),
        "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": sample,
                         "sample_flops_per_step": int(fl)},
        "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------ GPU arm
class Bench:
    def __init__(self, args, rank, world, local):
        import torch
        self.torch = torch
        self.args, self.rank, self.world, self.local = args, rank, world, local
        import spllt_b200 as sp
        from spllt_b200 import dist as spdist
        self.sp, self.spdist = sp, spdist
        self.L = sp.lib()
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)
        self.sptr_ = C.c_void_p(self.stream.cuda_stream)
        self.hbm, self.hbm_src = hbm_peak()

    def ev(self):
        return self.torch.cuda.Event(enable_timing=True)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def maxrank(self, ms):
        if self.world > 1:
            import torch.distributed as dist
            t = self.torch.tensor([ms], device="cuda", dtype=self.torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def peaks(self):
        out = {}
        for kind, nme in ((0, "dmma"), (1, "dfma")):
            self.L.spllt_b200_peak_probe(kind, 2000, self.sptr_)
            self.torch.cuda.synchronize()
            best = 0.0
            for _ in range(3):
                a, b = self.ev(), self.ev()
                a.record()
                fl = self.L.spllt_b200_peak_probe(kind, 20000, self.sptr_)
                b.record()
                self.torch.cuda.synchronize()
                best = max(best, fl / a.elapsed_time(b) / 1e9)
            out[nme] = best
        return out

    def setup(self, workload):
        mat, nb, desc = make_matrix(workload)
        n, ptr, row, val = mat
        solver = self.spdist.DistSpLLT(nb=nb, rank=self.rank, world=self.world, stream=self.stream)
        t0 = time.perf_counter()
        solver.analyse(n, ptr, row)
        t_analyse = time.perf_counter() - t0
        return mat, nb, desc, solver, t_analyse

    def time_factor(self, solver, d_val, steps, warmup):
        for _ in range(warmup):
            solver.factor_dev(d_val)
        self.barrier()
        a, b = self.ev(), self.ev()
        self.barrier()
        a.record()
        for _ in range(steps):
            solver.factor_dev(d_val)
        b.record()
        self.barrier()
        return self.maxrank(a.elapsed_time(b) / steps)

    def time_e2e(self, solver, hv, steps):
        for _ in range(2):
            solver.factor_host(hv)
            solver.wait()
        self.barrier()
        a, b = self.ev(), self.ev()
        a.record()
        for _ in range(steps):
            solver.factor_host(hv)          # spllt_factor: H2D copy of val + factorization
            solver.wait()                   # spllt_wait
            _ = solver.pivot_flag()         # D2H read of the step's result
        b.record()
        self.barrier()
        return self.maxrank(a.elapsed_time(b) / steps)

    def solve_bytes(self, s, n, nrhs):
        # per sweep 8 nnz(L) + 2*8*n*nrhs + 3*8*sum(m-n)*nrhs; fwd + bwd = 2x
        sptr, sparent, rptr, rlist = s.symbolic()
        upd = int(np.sum(np.diff(rptr) - np.diff(sptr)))
        return 2 * (8 * s.num_factor + 16 * n * nrhs + 24 * upd * nrhs)

    def time_solve(self, solver, mat, nrhs, reps):
        """seconds per solve (fwd + bwd, nrhs right-hand sides), backward errors, one solution"""
        torch = self.torch
        n, ptr, row, val = mat
        M = matrices_module()
        xs = np.asfortranarray(np.tile(np.arange(1.0, nrhs + 1), (n, 1)))
        if nrhs > 1:   # seeded random right-hand sides beside the reference's x = r convention
            rng = np.random.Generator(np.random.PCG64(20261018))
            xs[:, 1::2] = rng.standard_normal((n, xs[:, 1::2].shape[1]))
        rhs = np.asfortranarray(M.matvec(n, ptr, row, val, xs))
        d_rhs = [torch.tensor(rhs.T.copy(), device="cuda") for _ in range(reps + 2)]
        for d in d_rhs[:2]:
            solver.solve_dev(d, nrhs)
        self.barrier()
        a, b = self.ev(), self.ev()
        a.record()
        for d in d_rhs[2:]:
            solver.solve_dev(d, nrhs)
        b.record()
        self.barrier()
        ms = self.maxrank(a.elapsed_time(b) / reps)
        x = np.asfortranarray(d_rhs[2].cpu().numpy().T)
        ok, err = self.sp.chkerr(n, ptr, row, val, x, rhs)
        fwd_err = float(np.abs(x - xs).max() / np.abs(xs).max())
        return ms, ok, err, fwd_err, d_rhs[0]

    def solve_report(self, solver, mat, nrhs, reps):
        s = solver.local
        n = mat[0]
        ms, ok, err, fwd_err, d0 = self.time_solve(solver, mat, nrhs, reps)
        sbytes = self.solve_bytes(s, n, nrhs)
        if self.world > 1:
            path = solver.solve_path()
        elif nrhs <= self.L.spllt_b200_pipe_max_nrhs(s.akeep) and not os.environ.get("SPLLT_B200_SOLVE_LEVELSET"):
            path = "persistent pipelined kernels k_solve_pipe<fwd>/<bwd> (64-row strips, flags in HBM)"
        else:
            path = "level-set launches k_fwd_diag/k_fwd_upd/k_bwd_upd/k_bwd_diag"
        rep = {"nrhs": nrhs, "seconds": ms / 1e3, "seconds_per_rhs": ms / 1e3 / nrhs, "algorithmic_bytes": sbytes,
               "achieved_gbs": sbytes / ms / 1e6, "hbm_peak_gbs": self.hbm,
               "frac_of_hbm": sbytes / ms / 1e6 / self.hbm / max(self.world, 1), "path": path,
               "scaled_backward_error_max": float(err.max()), "rhs_ok": int(ok), "forward_error_max": fwd_err}
        return rep, d0


def main():
    args = parse()
    nrhs_list = [int(x) for x in str(args.nrhs).split(",")]
    args.nrhs = nrhs_list[0]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = Bench(args, rank, world, local)
    L = B.L
    peaks = B.peaks()
    warm = max(args.warmup, 3)

    mat, nb, desc, solver, t_analyse = B.setup(args.workload)
    n, ptr, row, val = mat
    s = solver.local
    flops = s.num_flops
    d_val = torch.tensor(val, device="cuda")
    h_val = torch.tensor(val).pin_memory()
    cfg_more = {"nnz_L": int(s.num_factor), "multi_gpu": solver.describe(),
                "factor_gb": L.spllt_b200_arena_doubles(s.akeep) * 8 / 1e9}
    launches_per_factor = int(solver.launches_per_factor())

    # ---------------- timed region: K factorizations, val resident in HBM
    clocks = ClockSampler(local)
    clocks.start()
    ms = B.time_factor(solver, d_val, args.steps, warm)
    clk = clocks.stop()
    pivot = solver.pivot_flag()
    value = flops / ms / 1e6

    # ---------------- e2e: reference-facing C ABI with host buffers
    ms_e2e = B.time_e2e(solver, h_val.numpy(), args.steps)
    e2e = {"value": flops / ms_e2e / 1e6, "unit": "GFLOP/s", "ms_per_step": ms_e2e,
           "h2d_bytes_per_step": int(val.nbytes), "d2h_bytes_per_step": 4}

    # ---------------- solve: seconds per RHS, achieved HBM bandwidth
    reps = max(min(args.steps, 10), 3)
    solve, d0 = B.solve_report(solver, mat, args.nrhs, reps)
    solve["launches"] = int(L.spllt_b200_solve_launches(s.fkeep, 0)) if world == 1 else None
    roofline_solve = {"bound": "hbm", "kernel": "k_solve_pipe<1,fwd> + k_solve_pipe<1,bwd> (one launch per sweep)"
                      if "pipelined" in solve["path"] else solve["path"],
                      "achieved": solve["achieved_gbs"] / max(world, 1), "peak": B.hbm, "unit": "GB/s",
                      "frac": solve["frac_of_hbm"], "peak_source": B.hbm_src,
                      "bytes_per_solve": solve["algorithmic_bytes"], "traffic": None}
    if world == 1:
        solve["profile_ms"] = s.profile_solve(d0.data_ptr(), args.nrhs)
        tpath = os.path.join(ROOT, "profiles", "solve_traffic.json")
        if os.path.exists(tpath):   # dram bytes of k_solve_pipe<fwd> + <bwd> from one ncu --set full capture
            tr = json.load(open(tpath))
            solve["traffic"] = tr
            if tr.get("workload") == args.workload:
                roofline_solve["traffic"] = tr.get("dram_bytes_read", 0) + tr.get("dram_bytes_write", 0)
    solve_more = []
    for nr in nrhs_list[1:]:
        rep, _ = B.solve_report(solver, mat, nr, 3)
        solve_more.append(rep)
    parity = {"scaled_backward_error_max": solve["scaled_backward_error_max"], "rhs_ok": solve["rhs_ok"],
              "nrhs": args.nrhs, "tol": 1e-14, "forward_error_max": solve["forward_error_max"],
              "pivot_flag": int(pivot)}
    if world > 1:
        # the distributed factor against a single-GPU factorization of the same matrix on every rank
        # (entries of the nodes the rank holds), outside the timed region
        parity["factor_vs_single_gpu"] = solver.compare_with_single_gpu(d_val)

    # ---------------- roofline of the dominant kernel (DMMA tile updates)
    roofline = None
    prof = solver.profile_factor(d_val)      # collective: every rank runs its un-graphed, event-timed pass
    if rank == 0:
        bd = np.zeros(4, dtype=np.int64)
        L.spllt_b200_launch_breakdown(s.akeep, bd.ctypes.data_as(C.POINTER(C.c_longlong)))
        issued = float(L.spllt_b200_tile_flops(s.akeep))            # padded tiles, masked halves included
        algo = float(L.spllt_b200_tile_flops_algo(s.akeep))         # 2 K per updated entry (i >= j)
        tile_ms = prof["tile_s"] + prof["tile_l"]
        n_tile_launch = int(bd[1] + bd[2])
        achieved = algo / tile_ms / 1e9 if tile_ms > 0 else 0.0      # TFLOP/s
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):   # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture
            traffic = json.load(open(tpath))
        roofline = {"bound": "tensor", "kernel": "k_tile_tma<64,2> (persistent TMA + DMMA.8x8x4, 128x64 tiles) + "
                                                 "k_tile<64,64,32,32,2> (small launches)",
                    "achieved": achieved, "peak": peaks["dmma"], "unit": "TFLOP/s",
                    "frac": achieved / peaks["dmma"] if peaks["dmma"] else None,
                    "flops": "algorithmic: 2*K per destination entry with i >= j of every tile update "
                             "(what the reference's dgemm/dsyrk calls count); this rank's tiles",
                    "achieved_issued": issued / tile_ms / 1e9 if tile_ms > 0 else 0.0,
                    "issued_over_algorithmic": issued / algo if algo else None,
                    # bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full); details beside it
                    "traffic": (traffic or {}).get("dram_bytes_per_launch"), "traffic_detail": traffic,
                    "peak_source": "own register-resident mma.sync.m8n8k4.f64 probe on 148 SMs, measured in this "
                                   "run (MEASURED_PEAKS.json has no FP64 figure); DFMA probe = %.1f TFLOP/s" % peaks["dfma"],
                    "flops_per_launch": algo / max(n_tile_launch, 1), "launches": n_tile_launch,
                    "avg_launch_ms": tile_ms / max(n_tile_launch, 1),
                    "share_of_step": tile_ms / max(sum(prof.values()), 1e-9),
                    "profile_ms": prof, "whole_factor_frac_of_peak": value / 1e3 / peaks["dmma"] / max(world, 1)}

    # ---------------- side measurements: BASELINE configs[1] and configs[4] (N = 1 only)
    extra = None
    if world == 1 and not args.no_extra and args.workload == HEADLINE:
        extra = {}
        del solver, s
        torch.cuda.empty_cache()
        for wl, nrhs_list in (("p3d64", (1,)), ("p3d80", (1, 16, 64))):
            m2, nb2, desc2, sol2, _ = B.setup(wl)
            dv2 = torch.tensor(m2[3], device="cuda")
            ms2 = B.time_factor(sol2, dv2, max(args.steps, 3), 3)
            ent = {"workload": desc2, "factor_ms": ms2, "factor_gflops": sol2.local.num_flops / ms2 / 1e6,
                   "factor_frac_of_dmma_peak": sol2.local.num_flops / ms2 / 1e9 / peaks["dmma"], "solve": []}
            for nr in nrhs_list:
                rep, _ = B.solve_report(sol2, m2, nr, 5)
                ent["solve"].append(rep)
            extra[wl] = ent
            del sol2, dv2
            torch.cuda.empty_cache()
        solver = None

    # ---------------- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        ref = CpuReference(mat, nb, cores)
        last, fl, what = ref.sampled(2, 20.0)     # ~10-20 s of CPU work in two steps
        sec = min(ref.run(last), ref.run(last))
        cpu = {"value": fl / sec / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": "port", "seconds": sec,
               "sample": "%s (best of 2); OpenMP tasks on %d threads + sequential OpenBLAS 0.3.31 (restated "
                         "reference OMP build)" % (what, cores)}

    if rank == 0:
        out = {
            "metric": "factor_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(desc, n, val.size, cfg_more["nnz_L"], flops, nb),
            "multi_gpu": cfg_more["multi_gpu"], "factor_gb": cfg_more["factor_gb"],
            "factor_seconds": ms / 1e3, "analyse_seconds_host": t_analyse,
            "clocks": clk, "e2e": e2e, "gpu_launches": None,
            "roofline": roofline, "roofline_solve": roofline_solve, "cpu_baseline": cpu, "solve": solve,
            "solve_more": solve_more or None, "parity": parity, "extra": extra,
        }
        out["gpu_launches"] = launches_per_factor * args.steps
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
