#!/usr/bin/env python
"""bench.py -- headline benchmark of the SpLLT numerical phase on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one numerical factorization (spllt_factor + spllt_wait) of BASELINE.json's
configs[1]: 3D Poisson 7-point 64^3 (n = 262144), nb = 512, METIS nested dissection, nemin = 32.
`value` = factor GFLOP/s with `val` already resident in HBM (flops = the reference's own count,
sum_nodes sum_j (m-n+j)^2, src/spllt_analyse_mod.F90:1013-1021); `e2e` = the same metric through
the reference-facing C ABI with HOST buffers (spllt_factor copies val H2D, spllt_wait, pivot flag
read back).  The solve (seconds per RHS, achieved HBM GB/s) is timed separately and reported in
`solve`.  Prints ONE JSON line on rank 0.

--impl reference times the CPU restatement of the reference's OpenMP build (oracle/, OpenMP tasks
+ sequential OpenBLAS) on the host cores, same config and metric.
"""
import argparse
import ctypes as C
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("SPLLT_BENCH_WORKLOAD", "p3d64"), choices=sorted(WORKLOADS))
    ap.add_argument("--nrhs", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def make_matrix(workload):
    from spllt_b200 import matrices as M
    gen, a, nb, desc = WORKLOADS[workload]
    return getattr(M, gen)(*a), nb, desc


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


def cpu_reference_factor(mat, nb, nthreads, runs):
    """The CPU arm: oracle port of the reference's OMP build (OpenMP tasks + sequential OpenBLAS)."""
    import spllt_b200 as sp
    from oracle.oracle import Oracle
    n, ptr, row, val = mat
    s = sp.SpLLT(nb=nb, ncpu=nthreads)     # the reference prunes the tree for ncpu workers
    s.analyse(n, ptr, row)
    sptr, sparent, rptr, rlist = s.symbolic()
    o = Oracle(n, ptr, row, s.order, sptr, sparent, rptr, rlist, nb, ncpu=nthreads)
    flops = s.num_flops
    times = []
    for _ in range(runs):
        t = time.perf_counter()
        o.factor(val, nthreads)
        times.append(time.perf_counter() - t)
    return flops, times, s, o


def run_reference(args, rank, world):
    if rank != 0:
        return
    mat, nb, desc = make_matrix(args.workload)
    cores = os.cpu_count() or 1
    flops, times, s, o = cpu_reference_factor(mat, nb, cores, args.warmup + args.steps)
    t = times[args.warmup:]
    sec = float(np.mean(t))
    val = flops / sec / 1e9
    out = {
        "impl": "reference", "metric": "factor_gflops", "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc + ", METIS nested dissection, nemin=32", "flops_per_step": flops},
        "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": cores, "kind": "port",
                         "sample": "full factorization of the workload per step, OpenMP tasks on %d threads + "
                                   "sequential OpenBLAS (restated reference OMP build; the Fortran reference "
                                   "cannot be compiled in this image)" % cores},
        "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import spllt_b200 as sp
    from spllt_b200 import dist as spdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = sp.lib()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sptr_ = C.c_void_p(stream.cuda_stream)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    mat, nb, desc = make_matrix(args.workload)
    n, ptr, row, val = mat
    solver = spdist.DistSpLLT(nb=nb, rank=rank, world=world, stream=stream)
    t0 = time.perf_counter()
    solver.analyse(n, ptr, row)
    t_analyse = time.perf_counter() - t0
    s = solver.local
    flops = s.num_flops
    d_val = torch.tensor(val, device="cuda")
    h_val = torch.tensor(val).pin_memory()

    # ---------------- FP64 tensor-pipe peak (no FP64 figure in MEASURED_PEAKS.json)
    peaks = {}
    for kind, nme in ((0, "dmma"), (1, "dfma")):
        L.spllt_b200_peak_probe(kind, 2000, sptr_)
        torch.cuda.synchronize()
        best = 0.0
        for _ in range(3):
            a, b = ev(), ev()
            a.record()
            fl = L.spllt_b200_peak_probe(kind, 20000, sptr_)
            b.record()
            torch.cuda.synchronize()
            best = max(best, fl / a.elapsed_time(b) / 1e9)
        peaks[nme] = best

    # ---------------- timed region: K factorizations, val resident in HBM
    for _ in range(max(args.warmup, 3)):
        solver.factor_dev(d_val)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    a, b = ev(), ev()
    barrier()
    a.record()
    for _ in range(args.steps):
        solver.factor_dev(d_val)
    b.record()
    barrier()
    ms = a.elapsed_time(b) / args.steps
    clk = clocks.stop()
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    pivot = solver.pivot_flag()
    work_units = solver.work_multiplier()      # 1 for the distributed factorization, N for replicas
    value = work_units * flops / ms / 1e6

    # ---------------- e2e: reference-facing C ABI with host buffers
    hv = h_val.numpy()
    for _ in range(2):
        solver.factor_host(hv)
        solver.wait()
    barrier()
    a2, b2 = ev(), ev()
    a2.record()
    for _ in range(args.steps):
        solver.factor_host(hv)          # spllt_factor: H2D copy of val + factorization
        solver.wait()                   # spllt_wait
        _ = solver.pivot_flag()         # D2H read of the step's result
    b2.record()
    barrier()
    ms_e2e = a2.elapsed_time(b2) / args.steps
    if world > 1:
        t = torch.tensor([ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e = {"value": work_units * flops / ms_e2e / 1e6, "unit": "GFLOP/s", "ms_per_step": ms_e2e,
           "h2d_bytes_per_step": int(val.nbytes), "d2h_bytes_per_step": 4}

    # ---------------- solve: seconds per RHS, achieved HBM bandwidth (single-GPU path)
    solve = None
    parity = None
    if True:
        nrhs = args.nrhs
        from spllt_b200 import matrices as M
        xs = np.asfortranarray(np.tile(np.arange(1.0, nrhs + 1), (n, 1)))
        rhs = np.asfortranarray(M.matvec(n, ptr, row, val, xs))
        reps = max(args.steps, 3)
        d_rhs = [torch.tensor(rhs.T.copy(), device="cuda") for _ in range(reps + 2)]
        for d in d_rhs[:2]:
            solver.solve_dev(d, nrhs)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a3, b3 = ev(), ev()
        a3.record()
        for d in d_rhs[2:]:
            solver.solve_dev(d, nrhs)
        b3.record()
        torch.cuda.synchronize()
        ms_solve = a3.elapsed_time(b3) / reps
        if world > 1:
            t = torch.tensor([ms_solve], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_solve = float(t.item())
        x = np.asfortranarray(d_rhs[2].cpu().numpy().T)
        ok, err = sp.chkerr(n, ptr, row, val, x, rhs)
        nfac = s.num_factor
        sptr, sparent, rptr, rlist = s.symbolic()
        upd = int(np.sum(np.diff(rptr) - np.diff(sptr)))
        # per sweep 8 nnz(L) + 2*8*n*nrhs + 3*8*sum(m-n)*nrhs; fwd + bwd = 2x
        sbytes = 2 * (8 * nfac + 16 * n * nrhs + 24 * upd * nrhs)
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] \
            if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        if world > 1:
            path = ("multi-GPU: persistent pipelined kernels on the rank's subtrees, upper tree redundantly, "
                    "two NCCL all-reduces of the work vector")
        elif nrhs <= L.spllt_b200_pipe_max_nrhs(s.akeep) and not os.environ.get("SPLLT_B200_SOLVE_LEVELSET"):
            path = "persistent pipelined kernels k_solve_pipe<fwd>/<bwd> (64-row strips, flags in HBM)"
        else:
            path = "level-set launches k_fwd_diag/k_fwd_upd/k_bwd_upd/k_bwd_diag"
        solve = {"nrhs": nrhs, "seconds": ms_solve / 1e3, "seconds_per_rhs": ms_solve / 1e3 / nrhs,
                 "algorithmic_bytes": sbytes, "achieved_gbs": sbytes / ms_solve / 1e6, "hbm_peak_gbs": hbm_peak,
                 "frac_of_hbm": sbytes / ms_solve / 1e6 / hbm_peak / max(world, 1),
                 "launches": int(L.spllt_b200_solve_launches(s.fkeep, 0)) if world == 1 else 8, "path": path}
        if world == 1:
            solve["profile_ms"] = s.profile_solve(d_rhs[0].data_ptr(), nrhs)
            tpath = os.path.join(ROOT, "profiles", "solve_traffic.json")
            if os.path.exists(tpath):   # dram bytes of k_solve_pipe<fwd> + <bwd> from one ncu --set full capture
                solve["traffic"] = json.load(open(tpath))
        parity = {"scaled_backward_error_max": float(err.max()), "rhs_ok": int(ok), "nrhs": nrhs, "tol": 1e-14,
                  "forward_error_max": float(np.abs(x - xs).max() / np.abs(xs).max()), "pivot_flag": int(pivot)}

    # ---------------- roofline of the dominant kernel (128x128 DMMA tile update)
    roofline = None
    if rank == 0:
        prof = s.profile_factor(d_val.data_ptr())
        bd = np.zeros(4, dtype=np.int64)
        L.spllt_b200_launch_breakdown(s.akeep, bd.ctypes.data_as(C.POINTER(C.c_longlong)))
        tile_flops = float(L.spllt_b200_tile_flops(s.akeep))          # flops issued by tile kernels
        tile_ms = prof["tile_s"] + prof["tile_l"]
        n_tile_launch = int(bd[2] + bd[3])
        achieved = tile_flops / tile_ms / 1e9 if tile_ms > 0 else 0.0  # TFLOP/s
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):   # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture
            traffic = json.load(open(tpath))
        roofline = {"bound": "tensor", "kernel": "k_tile_tma<64,2> (persistent TMA + DMMA.8x8x4, 128x64 tiles) + "
                                                 "k_tile<64,64,32,32,2> (small launches)",
                    "achieved": achieved, "peak": peaks["dmma"], "unit": "TFLOP/s",
                    "frac": achieved / peaks["dmma"] if peaks["dmma"] else None,
                    # bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full); details beside it
                    "traffic": (traffic or {}).get("dram_bytes_per_launch"), "traffic_detail": traffic,
                    "peak_source": "own register-resident mma.sync.m8n8k4.f64 probe on 148 SMs, measured in this "
                                   "run (MEASURED_PEAKS.json has no FP64 figure); DFMA probe = %.1f TFLOP/s" % peaks["dfma"],
                    "flops_per_launch": tile_flops / max(n_tile_launch, 1), "launches": n_tile_launch,
                    "avg_launch_ms": tile_ms / max(n_tile_launch, 1),
                    "share_of_step": tile_ms / sum(prof.values()),
                    "profile_ms": prof, "whole_factor_frac_of_peak": value / 1e3 / peaks["dmma"] / max(world, 1)}

    # ---------------- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        fl, times, _, orc = cpu_reference_factor(mat, nb, cores, 2)
        sec = min(times[1:]) if len(times) > 1 else times[0]
        cpu = {"value": fl / sec / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": "port", "seconds": sec,
               "sample": "one full factorization of the same workload (after one warm-up run), OpenMP tasks on "
                         "%d threads + sequential OpenBLAS 0.3.31 (restated reference OMP build)" % cores}
        # the reference's solve (sequential tree sweeps, src/spllt_solve_mod.F90:244-411) on the same factor
        try:
            from spllt_b200 import matrices as M2
            nr = args.nrhs
            xs2 = np.asfortranarray(np.tile(np.arange(1.0, nr + 1), (n, 1)))
            rhs2 = np.asfortranarray(M2.matvec(n, ptr, row, val, xs2))
            orc.prepare_solve(nr)
            ts = []
            for _ in range(3):
                xx = rhs2.copy(order="F")
                t0 = time.perf_counter()
                orc.solve(xx, 0)
                ts.append(time.perf_counter() - t0)
            cpu["solve_seconds_per_rhs"] = min(ts) / nr
            cpu["solve_sample"] = "forward + backward solve of the same system, nrhs=%d, restated reference solve on 1 thread" % nr
            if solve is not None:
                solve["cpu_seconds_per_rhs"] = cpu["solve_seconds_per_rhs"]
        except Exception as e:   # the baseline is a report, never a reason to lose the bench line
            cpu["solve_seconds_per_rhs"] = None
            cpu["solve_sample"] = "unavailable: %r" % (e,)

    if rank == 0:
        out = {
            "metric": "factor_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": solver.scaling(), "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc + ", METIS nested dissection, nemin=32", "n": n, "nnz_lower": int(val.size),
                       "nnz_L": int(s.num_factor), "flops_per_step": int(flops), "nb": nb,
                       "multi_gpu": solver.describe(),
                       "l2": "working set %.2f GB > 126 MB L2, no flush needed" %
                             (L.spllt_b200_arena_doubles(s.akeep) * 8 / 1e9)},
            "factor_seconds": ms / 1e3, "analyse_seconds_host": t_analyse,
            "clocks": clk, "e2e": e2e, "gpu_launches": int(solver.launches_per_factor() * args.steps),
            "roofline": roofline, "cpu_baseline": cpu, "solve": solve, "parity": parity,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
