"""One-process-per-GPU driver of the distributed factorization / solve.

world == 1: a thin pass-through to the C ABI.

world > 1: everything numerical happens inside libspllt_b200.so -- this module only
bootstraps the ranks (torch.distributed is the plumbing: it all-gathers the 128-byte CUDA IPC
records of spllt_b200_comm_export so that every rank can map every peer's factor arena) and then
calls the same entry points as on one GPU.

Factorization (spllt_b200/csrc/analyse.cpp partition_tree / build_factor_schedule, engine.cu):
  * subtrees of the assembly tree are dealt to the ranks by proportional mapping; each rank
    factorizes its subtrees in its own HBM and accumulates their contributions into the upper tree
    in one generated element per subtree (local HBM; the reference's subtree `buffer`,
    src/spllt_kernels_mod.F90:780-821); when a subtree is complete one kernel scatters the element
    straight into the arenas of the ranks that own the destination block columns (RED.ADD.F64 on
    peer-mapped addresses over NVLink: spllt_subtree_apply_buffer,
    src/spllt_factorization_mod.F90:39-191) -- nothing is all-reduced;
  * the upper tree is distributed by block column and walked in the same step order on every rank:
    the owner of a step factorizes the block column (panel chain), copies it into every peer's
    arena and raises its flag there; every rank applies it to the destination block columns it owns,
    the ones whose own step comes next first (static look-ahead), so the next owner's panel chain
    overlaps the other ranks' updates.  One in-order stream per rank, captured as one CUDA graph.

Solve: see DistSpLLT.solve_dev -- subtree sweeps on their owners, the upper tree (whose factor every
rank holds) redundantly, two all-reduces of the work vector.
"""
import ctypes as C

import numpy as np

from .api import SpLLT, lib

HANDLE_BYTES = 128


class _DevArray:
    """Zero-copy view of a device buffer for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}


class DistSpLLT:
    def __init__(self, nb, rank=0, world=1, stream=None, **options):
        self.rank, self.world = rank, world
        self.local = SpLLT(nb=nb, ncpu=max(world, 1), **options)
        self.stream = stream
        self._single = None

    # -------------------------------------------------------------- phases
    def analyse(self, n, ptr, row):
        s = self.local
        flag = s.analyse(n, ptr, row)
        L = s.L
        if self.world > 1:
            L.spllt_b200_partition(s.akeep, s.fkeep, self.rank, self.world)
            if self.stream is None:
                # the collectives of the solve are ordered against torch's current stream only
                import torch
                self.stream = torch.cuda.current_stream()
        if self.stream is not None:
            s.set_stream(self.stream.cuda_stream)
        if self.world > 1:
            self._attach()
        self._mat = (n, ptr, row)
        return flag

    def _attach(self):
        """all-gather of the IPC records, then every rank maps every peer's arena and flag block"""
        import torch
        import torch.distributed as dist
        s, L = self.local, self.local.L
        mine = np.zeros(HANDLE_BYTES, dtype=np.uint8)
        rc = L.spllt_b200_comm_export(s.fkeep, mine.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise RuntimeError("spllt_b200_comm_export failed (%d)" % rc)
        t = torch.from_numpy(mine).cuda()
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t)
        table = np.ascontiguousarray(torch.stack(out).cpu().numpy())
        rc = L.spllt_b200_comm_attach(s.fkeep, self.rank, self.world, table.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise RuntimeError("spllt_b200_comm_attach failed (%d)" % rc)
        dist.barrier()    # nobody starts pushing before everybody has mapped everybody

    def factor_dev(self, d_val):
        """d_val: torch CUDA tensor holding the user's val.  Asynchronous on the stream; with several
        ranks every rank calls it (the ranks meet inside the captured graph, not here)."""
        self.local.factor_dev(d_val.data_ptr())

    def factor_host(self, val):
        """Reference-facing call with a HOST val array (spllt_factor: H2D copy + factorization)."""
        self.local.factor(val)

    def solve_dev(self, d_x, nrhs=1):
        """d_x: torch CUDA tensor, nrhs x n (row r = right-hand side r, i.e. column-major n x nrhs);
        overwritten with the solution on EVERY rank.  Asynchronous on the stream.

        world > 1: every rank sweeps the subtrees it owns in its own HBM, the upper
        tree is swept redundantly by every rank (its factor is replicated), and the pivot-order work
        vector (8 n nrhs bytes) is summed over the ranks twice: after the forward sweep of the
        subtrees (contributions to the upper tree) and after the backward sweep (solution pieces)."""
        s = self.local
        if self.world == 1:
            s.solve_dev(d_x.data_ptr(), nrhs)
            return
        import torch
        import torch.distributed as dist
        L = s.L
        assert torch.cuda.current_stream().cuda_stream == self.stream.cuda_stream, \
            "the NCCL all-reduces are ordered against torch's current stream: make the engine's stream current"
        xw = torch.as_tensor(_DevArray(L.spllt_b200_xw_ptr(s.fkeep, nrhs), s.n * nrhs), device="cuda")
        px = C.c_void_p(d_x.data_ptr())
        L.spllt_b200_solve_phase(s.fkeep, nrhs, px, s.n, 0)
        L.spllt_b200_solve_phase(s.fkeep, nrhs, px, s.n, 1)
        dist.all_reduce(xw, op=dist.ReduceOp.SUM)
        L.spllt_b200_solve_phase(s.fkeep, nrhs, px, s.n, 2)
        L.spllt_b200_solve_phase(s.fkeep, nrhs, px, s.n, 3)
        L.spllt_b200_solve_phase(s.fkeep, nrhs, px, s.n, 4)
        dist.all_reduce(xw, op=dist.ReduceOp.SUM)
        L.spllt_b200_solve_phase(s.fkeep, nrhs, px, s.n, 5)

    def wait(self):
        self.local.wait()
        if self.world > 1:
            import torch
            torch.cuda.current_stream().synchronize()

    def pivot_flag(self):
        """0 = ok, else the smallest 1-based pivot column that failed on ANY rank."""
        f = self.local.pivot_flag()
        if self.world == 1:
            return f
        import torch
        import torch.distributed as dist
        big = 2 ** 31 - 1
        t = torch.tensor([f if f > 0 else big], device="cuda", dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        v = int(t.item())
        return 0 if v == big else v

    def profile_factor(self, d_val):
        """per-kernel-kind milliseconds of one un-graphed factorization (this rank's launches).
        Collective: with several ranks every rank must call it."""
        import os
        csv = os.environ.get("SPLLT_BENCH_PROFILE_CSV")     # diagnostics: one line per launch, per rank
        return self.local.profile_factor(d_val.data_ptr(), "%s.rank%d.csv" % (csv, self.rank) if csv else None)

    def compare_with_single_gpu(self, d_val):
        """max relative difference between this rank's part of the distributed factor (its subtrees +
        the upper tree) and a plain single-GPU factorization of the same matrix run on this rank's GPU;
        the maximum over the ranks is returned on every rank.  Collective."""
        import torch
        import torch.distributed as dist
        s = self.local
        if self._single is None:
            ref = SpLLT(nb=s.options.nb, ncpu=max(self.world, 1))
            ref.analyse(*self._mat)
            self._single = ref
        ref = self._single
        ref.factor_dev(d_val.data_ptr())
        ref.wait()
        self.wait()
        out = np.zeros(2)
        rc = s.L.spllt_b200_compare_factor(s.akeep, s.fkeep, ref.akeep, ref.fkeep, out.ctypes.data_as(C.POINTER(C.c_double)))
        if rc != 0 or out[1] <= 0:
            raise RuntimeError("spllt_b200_compare_factor failed")
        rel = float(out[0] / out[1])
        if self.world > 1:
            t = torch.tensor([rel], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            rel = float(t.item())
        return {"max_rel_diff": rel, "tol": 1e-12, "what": "every rank: entries of the nodes it holds (own subtrees + "
                "upper tree) vs a single-GPU factorization on the same GPU; max over ranks"}

    # -------------------------------------------------------------- reporting
    def scaling(self):
        return "strong"

    def solve_path(self):
        return ("multi-GPU: persistent pipelined kernels on the rank's subtrees, upper tree redundantly, "
                "two NCCL all-reduces of the work vector")

    def launches_per_factor(self):
        return int(self.local.L.spllt_b200_factor_launches(self.local.fkeep))

    def describe(self):
        if self.world == 1:
            return "single GPU"
        s = self.local
        own = np.array([s.L.spllt_b200_node_owner(s.akeep, k + 1) for k in range(s.nnodes)])
        w = s.weight()[:-1]
        # weight() is subtree-accumulated; per-node flops = weight - sum(children)
        nodes = s.nodes()
        per = w.copy()
        par = nodes[:, 2] - 1
        for k in range(s.nnodes):
            if par[k] < s.nnodes:
                per[par[k]] -= w[k]
        tot = float(per.sum())
        top = float(per[own < 0].sum())
        mine = float(per[own == self.rank].sum())
        nsub = int(np.sum((own >= 0) & ((par >= s.nnodes) | (own[np.minimum(par, s.nnodes - 1)] < 0))))
        nsteps = int(s.L.spllt_b200_num_top_steps(s.akeep))
        return ("subtree->GPU proportional mapping (%d subtrees; rank %d subtree share %.1f%% of flops); upper tree "
                "(%.0f%% of flops) distributed by block column, owner computes, %d steps with static look-ahead; "
                "generated elements of the subtrees scattered into the owners' HBM by peer RED.ADD.F64, finished "
                "block columns pushed into the peers' arenas (CUDA IPC mappings over NVLink), no all-reduce"
                % (nsub, self.rank, 100 * mine / tot, 100 * top / tot, nsteps))
