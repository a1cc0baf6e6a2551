"""One-process-per-GPU driver of the factorization (torch.distributed / NCCL is the plumbing).

world == 1: a thin pass-through to the C ABI.

world > 1: subtrees of the assembly tree are dealt to ranks by proportional
mapping; each rank factorizes its subtrees in its own HBM and accumulates their inter-node
updates into its private copy of the upper tree; a sum all-reduce of the contiguous upper-tree
slice of the arena over NVLink hands every rank the assembled upper tree.  The upper tree is then
factorized block-column-cyclically, owner computes: a rank factorizes the block columns it owns
and computes every update whose DESTINATION block column it owns; each finished block column is
broadcast once by its owner (NCCL) before anybody uses it as a source.  No reductions in the
upper tree.  (SPLLT_B200_REPLICATED_TOP=1: every rank factorizes the whole upper tree instead.)

Solve: see DistSpLLT.solve_dev -- subtree sweeps on their owners, the upper tree redundantly, two
all-reduces of the work vector.
"""
import ctypes as C

import numpy as np

from .api import SpLLT, lib


class _DevArray:
    """Zero-copy view of a device buffer for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}


class DistSpLLT:
    def __init__(self, nb, rank=0, world=1, stream=None, **options):
        self.rank, self.world = rank, world
        self.local = SpLLT(nb=nb, ncpu=max(world, 1), **options)
        self.stream = stream
        self.top = None
        self.dist_top = False
        self.program = []

    # -------------------------------------------------------------- phases
    def analyse(self, n, ptr, row):
        s = self.local
        flag = s.analyse(n, ptr, row)
        L = s.L
        if self.world > 1:
            L.spllt_b200_partition(s.akeep, s.fkeep, self.rank, self.world)
        if self.stream is not None:
            s.set_stream(self.stream.cuda_stream)
        if self.world > 1:
            import torch
            b, e = C.c_longlong(0), C.c_longlong(0)
            L.spllt_b200_shared_region(s.akeep, C.byref(b), C.byref(e))
            self.top_range = (b.value, e.value)
            base = L.spllt_b200_arena_ptr(s.fkeep)
            if e.value > b.value:
                self.top = torch.as_tensor(_DevArray(base + 8 * b.value, e.value - b.value), device="cuda")
            self.dist_top = bool(L.spllt_b200_dist_top(s.akeep))
            if self.dist_top:
                self._build_program(base)
        return flag

    def _build_program(self, arena_base):
        """Phase-1 program: runs of kernel launches separated by block-column broadcasts."""
        import torch
        s, L = self.local, self.local.L
        nrec = L.spllt_b200_num_launch_records(s.akeep)
        rec = np.zeros((max(nrec, 1), 8), dtype=np.int64)
        L.spllt_b200_get_launch_records(s.akeep, rec.ctypes.data_as(C.POINTER(C.c_longlong)))
        rec = rec[:nrec]
        self.program = []
        run_start = None
        maxbuf = 0
        off, ld, rows, cols = C.c_longlong(0), C.c_int(0), C.c_int(0), C.c_int(0)
        for i in range(nrec):
            kind, phase = int(rec[i, 0]), int(rec[i, 4])
            if phase != 1:
                continue
            if kind == 3:
                if run_start is not None:
                    self.program.append(("run", run_start, i))
                    run_start = None
                node, c, owner = int(rec[i, 2]), int(rec[i, 3]), int(rec[i, 5])
                L.spllt_b200_bcol_region(s.akeep, node + 1, c, C.byref(off), C.byref(ld), C.byref(rows), C.byref(cols))
                self.program.append(("bcast", node + 1, c, owner, rows.value * cols.value))
                maxbuf = max(maxbuf, rows.value * cols.value)
            elif run_start is None:
                run_start = i
        if run_start is not None:
            self.program.append(("run", run_start, nrec))
        self.staging = torch.empty(max(maxbuf, 1), dtype=torch.float64, device="cuda")

    def factor_dev(self, d_val):
        """d_val: torch CUDA tensor holding the user's val.  Asynchronous on the stream."""
        s = self.local
        if self.world == 1:
            s.factor_dev(d_val.data_ptr())
            return
        import torch.distributed as dist
        L = s.L
        L.spllt_b200_factor_phase(s.akeep, s.fkeep, C.c_void_p(d_val.data_ptr()), 0)
        if self.top is not None:
            dist.all_reduce(self.top, op=dist.ReduceOp.SUM)     # the exchange step (NCCL over NVLink)
        if not self.dist_top:
            L.spllt_b200_factor_phase(s.akeep, s.fkeep, C.c_void_p(d_val.data_ptr()), 1)
            return
        for op in self.program:
            if op[0] == "run":
                L.spllt_b200_run_launches(s.fkeep, op[1], op[2])
            else:
                _, node, c, owner, count = op
                buf = self.staging[:count]
                if owner == self.rank:
                    L.spllt_b200_pack_bcol(s.akeep, s.fkeep, node, c, C.c_void_p(buf.data_ptr()))
                dist.broadcast(buf, src=owner)
                if owner != self.rank:
                    L.spllt_b200_unpack_bcol(s.akeep, s.fkeep, node, c, C.c_void_p(buf.data_ptr()))

    def factor_host(self, val):
        """Reference-facing call with a HOST val array (spllt_factor)."""
        s = self.local
        if self.world == 1:
            s.factor(val)
            return
        import torch
        if getattr(self, "_dval", None) is None or self._dval.numel() != val.size:
            self._dval = torch.empty(val.size, dtype=torch.float64, device="cuda")
        self._dval.copy_(torch.from_numpy(val), non_blocking=True)
        self.factor_dev(self._dval)

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
        return self.local.pivot_flag()

    def profile_factor(self, d_val):
        """per-kernel-kind milliseconds of one un-graphed factorization (this rank's launches)"""
        return self.local.profile_factor(d_val.data_ptr())

    def solve_path(self):
        return ("multi-GPU: persistent pipelined kernels on the rank's subtrees, upper tree redundantly, "
                "two NCCL all-reduces of the work vector")

    # -------------------------------------------------------------- reporting
    def work_multiplier(self):
        return 1   # one factorization is shared by all ranks (strong scaling)

    def scaling(self):
        return "strong"

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
        how = ("distributed block-column-cyclically (owner computes, %d block-column broadcasts)"
               % sum(1 for op in self.program if op[0] == "bcast")) if self.dist_top else "replicated"
        return ("subtree->GPU proportional mapping (%d subtrees), sum all-reduce of the %.2f GB upper-tree "
                "slice, upper tree (%.0f%% of flops) %s; rank %d subtree share %.1f%%"
                % (nsub, (self.top_range[1] - self.top_range[0]) * 8 / 1e9, 100 * top / tot, how, self.rank,
                   100 * mine / tot))
