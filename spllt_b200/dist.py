"""One-process-per-GPU driver of the factorization (torch.distributed / NCCL is the plumbing).

world == 1: a thin pass-through to the C ABI.

world > 1: the pruned subtrees of the assembly tree (the reference's own unit of
tree parallelism, spllt_prune_tree with nth = number of GPUs) are dealt to ranks by weight; each
rank factorizes its subtrees in its own HBM and accumulates their inter-node updates into its
private copy of the upper tree; ONE exchange step -- a sum all-reduce of the contiguous
upper-tree slice of the arena over NVLink -- hands every rank the assembled upper tree, which
is then factorized (round 1: replicated on every rank; 2-D block-cyclic is the next step).
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
        return flag

    def factor_dev(self, d_val):
        """d_val: torch CUDA tensor holding the user's val.  Asynchronous on the stream."""
        s = self.local
        if self.world == 1:
            s.factor_dev(d_val.data_ptr())
            return
        import torch.distributed as dist
        s.L.spllt_b200_factor_phase(s.akeep, s.fkeep, C.c_void_p(d_val.data_ptr()), 0)
        if self.top is not None:
            dist.all_reduce(self.top, op=dist.ReduceOp.SUM)     # the exchange step (NCCL over NVLink)
        s.L.spllt_b200_factor_phase(s.akeep, s.fkeep, C.c_void_p(d_val.data_ptr()), 1)

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

    def wait(self):
        self.local.wait()
        if self.world > 1:
            import torch
            torch.cuda.current_stream().synchronize()

    def pivot_flag(self):
        return self.local.pivot_flag()

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
        return ("subtree->GPU proportional mapping (%d subtrees), sum all-reduce of the %.2f GB upper-tree "
                "slice, upper tree (%.0f%% of flops) replicated; rank %d subtree share %.1f%%"
                % (nsub, (self.top_range[1] - self.top_range[0]) * 8 / 1e9, 100 * top / tot, self.rank,
                   100 * mine / tot))
