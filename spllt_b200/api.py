"""ctypes mirror of the C ABI (include/spllt_iface.h + include/spllt_b200.h).

Host-side mirror of the reference's operator interface for the numerical phase:
`SpLLT.analyse / factor / wait / prepare_solve / solve` take the same arguments, with the same
meaning and error behaviour, as spllt_analyse / spllt_factor / spllt_wait /
spllt_prepare_solve / spllt_solve of the reference (src/spllt_analyse_mod.F90:23,
src/spllt_mod.F90:141,172, src/spllt_solve_mod.F90:32-224, interfaces/C/spllt_data_ciface.F90).
There is no Python or CPU compute path here: if libspllt_b200.so is missing the import fails,
and any numeric call without a CUDA device aborts inside the library.
"""
import ctypes as C
import os

import numpy as np

ORDER_NATURAL, ORDER_METIS, ORDER_USER = 0, 1, 2

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libspllt_b200.so")


class Options(C.Structure):
    """spllt_options_t, include/spllt_iface.h:14-31 (defaults :33-47)."""
    _fields_ = [(k, C.c_int) for k in (
        "print_level", "nrhs", "ncpu", "nb", "nemin", "prune_tree", "min_width_blas", "nb_min", "nb_max",
        "nrhs_min", "nrhs_max", "nb_linear_comp", "nrhs_linear_comp", "chunk")]

    def __init__(self, **kw):
        super().__init__(print_level=0, nrhs=1, ncpu=1, nb=16, nemin=32, prune_tree=1, min_width_blas=8,
                         nb_min=32, nb_max=32, nrhs_min=1, nrhs_max=1, nb_linear_comp=0,
                         nrhs_linear_comp=0, chunk=10)
        for k, v in kw.items():
            setattr(self, k, v)


class Inform(C.Structure):
    """spllt_inform_t, include/spllt_iface.h:49-57."""
    _fields_ = [(k, C.c_int) for k in ("flag", "maxdepth", "num_factor", "num_flops", "num_nodes", "stat")]


_lib = None


def load_library(path=None):
    """Loads libspllt_b200.so; raises if it has not been built (python spllt_b200/build.py)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or _LIBPATH
    if not os.path.exists(p):
        raise ImportError("%s not found: build it with `python spllt_b200/build.py` "
                          "(there is no fallback implementation)" % p)
    L = C.CDLL(p, mode=C.RTLD_GLOBAL)
    vp, vpp, ip, dp, llp = C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_double), \
        C.POINTER(C.c_longlong)
    OP, IP = C.POINTER(Options), C.POINTER(Inform)
    sig = {
        # reference ABI
        "spllt_analyse": (None, [vpp, vpp, OP, C.c_int, ip, ip, IP, ip]),
        "spllt_factor": (None, [vp, vp, OP, C.c_int, dp, IP]),
        "spllt_prepare_solve": (None, [vp, vp, C.c_int, C.c_int, C.POINTER(C.c_long), IP]),
        "spllt_set_mem_solve": (None, [vp, vp, C.c_int, C.c_int, C.c_long, dp, dp, IP]),
        "spllt_solve_workspace_size": (None, [vp, C.c_int, C.c_int, C.POINTER(C.c_long)]),
        "spllt_solve": (None, [vp, OP, ip, C.c_int, dp, IP, C.c_int]),
        "spllt_solve_worker": (None, [vp, OP, ip, C.c_int, dp, IP, C.c_int, dp, C.c_long, vp]),
        "spllt_wait": (None, []),
        "spllt_chkerr": (None, [C.c_int, ip, ip, dp, C.c_int, dp, dp]),
        "spllt_deallocate_fkeep": (None, [vpp, ip]),
        "spllt_deallocate_akeep": (None, [vpp, ip]),
        "spllt_task_manager_init": (None, [vpp]),
        "spllt_task_manager_deallocate": (None, [vpp, ip]),
        "spllt_all": (None, [vpp, vpp, OP, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, dp, dp, dp, IP]),
        # B200 additions
        "spllt_b200_analyse": (None, [vpp, vpp, OP, C.c_int, ip, ip, IP, ip, C.c_int]),
        "spllt_b200_num_factor": (C.c_longlong, [vp]),
        "spllt_b200_num_flops": (C.c_longlong, [vp]),
        "spllt_b200_arena_doubles": (C.c_longlong, [vp]),
        "spllt_b200_num_nodes": (C.c_int, [vp]),
        "spllt_b200_num_bcol": (C.c_int, [vp]),
        "spllt_b200_final_blk": (C.c_longlong, [vp]),
        "spllt_b200_maxmn": (C.c_int, [vp]),
        "spllt_b200_num_depth": (C.c_int, [vp]),
        "spllt_b200_get_symbolic": (None, [vp, ip, ip, llp, ip]),
        "spllt_b200_rlist_len": (C.c_longlong, [vp]),
        "spllt_b200_get_blocks": (None, [vp, llp]),
        "spllt_b200_get_nodes": (None, [vp, llp]),
        "spllt_b200_get_small": (None, [vp, ip]),
        "spllt_b200_get_weight": (None, [vp, llp]),
        "spllt_b200_lmap_len": (C.c_longlong, [vp, C.c_int]),
        "spllt_b200_get_lmap": (None, [vp, C.c_int, llp, llp]),
        "spllt_b200_num_sblocks": (C.c_int, [vp, C.c_int]),
        "spllt_b200_get_sblocks": (None, [vp, C.c_int, ip]),
        "spllt_b200_lcol_size": (C.c_longlong, [vp, C.c_int]),
        "spllt_b200_get_lcol": (None, [vp, C.c_int, dp]),
        "spllt_b200_factor_size": (C.c_longlong, [vp]),
        "spllt_b200_get_factor": (None, [vp, dp]),
        "spllt_b200_set_stream": (None, [vp, vp]),
        "spllt_b200_factor_dev": (None, [vp, vp, vp, IP]),
        "spllt_b200_solve_dev": (None, [vp, C.c_int, vp, C.c_int, C.c_int, IP]),
        "spllt_b200_get_fwd": (None, [vp, C.c_int, dp]),
        "spllt_b200_pivot_flag": (C.c_int, [vp]),
        "spllt_b200_chkerr": (C.c_int, [C.c_int, ip, ip, dp, C.c_int, dp, dp, dp]),
        "spllt_b200_factor_launches": (C.c_longlong, [vp]),
        "spllt_b200_solve_launches": (C.c_longlong, [vp, C.c_int]),
        "spllt_b200_tile_flops": (C.c_double, [vp]),
        "spllt_b200_tile_flops_algo": (C.c_double, [vp]),
        "spllt_b200_launch_breakdown": (None, [vp, llp]),
        "spllt_b200_profile_factor": (None, [vp, vp, dp, C.c_char_p]),
        "spllt_b200_node_owner": (C.c_int, [vp, C.c_int]),
        "spllt_b200_profile_solve": (None, [vp, C.c_int, vp, C.c_int, dp, C.c_char_p]),
        "spllt_b200_pipe_sizes": (None, [vp, C.POINTER(C.c_longlong)]),
        "spllt_b200_pipe_top_sizes": (None, [vp, C.POINTER(C.c_longlong)]),
        "spllt_b200_get_pipe_top": (None, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "spllt_b200_wide_frac": (C.c_double, [vp]),
        "spllt_b200_pipe_max_nrhs": (C.c_int, [vp]),
        "spllt_b200_trace_solve": (None, [vp, C.c_int, vp, C.c_int, C.POINTER(C.c_ulonglong),
                                          C.POINTER(C.c_ulonglong)]),
        "spllt_b200_get_pipe": (None, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                       C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "spllt_b200_peak_probe": (C.c_double, [C.c_int, C.c_int, vp]),
        "spllt_b200_arena_ptr": (vp, [vp]),
        "spllt_b200_partition": (None, [vp, vp, C.c_int, C.c_int]),
        "spllt_b200_partition_host": (None, [vp, C.c_int, C.c_int]),
        "spllt_b200_panel_coverage": (None, [vp, llp]),
        "spllt_b200_num_launch_records": (C.c_longlong, [vp]),
        "spllt_b200_get_launch_records": (None, [vp, llp]),
        "spllt_b200_num_tile_tasks": (C.c_longlong, [vp]),
        "spllt_b200_get_tile_tasks": (None, [vp, llp]),
        "spllt_b200_num_top_steps": (C.c_int, [vp]),
        "spllt_b200_get_top_steps": (None, [vp, ip]),
        "spllt_b200_get_bcol_owner": (None, [vp, ip]),
        "spllt_b200_comm_export": (C.c_int, [vp, vp]),
        "spllt_b200_comm_attach": (C.c_int, [vp, C.c_int, C.c_int, vp]),
        "spllt_b200_emulate_ranks_factor": (C.c_int, [vpp, C.c_int, vp]),
        "spllt_b200_compare_factor": (C.c_int, [vp, vp, vp, vp, dp]),
        "spllt_b200_dist_top": (C.c_int, [vp]),
        "spllt_b200_solve_phase": (None, [vp, C.c_int, vp, C.c_int, C.c_int]),
        "spllt_b200_xw_ptr": (vp, [vp, C.c_int]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)  # AttributeError = the library does not export a declared symbol
        f.restype = res
        f.argtypes = args
    L._signatures = sig
    if path is None:
        _lib = L
    return L


def lib():
    return load_library()


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _llp(a):
    return a.ctypes.data_as(C.POINTER(C.c_longlong))


def chkerr(n, ptr, row, val, x, rhs):
    """Scaled backward errors ||b-Ax|| / (||b|| + max|a_ij| ||x||) per RHS (src/utils_mod.F90:432-478)."""
    x = np.asfortranarray(x, dtype=np.float64).reshape(n, -1, order="F")
    rhs = np.asfortranarray(rhs, dtype=np.float64).reshape(n, -1, order="F")
    nrhs = x.shape[1]
    err = np.zeros(nrhs)
    ok = lib().spllt_b200_chkerr(n, _ip(ptr), _ip(row), _dp(val), nrhs, _dp(x), _dp(rhs), _dp(err))
    return ok, err


class SpLLT:
    """One analysed / factorized matrix: owns the opaque akeep / fkeep handles of the C ABI."""

    def __init__(self, **options):
        self.L = lib()
        self.options = Options(**options)
        self.info = Inform()
        self.akeep = C.c_void_p(None)
        self.fkeep = C.c_void_p(None)
        self.n = 0
        self.order = None
        self._keep = []  # host arrays that must outlive asynchronous calls

    # -------------------------------------------------------------- reference operators
    def analyse(self, n, ptr, row, ordering=ORDER_METIS, order=None):
        self.n = int(n)
        self.ptr = np.ascontiguousarray(ptr, dtype=np.int32)
        self.row = np.ascontiguousarray(row, dtype=np.int32)
        self.order = np.zeros(max(n, 1), dtype=np.int32)
        if ordering == ORDER_USER:
            self.order[:n] = order
        if ordering == ORDER_METIS:
            self.L.spllt_analyse(C.byref(self.akeep), C.byref(self.fkeep), C.byref(self.options), n,
                                 _ip(self.ptr), _ip(self.row), C.byref(self.info), _ip(self.order))
        else:
            self.L.spllt_b200_analyse(C.byref(self.akeep), C.byref(self.fkeep), C.byref(self.options), n,
                                      _ip(self.ptr), _ip(self.row), C.byref(self.info), _ip(self.order), ordering)
        return self.info.flag

    def factor(self, val):
        """Asynchronous, like the reference: call wait() before using the factors."""
        v = np.ascontiguousarray(val, dtype=np.float64)
        self._keep = [v]
        self.L.spllt_factor(self.akeep, self.fkeep, C.byref(self.options), v.size, _dp(v), C.byref(self.info))
        return self.info.flag

    def wait(self):
        self.L.spllt_wait()

    def prepare_solve(self, nrhs, nb=None):
        ws = C.c_long(0)
        nb = self.options.nb if nb is None else nb
        self.L.spllt_prepare_solve(self.akeep, self.fkeep, nb, nrhs, C.byref(ws), C.byref(self.info))
        self.worksize = ws.value
        self.y = np.zeros(max(self.n * nrhs, 1))
        self.workspace = np.zeros(max(self.worksize, 1))
        self.L.spllt_set_mem_solve(self.akeep, self.fkeep, nb, nrhs, ws.value, _dp(self.y), _dp(self.workspace),
                                   C.byref(self.info))
        return self.worksize

    def prepare_solve_size(self, nrhs, nb=None):
        ws = C.c_long(0)
        self.L.spllt_prepare_solve(self.akeep, self.fkeep, self.options.nb if nb is None else nb, nrhs,
                                   C.byref(ws), C.byref(self.info))
        return ws.value

    def solve(self, x, job=0):
        """x: n or n x nrhs (column-major), overwritten in place; job 0/1/2 as in the reference."""
        assert x.dtype == np.float64 and (x.ndim == 1 or x.flags.f_contiguous)
        nrhs = 1 if x.ndim == 1 else x.shape[1]
        self.L.spllt_solve(self.fkeep, C.byref(self.options), _ip(self.order), nrhs, _dp(x), C.byref(self.info), job)
        return self.info.flag

    def solve_worker(self, x, job=0):
        """Asynchronous form (spllt_solve_worker): wait() completes it."""
        nrhs = 1 if x.ndim == 1 else x.shape[1]
        self._keep.append(x)
        self.L.spllt_solve_worker(self.fkeep, C.byref(self.options), _ip(self.order), nrhs, _dp(x),
                                  C.byref(self.info), job, None, 0, None)
        return self.info.flag

    def free(self):
        st = C.c_int(0)
        if self.fkeep:
            self.L.spllt_deallocate_fkeep(C.byref(self.fkeep), C.byref(st))
        if self.akeep:
            self.L.spllt_deallocate_akeep(C.byref(self.akeep), C.byref(st))

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # -------------------------------------------------------------- device-resident path
    def set_stream(self, stream_ptr):
        self.L.spllt_b200_set_stream(self.fkeep, C.c_void_p(stream_ptr))

    def factor_dev(self, d_val_ptr):
        self.L.spllt_b200_factor_dev(self.akeep, self.fkeep, C.c_void_p(d_val_ptr), C.byref(self.info))

    def solve_dev(self, d_x_ptr, nrhs, ldx=None, job=0):
        self.L.spllt_b200_solve_dev(self.fkeep, nrhs, C.c_void_p(d_x_ptr), self.n if ldx is None else ldx, job,
                                    C.byref(self.info))
        return self.info.flag

    def profile_factor(self, d_val_ptr, csv=None):
        ms = np.zeros(4)
        self.L.spllt_b200_profile_factor(self.fkeep, C.c_void_p(d_val_ptr), _dp(ms),
                                         csv.encode() if csv else None)
        return dict(zip(("assemble", "panel", "tile_s", "tile_l"), ms.tolist()))

    def profile_solve(self, d_x_ptr, nrhs, csv=None):
        ms = np.zeros(6)
        self.L.spllt_b200_profile_solve(self.fkeep, nrhs, C.c_void_p(d_x_ptr), self.n, _dp(ms),
                                        csv.encode() if csv else None)
        return dict(zip(("fwd_diag", "fwd_upd", "bwd_upd", "bwd_diag", "fwd_pipe", "bwd_pipe"), ms.tolist()))

    def trace_solve(self, d_x_ptr, nrhs):
        """Per-task time stamps (ns) of the persistent solve kernels: (fwd [k,8], bwd [k,8])."""
        sz = np.zeros(4, np.int64)
        self.L.spllt_b200_pipe_sizes(self.akeep, sz.ctypes.data_as(C.POINTER(C.c_longlong)))
        ch = 1 if nrhs == 1 else (nrhs + 3) // 4
        f = np.zeros((int(sz[0]) * ch, 8), np.uint64)
        b = np.zeros((int(sz[1]) * ch, 8), np.uint64)
        up = lambda a: a.ctypes.data_as(C.POINTER(C.c_ulonglong))
        self.L.spllt_b200_trace_solve(self.fkeep, nrhs, C.c_void_p(d_x_ptr), self.n, up(f), up(b))
        return f, b

    def pipe_top_tables(self):
        """Multi-GPU: work lists of the shared upper tree: (tasks_f [k,6], tasks_b [k,6], expect [nstrips])."""
        sz = np.zeros(4, np.int64)
        self.L.spllt_b200_pipe_sizes(self.akeep, sz.ctypes.data_as(C.POINTER(C.c_longlong)))
        s2 = np.zeros(2, np.int64)
        self.L.spllt_b200_pipe_top_sizes(self.akeep, s2.ctypes.data_as(C.POINTER(C.c_longlong)))
        tf = np.zeros((int(s2[0]), 6), np.int32)
        tb = np.zeros((int(s2[1]), 6), np.int32)
        ex = np.zeros(max(int(sz[2]), 1), np.int32)
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        self.L.spllt_b200_get_pipe_top(self.akeep, ip(tf), ip(tb), ip(ex))
        return tf, tb, ex[:int(sz[2])]

    def pipe_tables(self):
        """Work lists of the pipelined solve:
        (tasks_f [k,6], tasks_b [k,6], nodes [nnodes,8], dest, nstrips, expect [nstrips])."""
        sz = np.zeros(4, np.int64)
        self.L.spllt_b200_pipe_sizes(self.akeep, sz.ctypes.data_as(C.POINTER(C.c_longlong)))
        tf = np.zeros((int(sz[0]), 6), np.int32)
        tb = np.zeros((int(sz[1]), 6), np.int32)
        nd = np.zeros((self.nnodes, 8), np.int32)
        de = np.zeros(max(int(sz[3]), 1), np.int32)
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        ex = np.zeros(max(int(sz[2]), 1), np.int32)
        self.L.spllt_b200_get_pipe(self.akeep, ip(tf), ip(tb), ip(nd), ip(de), ip(ex))
        return tf, tb, nd, de[:int(sz[3])], int(sz[2]), ex[:int(sz[2])]

    def pivot_flag(self):
        return self.L.spllt_b200_pivot_flag(self.fkeep)

    # -------------------------------------------------------------- tables (reference numbering)
    @property
    def nnodes(self):
        return self.L.spllt_b200_num_nodes(self.akeep)

    @property
    def nbcol(self):
        return self.L.spllt_b200_num_bcol(self.akeep)

    @property
    def num_flops(self):
        return self.L.spllt_b200_num_flops(self.akeep)

    @property
    def num_factor(self):
        return self.L.spllt_b200_num_factor(self.akeep)

    def symbolic(self):
        nn = self.nnodes
        sptr = np.zeros(nn + 1, dtype=np.int32)
        sparent = np.zeros(max(nn, 1), dtype=np.int32)
        rptr = np.zeros(nn + 1, dtype=np.int64)
        rlist = np.zeros(max(self.L.spllt_b200_rlist_len(self.akeep), 1), dtype=np.int32)
        self.L.spllt_b200_get_symbolic(self.akeep, _ip(sptr), _ip(sparent), _llp(rptr), _ip(rlist))
        return sptr, sparent[:nn], rptr, rlist[:self.L.spllt_b200_rlist_len(self.akeep)]

    def blocks(self):
        out = np.zeros((max(self.L.spllt_b200_final_blk(self.akeep), 1), 9), dtype=np.int64)
        self.L.spllt_b200_get_blocks(self.akeep, _llp(out))
        return out[:self.L.spllt_b200_final_blk(self.akeep)]

    def nodes(self):
        out = np.zeros((max(self.nnodes, 1), 8), dtype=np.int64)
        self.L.spllt_b200_get_nodes(self.akeep, _llp(out))
        return out[:self.nnodes]

    def small(self):
        out = np.zeros(max(self.nnodes, 1), dtype=np.int32)
        self.L.spllt_b200_get_small(self.akeep, _ip(out))
        return out[:self.nnodes]

    def weight(self):
        out = np.zeros(self.nnodes + 1, dtype=np.int64)
        self.L.spllt_b200_get_weight(self.akeep, _llp(out))
        return out

    def lmap(self, bcol):
        ln = self.L.spllt_b200_lmap_len(self.akeep, bcol)
        dst = np.zeros(max(ln, 1), dtype=np.int64)
        src = np.zeros(max(ln, 1), dtype=np.int64)
        self.L.spllt_b200_get_lmap(self.akeep, bcol, _llp(dst), _llp(src))
        return dst[:ln], src[:ln]

    def sblocks(self, nb=None):
        nb = self.options.nb if nb is None else nb
        cnt = self.L.spllt_b200_num_sblocks(self.akeep, nb)
        out = np.zeros((max(cnt, 1), 9), dtype=np.int32)
        self.L.spllt_b200_get_sblocks(self.akeep, nb, _ip(out))
        return out[:cnt]

    def factor_entries(self):
        """All block columns concatenated in the reference layout (lfact(bcol)%lcol)."""
        out = np.zeros(max(self.L.spllt_b200_factor_size(self.akeep), 1))
        self.L.spllt_b200_get_factor(self.fkeep, _dp(out))
        return out[:self.L.spllt_b200_factor_size(self.akeep)]

    def lcol(self, bcol):
        out = np.zeros(max(self.L.spllt_b200_lcol_size(self.akeep, bcol), 1))
        self.L.spllt_b200_get_lcol(self.fkeep, bcol, _dp(out))
        return out[:self.L.spllt_b200_lcol_size(self.akeep, bcol)]

    def fwd_result(self, nrhs):
        out = np.zeros((self.n, nrhs))
        self.L.spllt_b200_get_fwd(self.fkeep, nrhs, _dp(out))
        return out
