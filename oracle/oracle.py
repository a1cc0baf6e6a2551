"""ctypes wrapper of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py
may import this module.  It is the checker, never the product (see spllt_oracle.cpp header).
"""
import ctypes as C
import glob
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")


def build(force=False):
    src = os.path.join(HERE, "spllt_oracle.cpp")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return LIB


SSI = os.path.join(HERE, "libssids_standin.so")
_ssi = None


def build_ssids():
    srcs = [os.path.join(HERE, "ssids_standin.cpp"), os.path.join(HERE, "..", "spllt_b200", "csrc", "symbolic.cpp")]
    if not os.path.exists(SSI) or any(os.path.exists(f) and os.path.getmtime(SSI) < os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", HERE, "-B", "libssids_standin.so"], stdout=subprocess.DEVNULL)
    return SSI


def symbolic(n, ptr, row, nemin=32, ordering=1):
    """order, sptr, sparent, rptr, rlist (1-based, SSIDS conventions) from the SSIDS stand-in
    (METIS nested dissection + supernodal symbolic factorization), without the product library."""
    global _ssi
    if _ssi is None:
        build_ssids()
        L = C.CDLL(SSI)
        ip, llp = C.POINTER(C.c_int), C.POINTER(C.c_longlong)
        L.ssi_analyse.argtypes = [C.c_int, ip, ip, C.c_int, C.c_int]
        L.ssi_analyse.restype = C.c_void_p
        L.ssi_nnodes.argtypes = [C.c_void_p]
        L.ssi_rlist_len.argtypes = [C.c_void_p]
        L.ssi_rlist_len.restype = C.c_longlong
        L.ssi_get.argtypes = [C.c_void_p, ip, ip, ip, llp, ip]
        L.ssi_free.argtypes = [C.c_void_p]
        _ssi = L
    L = _ssi
    ptr = np.ascontiguousarray(ptr, dtype=np.int32)
    row = np.ascontiguousarray(row, dtype=np.int32)
    h = L.ssi_analyse(n, _ip(ptr), _ip(row), nemin, ordering)
    if not h:
        raise RuntimeError("symbolic analysis failed")
    nn = L.ssi_nnodes(h)
    order = np.zeros(max(n, 1), np.int32)
    sptr = np.zeros(nn + 1, np.int32)
    sparent = np.zeros(max(nn, 1), np.int32)
    rptr = np.zeros(nn + 1, np.int64)
    rlist = np.zeros(max(L.ssi_rlist_len(h), 1), np.int32)
    L.ssi_get(h, _ip(order), _ip(sptr), _ip(sparent), _llp(rptr), _ip(rlist))
    ln = L.ssi_rlist_len(h)
    L.ssi_free(h)
    return order[:n], sptr, sparent[:nn], rptr, rlist[:ln]


def _blas_path():
    import scipy
    c = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so"))
    if not c:
        raise RuntimeError("scipy's bundled OpenBLAS not found")
    return c[0]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        vp, ip, dp, llp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_longlong)
        L.orc_init_blas.argtypes = [C.c_char_p]
        L.orc_init_blas.restype = C.c_int
        L.orc_analyse.argtypes = [C.c_int, ip, ip, ip, C.c_int, ip, ip, llp, ip, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_analyse.restype = vp
        L.orc_free.argtypes = [vp]
        for f, r in (("orc_final_blk", C.c_longlong), ("orc_nbcol", C.c_int), ("orc_maxmn", C.c_int),
                     ("orc_factor_size", C.c_longlong), ("orc_num_sblocks", C.c_int)):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = r
        L.orc_get_blocks.argtypes = [vp, llp]
        L.orc_get_nodes.argtypes = [vp, llp]
        L.orc_get_small.argtypes = [vp, ip]
        L.orc_get_weight.argtypes = [vp, llp]
        L.orc_lmap_len.argtypes = [vp, C.c_int]
        L.orc_lmap_len.restype = C.c_longlong
        L.orc_get_lmap.argtypes = [vp, C.c_int, llp, llp]
        L.orc_lcol_size.argtypes = [vp, C.c_int]
        L.orc_lcol_size.restype = C.c_longlong
        L.orc_get_lcol.argtypes = [vp, C.c_int, dp]
        L.orc_get_factor.argtypes = [vp, dp]
        L.orc_factor.argtypes = [vp, dp, C.c_int]
        L.orc_factor_prefix.argtypes = [vp, dp, C.c_int, C.c_int]
        L.orc_prepare_solve.argtypes = [vp, C.c_int, C.c_int]
        L.orc_prepare_solve.restype = C.c_longlong
        L.orc_get_sblocks.argtypes = [vp, ip]
        L.orc_solve.argtypes = [vp, C.c_int, dp, C.c_int]
        L.orc_solve.restype = C.c_int
        L.orc_get_y.argtypes = [vp, dp]
        L.orc_chkerr.argtypes = [C.c_int, ip, ip, dp, C.c_int, dp, dp, dp]
        L.orc_chkerr.restype = C.c_int
        if L.orc_init_blas(_blas_path().encode()) != 0:
            raise RuntimeError("oracle: cannot load BLAS")
        _lib = L
    return _lib


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _llp(a):
    return a.ctypes.data_as(C.POINTER(C.c_longlong))


class Oracle:
    """CPU restatement of analyse (post-SSIDS) / factor / solve on given symbolic inputs."""

    def __init__(self, n, ptr, row, order, sptr, sparent, rptr, rlist, nb, ncpu=1, prune=1, min_width_blas=8):
        self.L = lib()
        self.n = n
        self.nnodes = len(sptr) - 1
        self.nb = nb
        a = [np.ascontiguousarray(v, dtype=np.int32) for v in (ptr, row, order, sptr, sparent, rlist)]
        self.ptr, self.row = a[0], a[1]
        rptr = np.ascontiguousarray(rptr, dtype=np.int64)
        self.h = self.L.orc_analyse(n, _ip(a[0]), _ip(a[1]), _ip(a[2]), self.nnodes, _ip(a[3]), _ip(a[4]),
                                    _llp(rptr), _ip(a[5]), nb, ncpu, prune, min_width_blas)

    def __del__(self):
        try:
            self.L.orc_free(self.h)
        except Exception:
            pass

    def blocks(self):
        out = np.zeros((max(self.L.orc_final_blk(self.h), 1), 9), dtype=np.int64)
        self.L.orc_get_blocks(self.h, _llp(out))
        return out[:self.L.orc_final_blk(self.h)]

    def nodes(self):
        out = np.zeros((max(self.nnodes, 1), 8), dtype=np.int64)
        self.L.orc_get_nodes(self.h, _llp(out))
        return out[:self.nnodes]

    def small(self):
        out = np.zeros(max(self.nnodes, 1), dtype=np.int32)
        self.L.orc_get_small(self.h, _ip(out))
        return out[:self.nnodes]

    def weight(self):
        out = np.zeros(self.nnodes + 1, dtype=np.int64)
        self.L.orc_get_weight(self.h, _llp(out))
        return out

    def nbcol(self):
        return self.L.orc_nbcol(self.h)

    def maxmn(self):
        return self.L.orc_maxmn(self.h)

    def lmap(self, bcol):
        ln = self.L.orc_lmap_len(self.h, bcol)
        dst = np.zeros(max(ln, 1), dtype=np.int64)
        src = np.zeros(max(ln, 1), dtype=np.int64)
        self.L.orc_get_lmap(self.h, bcol, _llp(dst), _llp(src))
        return dst[:ln], src[:ln]

    def factor(self, val, nthreads=1):
        v = np.ascontiguousarray(val, dtype=np.float64)
        self.L.orc_factor(self.h, _dp(v), nthreads)

    def factor_prefix(self, val, nthreads, last_node):
        """Bounded sample (bench.py's CPU arm): factorizes nodes 1..last_node only."""
        v = np.ascontiguousarray(val, dtype=np.float64)
        self.L.orc_factor_prefix(self.h, _dp(v), nthreads, int(last_node))

    def factor_entries(self):
        out = np.zeros(max(self.L.orc_factor_size(self.h), 1))
        self.L.orc_get_factor(self.h, _dp(out))
        return out[:self.L.orc_factor_size(self.h)]

    def lcol(self, bcol):
        out = np.zeros(max(self.L.orc_lcol_size(self.h, bcol), 1))
        self.L.orc_get_lcol(self.h, bcol, _dp(out))
        return out[:self.L.orc_lcol_size(self.h, bcol)]

    def prepare_solve(self, nrhs, nb=None):
        return self.L.orc_prepare_solve(self.h, self.nb if nb is None else nb, nrhs)

    def sblocks(self):
        cnt = self.L.orc_num_sblocks(self.h)
        out = np.zeros((max(cnt, 1), 9), dtype=np.int32)
        self.L.orc_get_sblocks(self.h, _ip(out))
        return out[:cnt]

    def solve(self, x, job=0):
        assert x.dtype == np.float64 and (x.ndim == 1 or x.flags.f_contiguous)
        nrhs = 1 if x.ndim == 1 else x.shape[1]
        return self.L.orc_solve(self.h, nrhs, _dp(x), job)

    def y(self, nrhs):
        out = np.zeros(self.n * nrhs)
        self.L.orc_get_y(self.h, _dp(out))
        return out


def chkerr(n, ptr, row, val, x, rhs):
    L = lib()
    x = np.asfortranarray(x, dtype=np.float64).reshape(n, -1, order="F")
    rhs = np.asfortranarray(rhs, dtype=np.float64).reshape(n, -1, order="F")
    nrhs = x.shape[1]
    err = np.zeros(nrhs)
    ptr = np.ascontiguousarray(ptr, dtype=np.int32)
    row = np.ascontiguousarray(row, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    ok = L.orc_chkerr(n, _ip(ptr), _ip(row), _dp(val), nrhs, _dp(x), _dp(rhs), _dp(err))
    return ok, err
