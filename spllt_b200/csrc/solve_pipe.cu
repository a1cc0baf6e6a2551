// Pipelined supernodal triangular solves: ONE persistent kernel per sweep (sm_100a).
//
// Replaces, for the solve phase, the reference's task DAG over solve tiles
// (solve_fwd / solve_bwd, src/spllt_solve_mod.F90:244-411; fwd_wdep / bwd_wdep dependency
// sets, src/spllt_solve_dep_mod.F90:27-248; task bodies
// src/include/spllt_solve_{fwd,bwd}_{block,update}_worker.F90.inc) and its kernels
// slv_solve / slv_fwd_update / slv_bwd_update (src/spllt_solve_kernels_mod.F90:11-210) and
// fwd_update_upd / bwd_update_upd (src/spllt_solve_dep_mod.F90:1684-1761).
//
// Design.  A level-set schedule pays one kernel boundary (>= 10 us with its tail) per block
// column on the critical path of the assembly tree and solves a 768-wide diagonal block on a
// single SM.  Here the unit is a 64-row STRIP, independent of the factorization's nb:
//
//   forward, node s with n columns, m rows, np = ceil(n/64) strips
//     DIAG(s,i)   owner of rows [64i, 64i+64) of the diagonal block.  Left-looking: streams the
//                 tiles L[strip i, strip j], j < i, as the x_j are published, accumulates in
//                 registers (no atomics), multiplies by the precomputed INVERSE of the 64 x 64
//                 diagonal block (k_invert_diag, once per factorization: a matrix-vector product
//                 instead of a 64-step substitution on the critical chain), publishes x_i.
//     BELOW(s,c)  rows below the diagonal block: the same streaming product, then
//                 xw[index[r]] -= sum (RED.ADD.F64) and one counter bump per ancestor STRIP hit.
//                 Nodes of <= 4 strips: x_s goes to shared memory once, chunks of up to 512 rows;
//                 wider nodes: 64-row chunks that walk the strips as they are published.
//     DIAG(s,i) reads its right-hand side once its strip's counter shows that every
//     contribution to those 64 rows has landed (so a parent starts after its child's first chunk).
//   backward: the mirror image (gather only): a BELOW chunk waits for the flags of the ancestor
//     strips its rows map to, gathers xw[index[r]], adds L^T y into the node's x (RED) and bumps
//     the node's counter; DIAG(s,i) streams L[strip j, strip i]^T x_j for j > i in decreasing j.
//   nodes with n <= 64 and few rows are ONE fused task (SMALL).
//
// Synchronisation, all in HBM: a flag per strip (st after __threadfence), counters
// (ld.acquire.gpu), and -- for the strip-to-strip chain inside a node -- a MAILBOX copy of x made
// of self-validating 8-byte words (see ld_mailbox), which takes the membar, the flag and one
// dependent load off every link of the chain.  Waiters far from the front back off (nanosleep).
//
// CTAs claim tasks from a global counter in list order; lists are topological, so a claimed
// task only waits on tasks already claimed by running CTAs: no deadlock, no co-residency
// requirement.  Every L entry is read exactly once per sweep, 512 contiguous bytes per row and
// warp instruction; HBM latency is taken by prefetch.global.L2 (a task's rows when it starts,
// tiles two steps ahead), the register loads of a tile are issued while the warp polls for x_j.
// Multi-GPU: the same kernels run on per-rank lists (own subtrees / shared upper tree), see
// Engine::solve_phase.
#include <cstdio>
#include <cstdlib>

#include "kernels.cuh"

#include "cuda_check.h"

namespace spllt {


namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int PT = 256;            // threads per CTA
constexpr int PWARPS = PT / 32;    // 8
constexpr int RPW = PS / PWARPS;   // 8 rows of a strip per warp
constexpr int PSL = PS + 1;        // padded leading dimension of the diagonal block in shared memory
#ifndef PIPE_OCC
#define PIPE_OCC 3   // resident CTAs per SM of the single-rhs kernels (register cap 85)
#endif
constexpr int FAT_NP = PIPE_FAT_NP;
constexpr int PREFETCH_BYTES = 128 * 1024;   // per task
static_assert(PS * PIPE_RC <= PT, "one right-hand-side entry of a strip per thread");

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_flag(int* p, int v) {
  asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// streaming read of the factor: read-only for the whole sweep, every entry used once
__device__ __forceinline__ double2 ld_stream2(const double* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void cp_async8z(void* s, const void* g, int bytes) {
  unsigned a = (unsigned)__cvta_generic_to_shared(s);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(a), "l"(g), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_commit_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// fire-and-forget reductions (SASS REDG): nothing here needs the old value, and a returning ATOMG
// keeps a register scoreboard busy until L2 answers (up to microseconds on contended lines)
__device__ __forceinline__ void red_f64(double* p, double v) {
  asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void red_s32(int* p, int v) {
  asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Mailbox: a second copy of x (xm, pre-set to an all-ones NaN pattern by a memset) whose entries are
// SELF-VALIDATING 8-byte words.  A strip writes its x there straight after the matrix-vector
// product -- before the barrier, the fence and the flag that everybody else waits for -- and the
// strips / rows below of the same node poll the values themselves: the critical chain of a node
// saves a membar and a dependent load per 64 columns.
__device__ __forceinline__ double ld_mailbox(const double* p) {
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_mailbox(double* p, double v) {
  if (__double_as_longlong(v) == -1LL) v = __longlong_as_double(0x7ff8000000000000LL);   // never publish the "empty" pattern
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ bool mailbox_empty(double v) { return __double_as_longlong(v) == -1LL; }

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// experiment switches (PipeArgs::mode, SPLLT_B200_PIPE_MODE):
//   1: fence.acq_rel.gpu instead of __threadfence() (MEMBAR.SC) before a flag / counter is raised
//   2: waiters do NOT back off (default: nanosleep in proportion to their distance from the critical path)
//   4: poll with ld.relaxed (no L1 invalidation per poll); x is read with L2-coherent loads behind
//      the control dependency   8: ... plus one fence.acq_rel after a successful poll
//  16: no L2 prefetch of a task's rows at task start   32: strips wait on flags, not on the mailbox
enum { M_FENCE_ACQREL = 1, M_NO_BACKOFF = 2, M_POLL_RELAXED = 4, M_POLL_FENCE = 8, M_NO_PREFETCH = 16, M_NO_MAILBOX = 32, M_NO_MAILBOX_BWD = 64 };
__device__ __forceinline__ void fence_gpu(int mode) {
  if (mode & M_FENCE_ACQREL)
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  else
    __threadfence();
}

// Rows [r0, r0 + nrows) x all n columns of a node -> L2, without registers: issued before a task
// starts waiting, so that the streaming loops afterwards hit L2 (~0.3 us) instead of HBM (~1 us)
// and more bytes are in flight than the register double-buffering alone allows.
__device__ __forceinline__ void prefetch_l2(const double* base, i64 ld, int nrows, int ncols) {
  const int row_bytes = ncols * 8;
  const int rows = min(nrows, max(8, PREFETCH_BYTES / max(row_bytes, 1)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < rows; r += PWARPS) {
    const char* p = (const char*)(base + (i64)r * ld);
    for (int off = lane * 128; off < row_bytes; off += 32 * 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
  }
}

// One instruction per warp: this warp's 8 rows x 64 columns (4 lines of 128 B per row) of a tile
// -> L2.  The register loads of a tile are issued when its turn comes (while the warp polls for
// the matching x_j); the HBM latency is taken two tiles earlier by this prefetch.
__device__ __forceinline__ void prefetch_tile(const double* row0, i64 ld, int rows_valid, int cols_valid, int lane) {
  const int u = lane >> 2, seg = lane & 3;
  if (u < rows_valid && seg * 16 < cols_valid) asm volatile("prefetch.global.L2 [%0];" ::"l"(row0 + (i64)u * ld + seg * 16));
}

// Flags f[0], f[dir], f[2 dir], ... (at most `limit` of them matter).  Spins until the first one
// is raised and returns how many consecutive ones are (1..32): the caller skips polling for those.
// dist: how many publications this waiter is away from being on the critical path (>= 1).
__device__ __forceinline__ int wait_run(const int* f, int dir, int limit, int lane, int mode, int dist) {
  for (;;) {
    int v = 0;
    if (lane < limit) v = (mode & M_POLL_RELAXED) ? ld_relaxed(f + dir * lane) : ld_acquire(f + dir * lane);
    unsigned b = __ballot_sync(FULL, v != 0);
    if (b & 1u) {
      if (mode & M_POLL_FENCE) asm volatile("fence.acq_rel.gpu;" ::: "memory");
      __syncwarp();
      return b == FULL ? 32 : __ffs(~b) - 1;
    }
    if (!(mode & M_NO_BACKOFF) && dist > 1) __nanosleep(min(dist - 1, 16) * 200);
  }
}
__device__ __forceinline__ void wait_count(const int* c, int expect, int mode) {
  if (mode & M_POLL_RELAXED) {
    while (ld_relaxed(c) < expect) {
    }
    if (mode & M_POLL_FENCE) asm volatile("fence.acq_rel.gpu;" ::: "memory");
  } else {
    while (ld_acquire(c) < expect) {
    }
  }
}

struct PipeArgs {
  const PTaskD* tasks;
  const int* dest;
  const int* expect;   // [nstrips] forward: contributions a strip waits for
  const double* arena;
  const double* dinv;
  const int* index;
  double* xw;
  double* xm;       // mailbox copy of x (see ld_mailbox), n x nrhs, all-ones pattern = not yet published
  int* sync;        // [0] claim counter; per chunk c: flags at 32 + c*stride, node counters after the flags
  int ntasks, nrhs, nchunk, stride, nstrips;
  int mode;                    // experiment switches, see M_*
  unsigned long long* trace;   // optional: 4 time stamps per claimed task
};

extern __shared__ __align__(16) double pipe_sm[];
// shared-memory layout (doubles): inverse diagonal block, backward partial sums, strip rhs, x of the node
template <int RC>
struct Sm {
  static constexpr int LS = 0;                         // [PS][PSL]
  static constexpr int RED = PS * PSL;                 // [PWARPS][PS][RC]
  static constexpr int RH = RED + PWARPS * PS * RC;    // [PS][RC]   rhs of the strip / gathered y of a below pass
  static constexpr int XS = RH + PS * RC;              // [FAT_NP * PS][RC]  x of the strip / of a narrow node
  static constexpr int TOTAL = XS + FAT_NP * PS * RC;
};

struct Ctx {      // per task; everything else is read from the kernel parameters / shared memory
  int rc0, nr;    // first right-hand side of this pass, how many (<= RC)
  int* flags;
  int* cnt;
  unsigned long long* tr;   // optional time stamps (thread 0): start, waits done, solved, end
};

// inverse of the diagonal block of strip i -> shared memory (lower triangle, rest zero), asynchronously
__device__ __forceinline__ void fetch_diag(const PNode& nd, int i, int rw, const PipeArgs& a, const Ctx& c) {
  const double* D = a.dinv + (i64)(nd.strip0 + i) * (PS * PS);
  for (int idx = threadIdx.x; idx < PS * PS; idx += PT) {
    const int r = idx >> 6, cc = idx & (PS - 1);
    const bool ok = cc <= r && r < rw;
    cp_async8z((pipe_sm + Sm<1>::LS) + r * PSL + cc, ok ? D + idx : D, ok ? 8 : 0);
  }
}

// x_i is final in global memory: make it visible, then raise the strip's flag
__device__ __forceinline__ void publish(int* flag, const PipeArgs& a, const Ctx& c) {
  __syncthreads();
  if (threadIdx.x == 0) {
    if (c.tr) c.tr[2] = gtime();
    fence_gpu(a.mode);
    st_flag(flag, 1);
    if (c.tr) c.tr[7] = gtime();
  }
}

// x = Linv * rh (TRANS: Linv^T * rh) for one strip: thread r < 64 owns row r, four independent
// accumulation chains; Ls is zero above the diagonal and rh is zero beyond the strip's width.
template <int RC, bool TRANS>
__device__ __forceinline__ void strip_matvec(int rw, double* xg, const PipeArgs& a, const Ctx& c) {
  const int r = threadIdx.x;
  if (r < PS) {
    double s[4][RC];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < RC; ++q) s[p][q] = 0.0;
#pragma unroll 4
    for (int k = 0; k < PS; k += 4) {
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const double l = TRANS ? (pipe_sm + Sm<1>::LS)[(k + p) * PSL + r] : (pipe_sm + Sm<1>::LS)[r * PSL + k + p];
#pragma unroll
        for (int q = 0; q < RC; ++q) s[p][q] = fma(l, (pipe_sm + Sm<RC>::RH)[(k + p) * RC + q], s[p][q]);
      }
    }
#pragma unroll
    for (int q = 0; q < RC; ++q) {
      const double x = (s[0][q] + s[1][q]) + (s[2][q] + s[3][q]);
      (pipe_sm + Sm<RC>::XS)[r * RC + q] = x;
      if (r < rw && q < c.nr) {
        st_mailbox(a.xm + (xg - a.xw) + (i64)r * a.nrhs + q, x);
        __stcg(xg + (i64)r * a.nrhs + q, x);
      }
    }
  }
}

// ------------------------------------------------------------------------------ forward
// DIAG(s, i), forward (a13 forward body + the intra-node part of a14).
template <int RC>
__device__ __forceinline__ void fwd_strip(const PNode& nd, int node, int i, bool pub, const PipeArgs& a, const Ctx& c) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = i * PS, rw = min(PS, nd.n - r0);
  const i64 ld = nd.ld;
  fetch_diag(nd, i, rw, a, c);
  const int expect = a.expect[nd.strip0 + i];   // tasks of descendants that add into this strip's rows
  // right-hand side of the strip: lane u < 8 of each warp owns row 8 warp + u.  It is final once
  // every contribution has landed -- usually long before the x_j arrive, so fetch it now.
  double* xg = a.xw + (i64)(nd.sa + r0) * a.nrhs + c.rc0;
  const int row = warp * RPW + lane;
  double b[RC];
  int have_b = 1;
  if (expect > 0) {
    have_b = lane == 0 ? (ld_acquire(c.cnt + nd.strip0 + i) >= expect) : 0;
    have_b = __shfl_sync(FULL, have_b, 0);
  }
  if (have_b) {
#pragma unroll
    for (int q = 0; q < RC; ++q)
      b[q] = (lane < RPW && row < rw && q < c.nr) ? __ldcg(xg + (i64)row * a.nrhs + q) : 0.0;
  }
  double acc[RPW][RC];
#pragma unroll
  for (int u = 0; u < RPW; ++u)
#pragma unroll
    for (int q = 0; q < RC; ++q) acc[u][q] = 0.0;
  if (i > 0) {
    const double* Lw = a.arena + nd.off + (i64)(r0 + warp * RPW) * ld;   // this warp's first row
    const double* Lr = Lw + 2 * lane;
    const int rv = rw - warp * RPW;                                      // its valid rows
    prefetch_tile(Lw, ld, rv, PS, lane);
    if (i > 1) prefetch_tile(Lw + PS, ld, rv, PS, lane);
    double2 t[RPW];
    int ready = 0;
    for (int j = 0; j < i; ++j) {
      if (j + 2 < i) prefetch_tile(Lw + (j + 2) * PS, ld, rv, PS, lane);
#pragma unroll
      for (int u = 0; u < RPW; ++u) t[u] = (u < rv) ? ld_stream2(Lr + (i64)u * ld + j * PS) : make_double2(0.0, 0.0);
      if (j >= ready) {
        if (!have_b) {   // about to wait anyway: have the contributions landed meanwhile?
          have_b = lane == 0 ? (ld_acquire(c.cnt + nd.strip0 + i) >= expect) : 0;
          have_b = __shfl_sync(FULL, have_b, 0);
          if (have_b) {
#pragma unroll
            for (int q = 0; q < RC; ++q)
              b[q] = (lane < RPW && row < rw && q < c.nr) ? __ldcg(xg + (i64)row * a.nrhs + q) : 0.0;
          }
        }
        if (a.mode & M_NO_MAILBOX) ready = j + wait_run(c.flags + nd.strip0 + j, 1, i - j, lane, a.mode, i - j);
      }
      double x0[RC], x1[RC];
      if (a.mode & M_NO_MAILBOX) {
        const double* xp = a.xw + (i64)(nd.sa + j * PS + 2 * lane) * a.nrhs + c.rc0;
#pragma unroll
        for (int q = 0; q < RC; ++q) {
          x0[q] = q < c.nr ? __ldcg(xp + q) : 0.0;
          x1[q] = q < c.nr ? __ldcg(xp + a.nrhs + q) : 0.0;
        }
      } else {
        const double* xp = a.xm + (i64)(nd.sa + j * PS + 2 * lane) * a.nrhs + c.rc0;
        for (;;) {   // poll the values themselves
          bool empty = false;
#pragma unroll
          for (int q = 0; q < RC; ++q) {
            x0[q] = q < c.nr ? ld_mailbox(xp + q) : 0.0;
            x1[q] = q < c.nr ? ld_mailbox(xp + a.nrhs + q) : 0.0;
            empty |= mailbox_empty(x0[q]) | mailbox_empty(x1[q]);
          }
          if (!__any_sync(FULL, empty)) break;
          if (!(a.mode & M_NO_BACKOFF) && i - j > 2) __nanosleep(min(i - j, 24) * 64);   // far from the front: poll rarely
        }
        ready = j + 1;
      }
      if (c.tr && tid == 0 && j == i - 1) c.tr[4] = gtime();
#pragma unroll
      for (int u = 0; u < RPW; ++u)
#pragma unroll
        for (int q = 0; q < RC; ++q) acc[u][q] = fma(t[u].x, x0[q], fma(t[u].y, x1[q], acc[u][q]));
    }
  }
  if (c.tr && tid == 0) c.tr[5] = gtime() + (long long)(acc[0][0] == 12345.678);   // after the last FMA
  if (!have_b) {   // every contribution to these 64 rows has landed (each warp reads its own rows)
    if (lane == 0) wait_count(c.cnt + nd.strip0 + i, expect, a.mode);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < RC; ++q)
      b[q] = (lane < RPW && row < rw && q < c.nr) ? __ldcg(xg + (i64)row * a.nrhs + q) : 0.0;
  }
  {
    if (i > 0) {
#pragma unroll
      for (int u = 0; u < RPW; ++u)
#pragma unroll
        for (int q = 0; q < RC; ++q) {
          double v = acc[u][q];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
          if (lane == u) b[q] -= v;
        }
    }
    if (lane < RPW) {
#pragma unroll
      for (int q = 0; q < RC; ++q) (pipe_sm + Sm<RC>::RH)[row * RC + q] = b[q];
    }
  }
  if (c.tr && tid == 0) c.tr[1] = gtime();
  cp_commit_wait_all();
  __syncthreads();
  if (c.tr && tid == 0) c.tr[6] = gtime();
  strip_matvec<RC, false>(rw, xg, a, c);
  if (pub)
    publish(c.flags + nd.strip0 + i, a, c);
  else
    __syncthreads();
}

// BELOW(s, rows [r0, r0+nrows)), forward, nodes with many strips (a14 slv_fwd_update + a15
// fwd_update_upd): xw[index[r]] -= L[r, 0..n) x_s, streaming over the node's strips as they are
// published; nrows <= 64.
template <int RC>
__device__ __forceinline__ void fwd_below(const PNode& nd, int r0, int nrows, const PipeArgs& a, const Ctx& c) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const i64 ld = nd.ld;
  const int np = nd.np;
  const double* Lr = a.arena + nd.off + (i64)(r0 + warp * RPW) * ld + 2 * lane;
  int myidx = 0;
  if (lane < RPW && warp * RPW + lane < nrows) myidx = a.index[nd.idx_off + r0 + warp * RPW + lane];
  double acc[RPW][RC];
#pragma unroll
  for (int u = 0; u < RPW; ++u)
#pragma unroll
    for (int q = 0; q < RC; ++q) acc[u][q] = 0.0;
  const double* Lw = Lr - 2 * lane;       // this warp's first row
  const int rv = nrows - warp * RPW;      // its valid rows
  prefetch_tile(Lw, ld, rv, nd.ld, lane);
  if (np > 1) prefetch_tile(Lw + PS, ld, rv, nd.ld - PS, lane);
  double2 t[RPW];
  int ready = 0;
  for (int j = 0; j < np; ++j) {
    if (j + 2 < np) prefetch_tile(Lw + (j + 2) * PS, ld, rv, nd.ld - (j + 2) * PS, lane);
    const bool colok = j * PS + 2 * lane < nd.ld;
#pragma unroll
    for (int u = 0; u < RPW; ++u)
      t[u] = (u < rv && colok) ? ld_stream2(Lr + (i64)u * ld + j * PS) : make_double2(0.0, 0.0);
    const int col = j * PS + 2 * lane;
    double x0[RC], x1[RC];
    if (a.mode & M_NO_MAILBOX) {
      if (j >= ready) ready = j + wait_run(c.flags + nd.strip0 + j, 1, np - j, lane, a.mode, np - j);
      const double* xp = a.xw + (i64)(nd.sa + col) * a.nrhs + c.rc0;
#pragma unroll
      for (int q = 0; q < RC; ++q) {
        x0[q] = (q < c.nr && col < nd.n) ? __ldcg(xp + q) : 0.0;
        x1[q] = (q < c.nr && col + 1 < nd.n) ? __ldcg(xp + a.nrhs + q) : 0.0;
      }
    } else {
      const double* xp = a.xm + (i64)(nd.sa + col) * a.nrhs + c.rc0;
      for (;;) {
        bool empty = false;
#pragma unroll
        for (int q = 0; q < RC; ++q) {
          x0[q] = (q < c.nr && col < nd.n) ? ld_mailbox(xp + q) : 0.0;
          x1[q] = (q < c.nr && col + 1 < nd.n) ? ld_mailbox(xp + a.nrhs + q) : 0.0;
          empty |= mailbox_empty(x0[q]) | mailbox_empty(x1[q]);
        }
        if (!__any_sync(FULL, empty)) break;
        if (!(a.mode & M_NO_BACKOFF)) __nanosleep(256);   // rows below a wide node are never on the critical chain of strips
      }
    }
#pragma unroll
    for (int u = 0; u < RPW; ++u)
#pragma unroll
      for (int q = 0; q < RC; ++q) acc[u][q] = fma(t[u].x, x0[q], fma(t[u].y, x1[q], acc[u][q]));
  }
  double mine[RC];
#pragma unroll
  for (int q = 0; q < RC; ++q) mine[q] = 0.0;
#pragma unroll
  for (int u = 0; u < RPW; ++u)
#pragma unroll
    for (int q = 0; q < RC; ++q) {
      double v = acc[u][q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
      if (lane == u) mine[q] = v;
    }
  if (lane < RPW && warp * RPW + lane < nrows) {
    double* dst = a.xw + (i64)myidx * a.nrhs + c.rc0;
#pragma unroll
    for (int q = 0; q < RC; ++q)
      if (q < c.nr) red_f64(dst + q, -mine[q]);
  }
}

// BELOW, forward, nodes with at most FAT_NP strips (most of L's bytes live in the rows below
// such nodes): the whole x_s is taken into registers once (from shared memory for the fused
// SMALL task, else from global memory after all of the node's flags are up), then up to
// PIPE_FAT_ROWS rows are streamed in passes of 64 with the next tile's loads always in flight.
template <int RC>
__device__ __forceinline__ void fwd_below_fat(const PNode& nd, int r0, int nrows, bool local, const PipeArgs& a, const Ctx& c) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const i64 ld = nd.ld;
  const int np = nd.np;
  const double* Lr = a.arena + nd.off + (i64)(r0 + warp * RPW) * ld + 2 * lane;
  auto load_tile = [&](double2(&d)[RPW], int p, int j) {
    const bool colok = j * PS + 2 * lane < nd.ld;
#pragma unroll
    for (int u = 0; u < RPW; ++u)
      d[u] = (p * PS + warp * RPW + u < nrows && colok) ? ld_stream2(Lr + (i64)(p * PS + u) * ld + j * PS)
                                                        : make_double2(0.0, 0.0);
  };
  double2 t[RPW];
  if (!local) {
    // x_s (n <= FAT_NP * 64 values per right-hand side) -> shared memory, once
    if (a.mode & M_NO_MAILBOX) {
      if (warp == 0) {
        int ready = 0;
        while (ready < np) ready += wait_run(c.flags + nd.strip0 + ready, 1, np - ready, lane, a.mode, 1);
      }
      __syncthreads();
      for (int k = threadIdx.x; k < np * PS * RC; k += PT) {
        const int col = k / RC, q = k - col * RC;
        (pipe_sm + Sm<RC>::XS)[k] = (col < nd.n && q < c.nr) ? __ldcg(a.xw + (i64)(nd.sa + col) * a.nrhs + c.rc0 + q) : 0.0;
      }
    } else {   // every thread polls its own entries of the mailbox: no flag, no fence on the path
      for (int k = threadIdx.x; k < np * PS * RC; k += PT) {
        const int col = k / RC, q = k - col * RC;
        double v = 0.0;
        if (col < nd.n && q < c.nr) {
          const double* mp = a.xm + (i64)(nd.sa + col) * a.nrhs + c.rc0 + q;
          while (mailbox_empty(v = ld_mailbox(mp))) {
          }
        }
        (pipe_sm + Sm<RC>::XS)[k] = v;
      }
    }
    __syncthreads();
  }
  if (c.tr && threadIdx.x == 0) c.tr[1] = gtime();
  const int npass = (nrows + PS - 1) / PS;
  for (int p = 0; p < npass; ++p) {
    const int row = p * PS + warp * RPW + lane;
    int myidx = 0;
    if (lane < RPW && row < nrows) myidx = a.index[nd.idx_off + r0 + row];
    double acc[RPW][RC];
#pragma unroll
    for (int u = 0; u < RPW; ++u)
#pragma unroll
      for (int q = 0; q < RC; ++q) acc[u][q] = 0.0;
#pragma unroll
    for (int j = 0; j < FAT_NP; ++j) {
      if (j < np) {
        load_tile(t, p, j);
        if (p + 1 < npass)   // next pass of this warp's rows -> L2 (the first 128 KB came in at task start)
          prefetch_tile(Lr - 2 * lane + (i64)((p + 1) * PS) * ld + j * PS, ld, nrows - (p + 1) * PS - warp * RPW,
                        nd.ld - j * PS, lane);
#pragma unroll
        for (int u = 0; u < RPW; ++u)
#pragma unroll
          for (int q = 0; q < RC; ++q) {
            const double x0 = (pipe_sm + Sm<RC>::XS)[(j * PS + 2 * lane) * RC + q], x1 = (pipe_sm + Sm<RC>::XS)[(j * PS + 2 * lane + 1) * RC + q];
            acc[u][q] = fma(t[u].x, x0, fma(t[u].y, x1, acc[u][q]));
          }
      }
    }
    double mine[RC];
#pragma unroll
    for (int q = 0; q < RC; ++q) mine[q] = 0.0;
#pragma unroll
    for (int u = 0; u < RPW; ++u)
#pragma unroll
      for (int q = 0; q < RC; ++q) {
        double v = acc[u][q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (lane == u) mine[q] = v;
      }
    if (lane < RPW && row < nrows) {
      double* dst = a.xw + (i64)myidx * a.nrhs + c.rc0;
#pragma unroll
      for (int q = 0; q < RC; ++q)
        if (q < c.nr) red_f64(dst + q, -mine[q]);
    }
  }
}

// the contributions above are complete: bump the counter of every ancestor node they hit
__device__ __forceinline__ void bump_dests(const int* dest, int count, int* cnt, int mode) {
  __syncthreads();
  for (int k = threadIdx.x; k < count; k += PT) {
    fence_gpu(mode);
    red_s32(cnt + dest[k], 1);
  }
}

// ------------------------------------------------------------------------------ backward
// DIAG(s, i), backward: x_i = L_ii^-T (b_i - sum_{j>i} L[strip j, strip i]^T x_j)
// (a13 backward body + the intra-node part of slv_bwd_update).
template <int RC>
__device__ __forceinline__ void bwd_strip(const PNode& nd, int node, int i, bool wait_below, const PipeArgs& a, const Ctx& c) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = i * PS, cw = min(PS, nd.n - r0);
  const i64 ld = nd.ld;
  const int np = nd.np, nj = np - 1 - i;
  fetch_diag(nd, i, cw, a, c);
  if (wait_below && i == np - 1 && tid == 0 && nd.expect_b > 0) wait_count(c.cnt + node, nd.expect_b, a.mode);
  double a0[RC], a1[RC];
#pragma unroll
  for (int q = 0; q < RC; ++q) a0[q] = a1[q] = 0.0;
  // this thread's entry of the strip's right-hand side (PS * RC <= PT entries); final as soon as
  // any flag of the node is up (the node's last strip waited for the below counter)
  double* xg = a.xw + (i64)(nd.sa + r0) * a.nrhs + c.rc0;
  const int brow = tid / RC, bq = tid - brow * RC;
  const bool bvalid = tid < PS * RC && brow < cw && bq < c.nr;
  double bpre = 0.0;
  if (nj > 0) {
    // tile rows: strip j of the node, this warp's 8 rows; columns: strip i (full, since i < np-1)
    const double* Lw = a.arena + nd.off + (i64)(warp * RPW) * ld + r0;   // this warp's rows of strip 0, column r0
    const double* Lc = Lw + 2 * lane;
    prefetch_tile(Lw + (i64)((np - 1) * PS) * ld, ld, nd.n - (np - 1) * PS - warp * RPW, PS, lane);
    if (nj > 1) prefetch_tile(Lw + (i64)((np - 2) * PS) * ld, ld, RPW, PS, lane);
    double2 t[RPW];
    int ready = 0;
    for (int jj = 0; jj < nj; ++jj) {
      const int j = np - 1 - jj;
      if (jj + 2 < nj) prefetch_tile(Lw + (i64)((j - 2) * PS) * ld, ld, RPW, PS, lane);
      const int rb = j * PS + warp * RPW;
#pragma unroll
      for (int u = 0; u < RPW; ++u)
        t[u] = (rb + u < nd.n) ? ld_stream2(Lc + (i64)(j * PS + u) * ld) : make_double2(0.0, 0.0);
      double xv[RPW][RC];
      if (a.mode & M_NO_MAILBOX_BWD) {
        if (jj >= ready) ready = jj + wait_run(c.flags + nd.strip0 + j, -1, nj - jj, lane, a.mode, nj - jj);
        const double* xp = a.xw + (i64)(nd.sa + rb) * a.nrhs + c.rc0;
#pragma unroll
        for (int u = 0; u < RPW; ++u)
#pragma unroll
          for (int q = 0; q < RC; ++q) xv[u][q] = (rb + u < nd.n && q < c.nr) ? __ldcg(xp + (i64)u * a.nrhs + q) : 0.0;
      } else {
        const double* xp = a.xm + (i64)(nd.sa + rb) * a.nrhs + c.rc0;
        for (;;) {   // the 8 x values of this warp's rows (same addresses in every lane): poll the values
          bool empty = false;
#pragma unroll
          for (int u = 0; u < RPW; ++u)
#pragma unroll
            for (int q = 0; q < RC; ++q) {
              xv[u][q] = (rb + u < nd.n && q < c.nr) ? ld_mailbox(xp + (i64)u * a.nrhs + q) : 0.0;
              empty |= mailbox_empty(xv[u][q]);
            }
          // a warp whose 8 rows lie beyond the (partial) last strip has nothing to poll, but it must
          // not run ahead: its threads read the strip's right-hand side next, which is final only
          // once the node's last strip has started publishing
          if (jj == 0) empty |= mailbox_empty(ld_mailbox(a.xm + (i64)(nd.sa + j * PS) * a.nrhs + c.rc0));
          if (!__any_sync(FULL, empty)) break;
          if (!(a.mode & M_NO_BACKOFF) && j - i > 2) __nanosleep(min(j - i, 24) * 64);
        }
      }
      if (jj == 0 && bvalid) bpre = __ldcg(xg + (i64)brow * a.nrhs + bq);
      if (c.tr && tid == 0 && jj == nj - 1) c.tr[4] = gtime();
#pragma unroll
      for (int u = 0; u < RPW; ++u)
#pragma unroll
        for (int q = 0; q < RC; ++q) {
          a0[q] = fma(t[u].x, xv[u][q], a0[q]);
          a1[q] = fma(t[u].y, xv[u][q], a1[q]);
        }
    }
  }
  if (c.tr && tid == 0) c.tr[5] = gtime() + (long long)(a0[0] == 12345.678);
#pragma unroll
  for (int q = 0; q < RC; ++q) {
    (pipe_sm + Sm<RC>::RED)[(warp * PS + 2 * lane) * RC + q] = a0[q];
    (pipe_sm + Sm<RC>::RED)[(warp * PS + 2 * lane + 1) * RC + q] = a1[q];
  }
  if (c.tr && tid == 0) c.tr[1] = gtime();
  __syncthreads();   // partial sums complete; for i == np-1 also: thread 0 has seen the below counter
  if (nj == 0 && bvalid) bpre = __ldcg(xg + (i64)brow * a.nrhs + bq);
  if (tid < PS * RC) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < PWARPS; ++w) s += (pipe_sm + Sm<RC>::RED)[(w * PS + brow) * RC + bq];
    (pipe_sm + Sm<RC>::RH)[tid] = bvalid ? bpre - s : 0.0;
  }
  cp_commit_wait_all();
  __syncthreads();
  if (c.tr && tid == 0) c.tr[6] = gtime();
  strip_matvec<RC, true>(cw, xg, a, c);
  publish(c.flags + nd.strip0 + i, a, c);
}

// BELOW(s, rows [r0, r0+nrows)), backward (a14 slv_bwd_update + a15 bwd_update_upd):
// x_s -= L[rows, 0..n)^T xw[index[rows]].  The ancestors are complete when this runs.
// Passes of 64 rows; within a pass the work items are (strip k, row group g).  Nodes with at
// most FAT_NP strips give every warp at most one item, whose sums stay in registers across all
// passes (one RED per column and task); wider nodes flush after every pass.
template <int RC>
__device__ __forceinline__ void bwd_below(const PNode& nd, int r0, int nrows, const PipeArgs& a, const Ctx& c) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const i64 ld = nd.ld;
  const int np = nd.np;
  int rpi = (PS * np / PWARPS) & ~7;
  rpi = np == 3 ? 32 : max(8, min(PS, rpi));
  const int ng = (PS + rpi - 1) / rpi, nitems = np * ng;
  const bool keep = nitems <= PWARPS;   // np <= FAT_NP
  const double* Lr = a.arena + nd.off + (i64)r0 * ld + 2 * lane;
  double a0[RC], a1[RC];
#pragma unroll
  for (int q = 0; q < RC; ++q) a0[q] = a1[q] = 0.0;
  const int npass = (nrows + PS - 1) / PS;
  for (int p = 0; p < npass; ++p) {
    const int cnt = min(PS, nrows - p * PS);
    const int* idx = a.index + nd.idx_off + r0 + p * PS;
    __syncthreads();   // rh may still be read by the previous pass / task
    for (int k = tid; k < PS * RC; k += PT) {
      const int r = k / RC, q = k - r * RC;
      double v = 0.0;
      if (r < cnt && q < c.nr) {
        if (a.mode & M_NO_MAILBOX_BWD) {
          v = __ldcg(a.xw + (i64)idx[r] * a.nrhs + c.rc0 + q);   // the task waited for the ancestors' flags
        } else {   // poll the ancestor's published value itself
          const double* mp = a.xm + (i64)idx[r] * a.nrhs + c.rc0 + q;
          while (mailbox_empty(v = ld_mailbox(mp))) {
          }
        }
      }
      (pipe_sm + Sm<RC>::RH)[k] = v;
    }
    __syncthreads();
    if (c.tr && tid == 0 && p == 0) c.tr[1] = gtime();
    for (int it = warp; it < nitems; it += PWARPS) {
      const int k = it / ng, g = it - k * ng;
      const int ra = g * rpi, rb = min(cnt, ra + rpi);
      const int col = k * PS + 2 * lane;
      const bool colok = col < nd.ld;
      // the whole item (<= 64 rows x 64 columns) -> L2 first; the register loads below then pay the
      // HBM latency once per item instead of once per 8 rows
      for (int rbase = ra; rbase < rb; rbase += RPW)
        prefetch_tile(Lr - 2 * lane + (i64)(p * PS + rbase) * ld + k * PS, ld, rb - rbase, nd.ld - k * PS, lane);
      for (int rbase = ra; rbase < rb; rbase += RPW) {
        double2 t[RPW];
#pragma unroll
        for (int u = 0; u < RPW; ++u)
          t[u] = (rbase + u < rb && colok) ? ld_stream2(Lr + (i64)(p * PS + rbase + u) * ld + k * PS)
                                           : make_double2(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < RPW; ++u) {
          const int r = min(rbase + u, PS - 1);
#pragma unroll
          for (int q = 0; q < RC; ++q) {
            const double y = (pipe_sm + Sm<RC>::RH)[r * RC + q];
            a0[q] = fma(t[u].x, y, a0[q]);
            a1[q] = fma(t[u].y, y, a1[q]);
          }
        }
      }
      if (!keep || p == npass - 1) {
        double* dst = a.xw + (i64)(nd.sa + col) * a.nrhs + c.rc0;
#pragma unroll
        for (int q = 0; q < RC; ++q) {
          if (q < c.nr && col < nd.n) red_f64(dst + q, -a0[q]);
          if (q < c.nr && col + 1 < nd.n) red_f64(dst + a.nrhs + q, -a1[q]);
          a0[q] = a1[q] = 0.0;
        }
      }
    }
  }
}

}  // namespace


template <int RC, bool FWD>
__global__ void __launch_bounds__(PT, RC == 1 ? PIPE_OCC : 1) k_solve_pipe(const __grid_constant__ PipeArgs a) {
  __shared__ int s_next;
  __shared__ PTaskD s_td;
  static_assert(sizeof(PTaskD) == 96, "PTaskD is copied as 24 ints");
  Ctx c;
  const int tid = threadIdx.x;
  const int total = a.ntasks * a.nchunk;
  if (tid == 0) s_next = atomicAdd(a.sync, 1);
  __syncthreads();
  int t = s_next;
  while (t < total) {
    const int ti = t / a.nchunk, ch = t - ti * a.nchunk;
    int nxt = 0;
    if (tid == 0) nxt = atomicAdd(a.sync, 1);   // claimed early, consumed after this task
    if (tid < 24) ((int*)&s_td)[tid] = ((const int*)(a.tasks + ti))[tid];
    __syncthreads();   // descriptor visible; everybody has read s_next
    const PTask& tk = s_td.t;
    const PNode& nd = s_td.n;
    c.rc0 = ch * RC;
    c.nr = min(RC, a.nrhs - c.rc0);
    c.flags = a.sync + 32 + (i64)ch * a.stride;
    c.cnt = c.flags + (FWD ? 1 : 2) * a.nstrips;   // forward: per-strip counters; backward: per-node counters
    c.tr = a.trace ? a.trace + 8 * (i64)t : nullptr;
    if (c.tr && tid == 0) c.tr[0] = c.tr[1] = c.tr[2] = gtime();
    // rows below the diagonal block handled by this task (none for a DIAG task)
    const int rb0 = tk.kind == P_BELOW ? tk.r0 : nd.n;
    const int rb1 = tk.kind == P_BELOW ? tk.r0 + tk.nrows : (tk.kind == P_SMALL ? nd.m : nd.n);
    const bool fat = nd.np <= FAT_NP;
    if (rb1 > rb0 && !(a.mode & M_NO_PREFETCH))
      prefetch_l2(a.arena + nd.off + (i64)rb0 * nd.ld, nd.ld, rb1 - rb0, nd.n);
    if (FWD) {
      // a SMALL node's x is only read by its own rows below: no flag needed in the forward sweep
      if (tk.kind != P_BELOW) fwd_strip<RC>(nd, tk.node, tk.kind == P_DIAG ? tk.r0 : 0, tk.kind == P_DIAG, a, c);
      if (tk.kind != P_DIAG) {
        if (fat)
          fwd_below_fat<RC>(nd, rb0, rb1 - rb0, tk.kind == P_SMALL, a, c);
        else
          fwd_below<RC>(nd, rb0, rb1 - rb0, a, c);
        if (c.tr && tid == 0 && tk.kind == P_BELOW) c.tr[2] = gtime();
        bump_dests(a.dest + tk.dest_begin, tk.dest_count, c.cnt, a.mode);
      }
    } else {
      if (tk.kind != P_DIAG) {
        // every ancestor strip this task's rows map to has published its x
        // (with the mailbox the gather itself waits, value by value)
        if (a.mode & M_NO_MAILBOX_BWD)
          for (int k = tid; k < tk.dest_count; k += PT) wait_count(c.flags + a.dest[tk.dest_begin + k], 1, a.mode);
        bwd_below<RC>(nd, rb0, rb1 - rb0, a, c);
        if (c.tr && tid == 0 && tk.kind == P_BELOW) c.tr[2] = gtime();
        if (tk.kind == P_SMALL) fence_gpu(a.mode);   // own REDs, re-read by this CTA right below
        __syncthreads();
        if (tk.kind == P_BELOW && tid == 0) {
          fence_gpu(a.mode);
          red_s32(c.cnt + tk.node, 1);
        }
      }
      if (tk.kind != P_BELOW) bwd_strip<RC>(nd, tk.node, tk.kind == P_DIAG ? tk.r0 : 0, tk.kind == P_DIAG, a, c);
    }
    __syncthreads();
    if (c.tr && tid == 0) c.tr[3] = gtime();
    if (tid == 0) s_next = nxt;
    __syncthreads();
    t = s_next;
  }
}

// ------------------------------------------------------------------------------ diagonal inverses
// One CTA per strip: the 64 x 64 (or narrower) lower-triangular diagonal block L_ii of the factor
// is inverted by forward substitution, thread c computing column c of L_ii^-1.  Runs once per
// factorization; the solves then replace the latency-bound substitution (a13 slv_solve: dtrsv /
// dtrsm on the diagonal tile) by a 64 x 64 matrix-vector product.
constexpr int ISL = PS + 2;   // even leading dimension: pairs of a row are 16-byte aligned
__global__ void __launch_bounds__(PS) k_invert_diag(const PNode* __restrict__ nodes, const int* __restrict__ strip_node,
                                                    const double* __restrict__ arena, double* __restrict__ dinv) {
  __shared__ __align__(16) double Ls[PS * ISL];
  const int strip = blockIdx.x, c = threadIdx.x;
  const PNode nd = nodes[strip_node[strip]];
  const int i = strip - nd.strip0, r0 = i * PS, rw = min(PS, nd.n - r0);
  const double* D = arena + nd.off + (i64)r0 * nd.ld + r0;
  for (int idx = c; idx < PS * PS; idx += PS) {
    const int r = idx >> 6, cc = idx & (PS - 1);
    Ls[r * ISL + cc] = (r < rw && cc <= r) ? D[(i64)r * nd.ld + cc] : 0.0;
  }
  __syncthreads();
  {   // the diagonal is only ever divided by: keep its reciprocal (0 for rows beyond the strip)
    const double d = Ls[c * ISL + c];
    Ls[c * ISL + c] = d != 0.0 ? 1.0 / d : 0.0;
  }
  __syncthreads();
  // column c of the inverse stays in registers (the loops are fully unrolled, so y[] is indexed
  // statically): y_r = (delta_rc - sum_{k<r} L[r][k] y_k) / L[r][r]; rows above c come out as 0
  double y[PS];
#pragma unroll
  for (int r = 0; r < PS; ++r) {
    double s0 = (r == c) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k + 1 < r; k += 2) {
      const double2 l = *reinterpret_cast<const double2*>(Ls + r * ISL + k);
      s0 = fma(-l.x, y[k], s0);
      s1 = fma(-l.y, y[k + 1], s1);
    }
    if (r & 1) s0 = fma(-Ls[r * ISL + r - 1], y[r - 1], s0);
    y[r] = (s0 + s1) * Ls[r * ISL + r];
  }
  double* out = dinv + (i64)strip * (PS * PS);
#pragma unroll
  for (int r = 0; r < PS; ++r) out[r * PS + c] = y[r];
}

void launch_invert_diag(const PNode* nodes, const int* strip_node, int nstrips, const double* arena, double* dinv,
                        cudaStream_t st) {
  if (nstrips <= 0) return;
  k_invert_diag<<<nstrips, PS, 0, st>>>(nodes, strip_node, arena, dinv);
}

// multi-GPU: zero the entries of the work vector this rank does not own before it is summed
__global__ void k_mask_rows(double* __restrict__ xw, const char* __restrict__ keep, i64 n, int nrhs) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * nrhs && !keep[i / nrhs]) xw[i] = 0.0;
}
void launch_mask_rows(double* xw, const char* keep, int n, int nrhs, cudaStream_t st) {
  const i64 tot = (i64)n * nrhs;
  if (tot > 0) k_mask_rows<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(xw, keep, n, nrhs);
}

static int pipe_smem(int rc) { return (rc == 1 ? Sm<1>::TOTAL : Sm<PIPE_RC>::TOTAL) * (int)sizeof(double); }
static int g_pipe_grid[2][2] = {{0, 0}, {0, 0}};   // [rc index][fwd] resident CTAs on the whole device

template <int RC, bool FWD>
static int pipe_prepare() {
  CK(cudaFuncSetAttribute(k_solve_pipe<RC, FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, pipe_smem(RC)));
  int dev = 0, sms = 0, per = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_solve_pipe<RC, FWD>, PT, pipe_smem(RC)));
  if (per < 1) {
    fprintf(stderr, "spllt_b200: pipelined solve kernel does not fit on an SM\n");
    throw CudaFailure{cudaErrorLaunchOutOfResources, __FILE__, __LINE__};
  }
  return per * sms;
}

void pipe_init() {
  g_pipe_grid[0][1] = pipe_prepare<1, true>();
  g_pipe_grid[0][0] = pipe_prepare<1, false>();
  g_pipe_grid[1][1] = pipe_prepare<PIPE_RC, true>();
  g_pipe_grid[1][0] = pipe_prepare<PIPE_RC, false>();
}

int pipe_rc(int nrhs) { return nrhs == 1 ? 1 : PIPE_RC; }
int pipe_chunks(int nrhs) { return (nrhs + pipe_rc(nrhs) - 1) / pipe_rc(nrhs); }
i64 pipe_sync_ints(int nstrips, int nnodes, int nrhs) {
  return 32 + (i64)pipe_chunks(nrhs) * pipe_sync_stride(nstrips, nnodes);
}

void launch_solve_pipe(bool fwd, const PTaskD* tasks, int ntasks, const int* dest, const int* expect,
                       const double* arena, const double* dinv, const int* index, double* xw, double* xm, int nrhs,
                       int nstrips, int nnodes, int n, int* sync, cudaStream_t st, unsigned long long* trace,
                       bool keep_flags) {
  if (ntasks <= 0) return;
  const int mode = getenv("SPLLT_B200_PIPE_MODE") ? atoi(getenv("SPLLT_B200_PIPE_MODE")) : 0;
  PipeArgs a;
  a.tasks = tasks;
  a.dest = dest;
  a.expect = expect;
  a.arena = arena;
  a.dinv = dinv;
  a.index = index;
  a.xw = xw;
  a.xm = xm;
  a.sync = sync;
  a.ntasks = ntasks;
  a.nrhs = nrhs;
  a.nchunk = pipe_chunks(nrhs);
  a.stride = pipe_sync_stride(nstrips, nnodes);
  a.nstrips = nstrips;
  a.mode = mode;
  a.trace = trace;
  // keep_flags: only the claim counter is reset -- the flags raised by the previous launch on the
  // same region stay up (multi-GPU backward sweep: upper tree first, then this rank's subtrees)
  CK(cudaMemsetAsync(sync, 0, (keep_flags ? 32 : pipe_sync_ints(nstrips, nnodes, nrhs)) * sizeof(int), st));
  if (!keep_flags) CK(cudaMemsetAsync(xm, 0xff, (size_t)n * nrhs * sizeof(double), st));   // mailbox: nothing published
  const i64 total = (i64)ntasks * a.nchunk;
  const int rci = nrhs == 1 ? 0 : 1;
  const int grid = (int)std::min<i64>(total, g_pipe_grid[rci][fwd ? 1 : 0]);
  if (nrhs == 1) {
    if (fwd)
      k_solve_pipe<1, true><<<grid, PT, pipe_smem(1), st>>>(a);
    else
      k_solve_pipe<1, false><<<grid, PT, pipe_smem(1), st>>>(a);
  } else {
    if (fwd)
      k_solve_pipe<PIPE_RC, true><<<grid, PT, pipe_smem(PIPE_RC), st>>>(a);
    else
      k_solve_pipe<PIPE_RC, false><<<grid, PT, pipe_smem(PIPE_RC), st>>>(a);
  }
}

}  // namespace spllt
