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
//                 tiles L[strip i, strip j], j < i, as the x_j are published (flag per strip),
//                 accumulates in registers (no atomics), solves the 64 x 64 diagonal block,
//                 publishes x_i and raises flag (s,i).
//     BELOW(s,c)  64 rows below the diagonal block: same streaming product over all np strips,
//                 then xw[index[r]] -= sum (RED.ADD.F64) and one counter bump per ancestor node hit.
//     DIAG(s,0) waits until the node's counter shows that every contribution has arrived.
//   backward: the mirror image (gather only): BELOW chunks wait for the parent, gather
//     xw[index[r]], add L^T y into the node's x (RED) and bump the node's counter; DIAG(s,i)
//     streams L[strip j, strip i]^T x_j for j > i in decreasing j, then solves L_ii^T.
//   nodes with n <= 64 and few rows are ONE fused task (SMALL).
//
// CTAs claim tasks from a global counter in list order; lists are topological, so a claimed
// task only waits on tasks already claimed by running CTAs: no deadlock, no co-residency
// requirement.  Every L entry is read exactly once per sweep, 512 contiguous bytes per row and
// warp instruction, with the next tile's loads in flight while the warp polls for x_j.
#include <cstdio>
#include <cstdlib>

#include "kernels.cuh"

namespace spllt {

#define CK(x)                                                                                            \
  do {                                                                                                   \
    cudaError_t e_ = (x);                                                                                \
    if (e_ != cudaSuccess) {                                                                             \
      fprintf(stderr, "spllt_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      abort();                                                                                           \
    }                                                                                                    \
  } while (0)

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int PT = 256;            // threads per CTA
constexpr int PWARPS = PT / 32;    // 8
constexpr int RPW = PS / PWARPS;   // 8 rows of a strip per warp
constexpr int PSL = PS + 1;        // padded leading dimension of the diagonal block in shared memory

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

// Flags f[0], f[dir], f[2 dir], ... (at most `limit` of them matter).  Spins until the first one
// is raised and returns how many consecutive ones are (1..32): the caller skips polling for those.
__device__ __forceinline__ int wait_run(const int* f, int dir, int limit, int lane) {
  for (;;) {
    int v = lane < limit ? ld_acquire(f + dir * lane) : 0;
    unsigned b = __ballot_sync(FULL, v != 0);
    if (b & 1u) {
      __syncwarp();
      return b == FULL ? 32 : __ffs(~b) - 1;
    }
  }
}
__device__ __forceinline__ void wait_count(const int* c, int expect) {
  while (ld_acquire(c) < expect) {
  }
}

struct Ctx {
  const double* arena;
  const int* index;
  double* xw;
  int nrhs, rc0, nr;
  int* flags;
  int* cnt;
  double* Ls;   // [PS][PSL]
  double* red;  // forward [PS][RC]; backward [PWARPS][PS][RC]
  double* ys;   // backward below chunks: [PS][RC]
};

// diagonal block of strip i -> shared memory (lower triangle, rest zero), asynchronously
__device__ __forceinline__ void fetch_diag(const PNode& nd, int r0, int rw, const Ctx& c) {
  const double* D = c.arena + nd.off + (i64)r0 * nd.ld + r0;
  for (int idx = threadIdx.x; idx < rw * PS; idx += PT) {
    const int r = idx >> 6, cc = idx & (PS - 1);
    const bool ok = cc <= r;
    cp_async8z(c.Ls + r * PSL + cc, ok ? D + (i64)r * nd.ld + cc : D, ok ? 8 : 0);
  }
}

// x_i is final in global memory: make it visible, then raise the strip's flag
template <int RC>
__device__ __forceinline__ void publish(int* flag) {
  if (RC == 1)
    __syncwarp();   // only warp 0 wrote
  else
    __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    st_flag(flag, 1);
  }
}

// ------------------------------------------------------------------------------ forward
// DIAG(s, i), forward (a13 forward body + the intra-node part of a14).
template <int RC>
__device__ __forceinline__ void fwd_strip(const PNode& nd, int node, int i, const Ctx& c) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = i * PS, rw = min(PS, nd.n - r0);
  const i64 ld = nd.ld;
  fetch_diag(nd, r0, rw, c);
  double acc[RPW][RC];
#pragma unroll
  for (int u = 0; u < RPW; ++u)
#pragma unroll
    for (int q = 0; q < RC; ++q) acc[u][q] = 0.0;
  if (i > 0) {
    const double* Lr = c.arena + nd.off + (i64)(r0 + warp * RPW) * ld + 2 * lane;
    double2 t[RPW], tn[RPW];
#pragma unroll
    for (int u = 0; u < RPW; ++u)
      t[u] = (warp * RPW + u < rw) ? ld_stream2(Lr + (i64)u * ld) : make_double2(0.0, 0.0);
    int ready = 0;
    for (int j = 0; j < i; ++j) {
      if (j + 1 < i) {
#pragma unroll
        for (int u = 0; u < RPW; ++u)
          tn[u] = (warp * RPW + u < rw) ? ld_stream2(Lr + (i64)u * ld + (j + 1) * PS) : make_double2(0.0, 0.0);
      }
      if (j >= ready) ready = j + wait_run(c.flags + nd.strip0 + j, 1, i - j, lane);
      const double* xp = c.xw + (i64)(nd.sa + j * PS + 2 * lane) * c.nrhs + c.rc0;
      double x0[RC], x1[RC];
#pragma unroll
      for (int q = 0; q < RC; ++q) {
        x0[q] = q < c.nr ? __ldcg(xp + q) : 0.0;
        x1[q] = q < c.nr ? __ldcg(xp + c.nrhs + q) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < RPW; ++u)
#pragma unroll
        for (int q = 0; q < RC; ++q) acc[u][q] = fma(t[u].x, x0[q], fma(t[u].y, x1[q], acc[u][q]));
#pragma unroll
      for (int u = 0; u < RPW; ++u) t[u] = tn[u];
    }
#pragma unroll
    for (int u = 0; u < RPW; ++u)
#pragma unroll
      for (int q = 0; q < RC; ++q) {
        double v = acc[u][q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        acc[u][q] = v;
      }
  } else if (tid == 0 && nd.expect_f > 0) {
    wait_count(c.cnt + node, nd.expect_f);   // every contribution of the descendants has landed
  }
  if (lane == 0) {
#pragma unroll
    for (int u = 0; u < RPW; ++u)
#pragma unroll
      for (int q = 0; q < RC; ++q) c.red[(warp * RPW + u) * RC + q] = acc[u][q];
  }
  cp_commit_wait_all();
  __syncthreads();
  double* xg = c.xw + (i64)(nd.sa + r0) * c.nrhs + c.rc0;
  const double* Ls = c.Ls;
  for (int q = warp; q < c.nr; q += PWARPS) {
    const int i0 = lane, i1 = lane + 32;
    double x0 = i0 < rw ? __ldcg(xg + (i64)i0 * c.nrhs + q) - c.red[i0 * RC + q] : 0.0;
    double x1 = i1 < rw ? __ldcg(xg + (i64)i1 * c.nrhs + q) - c.red[i1 * RC + q] : 0.0;
    const double d0 = i0 < rw ? 1.0 / Ls[i0 * PSL + i0] : 0.0, d1 = i1 < rw ? 1.0 / Ls[i1 * PSL + i1] : 0.0;
    for (int k = 0; k < min(rw, 32); ++k) {
      const double xk = __shfl_sync(FULL, x0 * d0, k);
      if (lane == k) x0 = xk;
      if (i0 > k) x0 -= Ls[i0 * PSL + k] * xk;
      if (i1 < rw) x1 -= Ls[i1 * PSL + k] * xk;
    }
    for (int k = 32; k < rw; ++k) {
      const double xk = __shfl_sync(FULL, x1 * d1, k - 32);
      if (lane == k - 32) x1 = xk;
      if (i1 > k && i1 < rw) x1 -= Ls[i1 * PSL + k] * xk;
    }
    if (i0 < rw) __stcg(xg + (i64)i0 * c.nrhs + q, x0);
    if (i1 < rw) __stcg(xg + (i64)i1 * c.nrhs + q, x1);
  }
  publish<RC>(c.flags + nd.strip0 + i);
}

// BELOW(s, rows [r0, r0+nrows)), forward (a14 slv_fwd_update + a15 fwd_update_upd):
// xw[index[r]] -= L[r, 0..n) x_s, streaming over the node's strips as they are published.
template <int RC>
__device__ __forceinline__ void fwd_below(const PNode& nd, int r0, int nrows, const Ctx& c) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const i64 ld = nd.ld;
  const int np = nd.np;
  const double* Lr = c.arena + nd.off + (i64)(r0 + warp * RPW) * ld + 2 * lane;
  double acc[RPW][RC];
#pragma unroll
  for (int u = 0; u < RPW; ++u)
#pragma unroll
    for (int q = 0; q < RC; ++q) acc[u][q] = 0.0;
  double2 t[RPW], tn[RPW];
#pragma unroll
  for (int u = 0; u < RPW; ++u)
    t[u] = (warp * RPW + u < nrows && 2 * lane < nd.ld) ? ld_stream2(Lr + (i64)u * ld) : make_double2(0.0, 0.0);
  int ready = 0;
  for (int j = 0; j < np; ++j) {
    if (j + 1 < np) {
      const bool colok = (j + 1) * PS + 2 * lane < nd.ld;
#pragma unroll
      for (int u = 0; u < RPW; ++u)
        tn[u] = (warp * RPW + u < nrows && colok) ? ld_stream2(Lr + (i64)u * ld + (j + 1) * PS) : make_double2(0.0, 0.0);
    }
    if (j >= ready) ready = j + wait_run(c.flags + nd.strip0 + j, 1, np - j, lane);
    const int col = j * PS + 2 * lane;
    const double* xp = c.xw + (i64)(nd.sa + col) * c.nrhs + c.rc0;
    double x0[RC], x1[RC];
#pragma unroll
    for (int q = 0; q < RC; ++q) {
      x0[q] = (q < c.nr && col < nd.n) ? __ldcg(xp + q) : 0.0;
      x1[q] = (q < c.nr && col + 1 < nd.n) ? __ldcg(xp + c.nrhs + q) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < RPW; ++u)
#pragma unroll
      for (int q = 0; q < RC; ++q) acc[u][q] = fma(t[u].x, x0[q], fma(t[u].y, x1[q], acc[u][q]));
#pragma unroll
    for (int u = 0; u < RPW; ++u) t[u] = tn[u];
  }
#pragma unroll
  for (int u = 0; u < RPW; ++u)
#pragma unroll
    for (int q = 0; q < RC; ++q) {
      double v = acc[u][q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
      acc[u][q] = v;
    }
  const int* idx = c.index + nd.idx_off + r0 + warp * RPW;
#pragma unroll
  for (int u = 0; u < RPW; ++u) {
    if (lane == u && warp * RPW + u < nrows) {
      double* dst = c.xw + (i64)idx[u] * c.nrhs + c.rc0;
#pragma unroll
      for (int q = 0; q < RC; ++q)
        if (q < c.nr) atomicAdd(dst + q, -acc[u][q]);
    }
  }
}

// the contributions above are complete: bump the counter of every ancestor node they hit
__device__ __forceinline__ void bump_dests(const int* dest, int count, int* cnt) {
  __syncthreads();
  for (int k = threadIdx.x; k < count; k += PT) {
    __threadfence();
    atomicAdd(cnt + dest[k], 1);
  }
}

// ------------------------------------------------------------------------------ backward
// DIAG(s, i), backward: x_i = L_ii^-T (b_i - sum_{j>i} L[strip j, strip i]^T x_j)
// (a13 backward body + the intra-node part of slv_bwd_update).
template <int RC>
__device__ __forceinline__ void bwd_strip(const PNode& nd, int node, int i, bool wait_below, const Ctx& c) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = i * PS, cw = min(PS, nd.n - r0);
  const i64 ld = nd.ld;
  const int np = nd.np, nj = np - 1 - i;
  fetch_diag(nd, r0, cw, c);
  if (wait_below && i == np - 1 && tid == 0 && nd.expect_b > 0) wait_count(c.cnt + node, nd.expect_b);
  double a0[RC], a1[RC];
#pragma unroll
  for (int q = 0; q < RC; ++q) a0[q] = a1[q] = 0.0;
  if (nj > 0) {
    // tile rows: strip j of the node, this warp's 8 rows; columns: strip i (full, since i < np-1)
    const double* Lc = c.arena + nd.off + (i64)(warp * RPW) * ld + r0 + 2 * lane;
    double2 t[RPW], tn[RPW];
    {
      const int rb = (np - 1) * PS + warp * RPW;
#pragma unroll
      for (int u = 0; u < RPW; ++u)
        t[u] = (rb + u < nd.n) ? ld_stream2(Lc + (i64)((np - 1) * PS + u) * ld) : make_double2(0.0, 0.0);
    }
    int ready = 0;
    for (int jj = 0; jj < nj; ++jj) {
      const int j = np - 1 - jj;
      if (jj + 1 < nj) {   // strip j-1 is a full strip
#pragma unroll
        for (int u = 0; u < RPW; ++u) tn[u] = ld_stream2(Lc + (i64)((j - 1) * PS + u) * ld);
      }
      if (jj >= ready) ready = jj + wait_run(c.flags + nd.strip0 + j, -1, nj - jj, lane);
      const int rb = j * PS + warp * RPW;
      const double* xp = c.xw + (i64)(nd.sa + rb) * c.nrhs + c.rc0;
#pragma unroll
      for (int u = 0; u < RPW; ++u) {
        const bool ok = rb + u < nd.n;
#pragma unroll
        for (int q = 0; q < RC; ++q) {
          const double xv = (ok && q < c.nr) ? __ldcg(xp + (i64)u * c.nrhs + q) : 0.0;
          a0[q] = fma(t[u].x, xv, a0[q]);
          a1[q] = fma(t[u].y, xv, a1[q]);
        }
      }
#pragma unroll
      for (int u = 0; u < RPW; ++u) t[u] = tn[u];
    }
  }
#pragma unroll
  for (int q = 0; q < RC; ++q) {
    c.red[(warp * PS + 2 * lane) * RC + q] = a0[q];
    c.red[(warp * PS + 2 * lane + 1) * RC + q] = a1[q];
  }
  cp_commit_wait_all();
  __syncthreads();
  double* xg = c.xw + (i64)(nd.sa + r0) * c.nrhs + c.rc0;
  const double* Ls = c.Ls;
  for (int q = warp; q < c.nr; q += PWARPS) {
    const int i0 = lane, i1 = lane + 32;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int w = 0; w < PWARPS; ++w) {
      s0 += c.red[(w * PS + i0) * RC + q];
      s1 += c.red[(w * PS + i1) * RC + q];
    }
    double x0 = i0 < cw ? __ldcg(xg + (i64)i0 * c.nrhs + q) - s0 : 0.0;
    double x1 = i1 < cw ? __ldcg(xg + (i64)i1 * c.nrhs + q) - s1 : 0.0;
    const double d0 = i0 < cw ? 1.0 / Ls[i0 * PSL + i0] : 0.0, d1 = i1 < cw ? 1.0 / Ls[i1 * PSL + i1] : 0.0;
    // x_k = x_k / L_kk, then x_i -= L[k][i] x_k for i < k
    for (int k = cw - 1; k >= 32; --k) {
      const double xk = __shfl_sync(FULL, x1 * d1, k - 32);
      if (lane == k - 32) x1 = xk;
      if (i1 < k) x1 -= Ls[k * PSL + i1] * xk;
      x0 -= Ls[k * PSL + i0] * xk;
    }
    for (int k = min(cw, 32) - 1; k >= 0; --k) {
      const double xk = __shfl_sync(FULL, x0 * d0, k);
      if (lane == k) x0 = xk;
      if (i0 < k) x0 -= Ls[k * PSL + i0] * xk;
    }
    if (i0 < cw) __stcg(xg + (i64)i0 * c.nrhs + q, x0);
    if (i1 < cw) __stcg(xg + (i64)i1 * c.nrhs + q, x1);
  }
  publish<RC>(c.flags + nd.strip0 + i);
}

// BELOW(s, rows [r0, r0+nrows)), backward (a14 slv_bwd_update + a15 bwd_update_upd):
// x_s -= L[rows, 0..n)^T xw[index[rows]].  The ancestors are complete when this runs.
template <int RC>
__device__ __forceinline__ void bwd_below(const PNode& nd, int r0, int nrows, const Ctx& c) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const i64 ld = nd.ld;
  const int np = nd.np;
  const int* idx = c.index + nd.idx_off + r0;
  __syncthreads();   // ys may still be read by the previous chunk
  for (int k = tid; k < PS * RC; k += PT) {
    const int r = k / RC, q = k - r * RC;
    c.ys[k] = (r < nrows && q < c.nr) ? __ldcg(c.xw + (i64)idx[r] * c.nrhs + c.rc0 + q) : 0.0;
  }
  __syncthreads();
  // work items (strip k, row group g): enough of them to occupy the 8 warps of thin nodes,
  // as few as possible otherwise (one RED per column and item)
  int rpi = (PS * np / PWARPS) & ~7;
  rpi = max(8, min(PS, rpi));
  const int ng = (nrows + rpi - 1) / rpi, nitems = np * ng;
  const double* Lr = c.arena + nd.off + (i64)r0 * ld + 2 * lane;
  for (int it = warp; it < nitems; it += PWARPS) {
    const int k = it / ng, g = it - k * ng;
    const int ra = g * rpi, rb = min(nrows, ra + rpi);
    const int col = k * PS + 2 * lane;
    const bool colok = col < nd.ld;
    double a0[RC], a1[RC];
#pragma unroll
    for (int q = 0; q < RC; ++q) a0[q] = a1[q] = 0.0;
    for (int rbase = ra; rbase < rb; rbase += RPW) {
      double2 t[RPW];
#pragma unroll
      for (int u = 0; u < RPW; ++u)
        t[u] = (rbase + u < rb && colok) ? ld_stream2(Lr + (i64)(rbase + u) * ld + k * PS) : make_double2(0.0, 0.0);
#pragma unroll
      for (int u = 0; u < RPW; ++u) {
        const int r = min(rbase + u, PS - 1);
#pragma unroll
        for (int q = 0; q < RC; ++q) {
          const double y = c.ys[r * RC + q];
          a0[q] = fma(t[u].x, y, a0[q]);
          a1[q] = fma(t[u].y, y, a1[q]);
        }
      }
    }
    double* dst = c.xw + (i64)(nd.sa + col) * c.nrhs + c.rc0;
#pragma unroll
    for (int q = 0; q < RC; ++q) {
      if (q < c.nr && col < nd.n) atomicAdd(dst + q, -a0[q]);
      if (q < c.nr && col + 1 < nd.n) atomicAdd(dst + c.nrhs + q, -a1[q]);
    }
  }
}

}  // namespace

struct PipeArgs {
  const PTask* tasks;
  const PNode* nodes;
  const int* dest;
  const double* arena;
  const int* index;
  double* xw;
  int* sync;        // [0] claim counter; per chunk c: flags at 32 + c*stride, node counters after the flags
  int ntasks, nrhs, nchunk, stride, nstrips;
};

template <int RC, bool FWD>
__global__ void __launch_bounds__(PT, RC == 1 ? 2 : 1) k_solve_pipe(const PipeArgs a) {
  extern __shared__ __align__(16) double sm[];
  __shared__ int s_next;
  Ctx c;
  c.arena = a.arena;
  c.index = a.index;
  c.xw = a.xw;
  c.nrhs = a.nrhs;
  c.Ls = sm;
  c.red = sm + PS * PSL;
  c.ys = c.red + PWARPS * PS * RC;
  const int tid = threadIdx.x;
  const int total = a.ntasks * a.nchunk;
  if (tid == 0) s_next = atomicAdd(a.sync, 1);
  __syncthreads();
  int t = s_next;
  while (t < total) {
    __syncthreads();   // everybody has read s_next
    int nxt = 0;
    if (tid == 0) nxt = atomicAdd(a.sync, 1);   // claimed early, consumed after this task
    const int ti = t / a.nchunk, ch = t - ti * a.nchunk;
    const PTask tk = a.tasks[ti];
    const PNode nd = a.nodes[tk.node];
    c.rc0 = ch * RC;
    c.nr = min(RC, a.nrhs - c.rc0);
    c.flags = a.sync + 32 + (i64)ch * a.stride;
    c.cnt = c.flags + a.nstrips;
    // rows below the diagonal block handled by this task (none for a DIAG task)
    const int rb0 = tk.kind == P_BELOW ? tk.r0 : nd.n;
    const int rb1 = tk.kind == P_BELOW ? tk.r0 + tk.nrows : (tk.kind == P_SMALL ? nd.m : nd.n);
    if (FWD) {
      if (tk.kind != P_BELOW) fwd_strip<RC>(nd, tk.node, tk.kind == P_DIAG ? tk.r0 : 0, c);
      if (tk.kind != P_DIAG) {
        for (int r = rb0; r < rb1; r += PS) fwd_below<RC>(nd, r, min(PS, rb1 - r), c);
        bump_dests(a.dest + tk.dest_begin, tk.dest_count, c.cnt);
      }
    } else {
      if (tk.kind != P_DIAG) {
        if (tid == 0 && nd.pflag >= 0) wait_count(c.flags + nd.pflag, 1);   // parent (hence every ancestor) done
        for (int r = rb0; r < rb1; r += PS) bwd_below<RC>(nd, r, min(PS, rb1 - r), c);
        __threadfence();
        __syncthreads();
        if (tk.kind == P_BELOW && tid == 0) atomicAdd(c.cnt + tk.node, 1);
      }
      if (tk.kind != P_BELOW) bwd_strip<RC>(nd, tk.node, tk.kind == P_DIAG ? tk.r0 : 0, tk.kind == P_DIAG, c);
    }
    __syncthreads();
    if (tid == 0) s_next = nxt;
    __syncthreads();
    t = s_next;
  }
}

static int pipe_smem(int rc) { return (PS * PSL + PWARPS * PS * rc + PS * rc) * (int)sizeof(double); }
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
    abort();
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

void launch_solve_pipe(bool fwd, const PTask* tasks, int ntasks, const PNode* nodes, const int* dest,
                       const double* arena, const int* index, double* xw, int nrhs, int nstrips, int nnodes,
                       int* sync, cudaStream_t st) {
  if (ntasks <= 0) return;
  PipeArgs a;
  a.tasks = tasks;
  a.nodes = nodes;
  a.dest = dest;
  a.arena = arena;
  a.index = index;
  a.xw = xw;
  a.sync = sync;
  a.ntasks = ntasks;
  a.nrhs = nrhs;
  a.nchunk = pipe_chunks(nrhs);
  a.stride = pipe_sync_stride(nstrips, nnodes);
  a.nstrips = nstrips;
  CK(cudaMemsetAsync(sync, 0, pipe_sync_ints(nstrips, nnodes, nrhs) * sizeof(int), st));
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
