// sm_100a kernels of the SpLLT numerical phase.
//
//   k_assemble   a10  spllt_init_node          src/spllt_kernels_mod.F90:2301-2364
//   k_panel      a1   spllt_factor_diag_block  :1168-1189   (inner panel diagonal block)
//                a2   spllt_solve_block        :1217-1229   (rows * L_pp^-T), fused
//   k_tile       a3   spllt_update_block       :1261-1292   (intra-node, src < 0)
//                a4-a7 spllt_update_between + expand_buffer / update_direct
//                     :2108-2237, :2010-2053, :14-93        (inter-node, fused scatter)
//   k_fwd_* / k_bwd_*  a13-a15 slv_solve / slv_fwd_update / slv_bwd_update
//                     src/spllt_solve_kernels_mod.F90:11-210
//
// FP64 tensor path: tcgen05 has no f64 kind, so the dense contractions run on
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) fed from multi-stage shared-memory tiles.
#include <cstdio>

#include "kernels.cuh"

#include "cuda_check.h"

namespace spllt {

#define FULL 0xffffffffu

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* s, const void* g, int bytes) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(s);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(g), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void* s, const void* g, int bytes) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(s);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(g), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------------------ assemble
__global__ void k_assemble(double* __restrict__ arena, const i64* __restrict__ dst, const i64* __restrict__ src,
                           const double* __restrict__ val, i64 cnt) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  i64 stride = (i64)gridDim.x * blockDim.x;
  for (; i < cnt; i += stride) arena[dst[i]] = val[src[i]];
}

// Extend-add of one entry: fire-and-forget FP64 reduction at the destination's L2.  `addr` is an
// absolute GLOBAL address (own arena or a peer's mapping); stating the state space keeps this one
// REDG instruction (an atomicAdd on a generic pointer expands to a shared / global dispatch with a
// returning ATOM).  sys: several GPUs may add into the same entry -> system scope.
template <bool SYS>
__device__ __forceinline__ void red_add_f64(i64 addr, double v) {
  if (SYS)
    asm volatile("red.relaxed.sys.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
  else
    asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra D_%=;\n"
      "bra W_%=;\n"
      "D_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// same, with back-off: used by warps that wait for a long time next to latency-critical warps
__device__ __forceinline__ void mbar_wait_backoff(unsigned long long* bar, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!done) __nanosleep(200);
  }
}
// ------------------------------------------------------------------------------ panel
// One inner panel: Cholesky of the pw x pw diagonal block (a1) + triangular solve of a chunk of
// rows against it (a2), pipelined inside the CTA.
//   warp 0    : factorizes the diagonal block, 16 columns at a time; each lane owns rows l and
//               l + 32 in registers.  Per 16-column block: left-looking update from the finished
//               columns, then column by column (column published through a double-buffered
//               shared array, __syncwarp only): a_rj -= (a_rk / a_kk) * a_jk, so only rsqrt sits
//               on the per-column critical path; l_rk = a_rk * rsqrt(a_kk) is produced off it.
//               After each block it arrives on that block's mbarrier.
//   warps 1-4 : thread r owns one row of the chunk; as soon as block cb of L is published they
//               run x_blk -= X_done * L^T (left-looking) and the 16 x 16 diagonal solve, i.e. the
//               solve hides behind the factorization except for its last block.
// Column c of L is contiguous in shared memory (Lt is L transposed).  Everything is written as
// 16-wide register blocks with run-time outer loops: a fully unrolled 64-column version is
// ~300 KB of straight-line SASS and runs instruction-fetch bound (measured 67 us per panel).
constexpr int PLD = IB + 1;   // X rows: conflict-free when thread r reads X[r][c]
constexpr int LTD = IB + 2;   // Lt rows: even, so (c*LTD + j) is 16-byte aligned for even j
constexpr int PB = 16;        // register block
constexpr int PANEL_THREADS = 128 + TRSM_ROWS;
constexpr int SMEM_PANEL = (IB * LTD + TRSM_ROWS * PLD + 2 * IB + IB + 8 + 2) * 8;

template <bool DBG>
__global__ void __launch_bounds__(PANEL_THREADS) k_panel(const PanelTask* __restrict__ tasks, double* __restrict__ arena,
                                                          int* __restrict__ info, int* __restrict__ pcount,
                                                          long long* __restrict__ dbg) {
  extern __shared__ __align__(16) double sm[];
  double* Lt = sm;                        // [IB][LTD]   Lt[c][r] = L[r][c]
  // the un-factorized block is staged in the same buffer, in the same (transposed) layout:
  // column k is consumed into registers before step k overwrites it with L's column k
  double* X = Lt + IB * LTD;              // [TRSM_ROWS][PLD]
  double* col = X + TRSM_ROWS * PLD;      // [2][IB]
  double* dinv = col + 2 * IB;            // [IB]
  unsigned long long* blk_done = reinterpret_cast<unsigned long long*>(dinv + IB);  // [4]
  int* store_flag = reinterpret_cast<int*>(blk_done + IB / PB);
  const PanelTask t = tasks[blockIdx.x];
  const int pw = t.pw, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double* gd = arena + t.d_off;
  double* gr = arena + t.r_off;
  if (DBG && tid == 0) dbg[blockIdx.x * 8 + 0] = clock64();
  if (tid == 0) {
    for (int i = 0; i < IB / PB; ++i) mbar_init(blk_done + i, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {  // diagonal block: lane -> column, PANEL_THREADS / 64 rows per pass, no integer division,
     // column test hoisted so that the loads of the unrolled passes are issued together
    static_assert(PANEL_THREADS % IB == 0 && TRSM_ROWS % IB == 0, "copy loops use groups of 64 threads");
    const int cc = tid & (IB - 1), r2 = tid >> 6;
    if (cc < pw) {
#pragma unroll 8
      for (int r = r2; r < pw; r += PANEL_THREADS / IB) Lt[cc * LTD + r] = (cc <= r) ? gd[(i64)r * t.ld + cc] : 0.0;
    }
  }
  // Every read of the un-factorized block by this CTA is complete (its values sit in shared
  // memory).  Each CTA of the panel draws a ticket; the one that draws the LAST ticket knows that
  // every other CTA of the panel has read the original block too, so it is the one that stores
  // L_pp in place.  No CTA ever waits for another one: no assumption on dispatch order.
  // The ticket is only needed when the factorization is finished: its round trip to L2 is hidden.
  __syncthreads();
  int ticket = t.ngroup - 1;
  if (tid == 0 && t.ngroup > 1) {
    __threadfence();
    ticket = atomicAdd(pcount + t.group, 1);
  }
  if (DBG && tid == 0) dbg[blockIdx.x * 8 + 1] = clock64();

  if (tid < 128) {
    // ---------------- factorization warps (4): thread (r, h) owns row r, columns h*8 .. h*8+7 of
    // the current 16-column block.  A lone warp issues one DFMA per ~4 cycles, so the block
    // updates are spread over the four SM sub-partitions; columns are exchanged through a
    // double-buffered shared array with one 128-thread named barrier per column.
    const int r = tid & 63, h = tid >> 6;
    for (int c0 = 0; c0 < pw; c0 += PB) {
      const int cbase = c0 + h * 8;
      double a[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = (r < pw && cbase + j < pw && cbase + j <= r) ? Lt[(cbase + j) * LTD + r] : 0.0;
      if (r >= c0 && r < pw) {
#pragma unroll 4
        for (int c = 0; c < c0; ++c) {
          double lrc = Lt[c * LTD + r];
          const double2* lc = reinterpret_cast<const double2*>(Lt + c * LTD + cbase);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            double2 l2 = lc[j];
            a[2 * j] -= lrc * l2.x;
            a[2 * j + 1] -= lrc * l2.y;
          }
        }
      }
#pragma unroll
      for (int kk = 0; kk < PB; ++kk) {
        const int k = c0 + kk;
        if (k < pw) {
          double* cb = col + (kk & 1) * IB;
          const int hh = kk >> 3;
          if (h == hh && r >= k) cb[r] = a[kk & 7];
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (h >= hh) {
            double akk = cb[k];
            double me = cb[r];
            double cj[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) cj[j] = cb[cbase + j];
            double s = rsqrt(akk);
            double tt = me * (s * s);
            // columns <= k of this thread (only when h == hh) receive garbage: they are final
            // and already published; entries above the diagonal are never stored
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] -= tt * cj[j];
            if (h == hh) {
              if (r >= k && r < pw) Lt[k * LTD + r] = me * s;
              if (r == k) {
                dinv[k] = s;
                if (!(akk > 0.0)) atomicMin(info, t.col0 + k + 1);   // every CTA of the panel sees the same pivot
              }
            }
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if ((tid & 31) == 0) mbar_arrive(blk_done + c0 / PB);
    }
    if (DBG && tid == 0) dbg[blockIdx.x * 8 + 2] = clock64();
    // store L_pp: only the CTA that drew the panel's last ticket (see above)
    if (tid == 0) *store_flag = ticket == t.ngroup - 1;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (*store_flag) {
      const int cc = tid & 63, r2 = tid >> 6;
      if (cc < pw) {
#pragma unroll 8
        for (int rr = r2; rr < pw; rr += 2)
          if (cc <= rr) gd[(i64)rr * t.ld + cc] = Lt[cc * LTD + rr];
      }
    }
    return;
  }

  // ---------------- solve warps
  const int row = tid - 128;
  {
    const int cc = row & (IB - 1), r2 = row >> 6;
    if (cc < pw) {
#pragma unroll 8
      for (int r = r2; r < t.nrows; r += TRSM_ROWS / IB) X[r * PLD + cc] = gr[(i64)r * t.ld + cc];
    }
  }
  asm volatile("bar.sync 2, %0;" ::"n"(TRSM_ROWS) : "memory");
  if (row < t.nrows) {
    double* xr = X + row * PLD;
    for (int c0 = 0; c0 < pw; c0 += PB) {
      double x[PB];
#pragma unroll
      for (int j = 0; j < PB; ++j) x[j] = (c0 + j < pw) ? xr[c0 + j] : 0.0;
      mbar_wait_backoff(blk_done + c0 / PB, 0);
#pragma unroll 4
      for (int c = 0; c < c0; ++c) {
        double xc = xr[c];
        const double2* lc = reinterpret_cast<const double2*>(Lt + c * LTD + c0);
#pragma unroll
        for (int j = 0; j < PB / 2; ++j) {
          double2 l2 = lc[j];
          x[2 * j] -= xc * l2.x;
          x[2 * j + 1] -= xc * l2.y;
        }
      }
#pragma unroll
      for (int kk = 0; kk < PB; ++kk) {
        if (c0 + kk < pw) {
          double xc = x[kk] * dinv[c0 + kk];
          x[kk] = xc;
          const double* lc = Lt + (c0 + kk) * LTD + c0;
#pragma unroll
          for (int jj = kk + 1; jj < PB; ++jj) x[jj] -= xc * lc[jj];
        }
      }
#pragma unroll
      for (int j = 0; j < PB; ++j)
        if (c0 + j < pw) xr[c0 + j] = x[j];
    }
  }
  asm volatile("bar.sync 2, %0;" ::"n"(TRSM_ROWS) : "memory");
  if (DBG && row == 0) dbg[blockIdx.x * 8 + 3] = clock64();
  {
    const int cc = row & (IB - 1), r2 = row >> 6;
    if (cc < pw) {
#pragma unroll 8
      for (int r = r2; r < t.nrows; r += TRSM_ROWS / IB) gr[(i64)r * t.ld + cc] = X[r * PLD + cc];
    }
  }
  if (DBG && row == 0) dbg[blockIdx.x * 8 + 4] = clock64();
}

// ------------------------------------------------------------------------------ tile update
constexpr int KC = 32;          // K chunk per pipeline stage
constexpr int SLD = KC + 4;     // padded shared row: (r*SLD + k) mod 16 distinct for r<4, k<4

template <int BM, int BN, int WM, int WN, int NSTAGE, bool SYS>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32)
    k_tile(const TileTask* __restrict__ tasks, double* __restrict__ arena, DevMaps mp) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  constexpr int FM = WM / 8, FN = WN / 8;
  extern __shared__ __align__(16) double sm[];
  double* As = sm;                          // [NSTAGE][BM][SLD]
  double* Bs = sm + NSTAGE * BM * SLD;      // [NSTAGE][BN][SLD]
  const TileTask t = tasks[blockIdx.x];
  const double* base = arena + t.off;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm0 = (warp / (BN / WN)) * WM, wn0 = (warp % (BN / WN)) * WN;
  const bool al = ((t.k0 & 1) == 0);
  const int nch = (t.kk + KC - 1) / KC;

  auto load_stage = [&](int ch, int st) {
    const int kb = ch * KC;
    for (int c = tid; c < (BM + BN) * (KC / 2); c += NT) {
      int r = c / (KC / 2), q = (c % (KC / 2)) * 2;
      double* s;
      int row;
      if (r < BM) {
        s = As + (st * BM + r) * SLD + q;
        row = t.i0 + min(r, t.mt - 1);
      } else {
        s = Bs + (st * BN + (r - BM)) * SLD + q;
        row = t.j0 + min(r - BM, t.nt - 1);
      }
      int k = kb + q;
      int valid = min(max(t.kk - k, 0), 2);
      const double* g = base + (i64)row * t.ld + t.k0 + k;
      if (al) {
        cp_async16(s, valid ? g : base, valid * 8);
      } else {
        cp_async8(s, valid > 0 ? g : base, valid > 0 ? 8 : 0);
        cp_async8(s + 1, valid > 1 ? g + 1 : base, valid > 1 ? 8 : 0);
      }
    }
  };

  double acc[FM][FN][2];
#pragma unroll
  for (int i = 0; i < FM; ++i)
#pragma unroll
    for (int j = 0; j < FN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // warps whose sub-tile lies strictly above the diagonal have nothing to contribute
  const bool active = (t.i0 + wm0 + WM - 1 >= t.j0 + wn0) && (wm0 < t.mt) && (wn0 < t.nt);

#pragma unroll
  for (int s = 0; s < NSTAGE - 1; ++s) {
    if (s < nch) load_stage(s, s);
    cp_commit();
  }
  for (int ch = 0; ch < nch; ++ch) {
    cp_wait<NSTAGE - 2>();
    __syncthreads();
    int nx = ch + NSTAGE - 1;
    if (nx < nch) load_stage(nx, nx % NSTAGE);
    cp_commit();
    if (active) {
      const double* a = As + ((ch % NSTAGE) * BM + wm0 + (lane >> 2)) * SLD + (lane & 3);
      const double* b = Bs + ((ch % NSTAGE) * BN + wn0 + (lane >> 2)) * SLD + (lane & 3);
#pragma unroll
      for (int k4 = 0; k4 < KC; k4 += 4) {
        double af[FM], bf[FN];
#pragma unroll
        for (int i = 0; i < FM; ++i) af[i] = a[i * 8 * SLD + k4];
#pragma unroll
        for (int j = 0; j < FN; ++j) bf[j] = b[j * 8 * SLD + k4];
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
          for (int j = 0; j < FN; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
    }
  }
  cp_wait<0>();
  if (!active) return;

  // Epilogue.  All loads of a batch are issued before the first store / atomic of that batch:
  // a naive `*c -= v` per element serialises 64 global round trips per thread (measured: 27 us
  // of a 36 us K = 64 tile).
  const bool scatter = t.src >= 0;
  const int li = lane >> 2, lj = 2 * (lane & 3);
  if (!scatter) {
    double* cbase = arena + t.off;
#pragma unroll
    for (int i = 0; i < FM; i += 2) {
      double cv[2][FN][2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        int ii = wm0 + (i + u) * 8 + li, gi = t.i0 + ii;
#pragma unroll
        for (int j = 0; j < FN; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            int jj = wn0 + j * 8 + lj + e, gj = t.j0 + jj;
            bool ok = ii < t.mt && jj < t.nt && gi >= gj;
            cv[u][j][e] = ok ? cbase[(i64)gi * t.ld + gj] : 0.0;
          }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        int ii = wm0 + (i + u) * 8 + li, gi = t.i0 + ii;
#pragma unroll
        for (int j = 0; j < FN; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            int jj = wn0 + j * 8 + lj + e, gj = t.j0 + jj;
            bool ok = ii < t.mt && jj < t.nt && gi >= gj;
            if (ok) cbase[(i64)gi * t.ld + gj] = cv[u][j][e] - acc[i + u][j][e];
          }
      }
    }
  } else {
    // per-thread destination columns are the same for every row block: fetch their maps once
    i64 qb[FN][2], qr[FN][2];
    int ql[FN][2];
#pragma unroll
    for (int j = 0; j < FN; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        int jj = wn0 + j * 8 + lj + e;
        i64 q = t.qoff + t.j0 + min(jj, t.nt - 1);
        qb[j][e] = mp.q_base[q];
        qr[j][e] = mp.q_rp[q];
        ql[j][e] = mp.q_ld[q];
      }
#pragma unroll
    for (int i = 0; i < FM; ++i) {
      int ii = wm0 + i * 8 + li, gi = t.i0 + min(ii, t.mt - 1);
      int rp[FN][2];
#pragma unroll
      for (int j = 0; j < FN; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          int jj = wn0 + j * 8 + lj + e, gj = t.j0 + jj;
          bool ok = ii < t.mt && jj < t.nt && gi >= gj;
          rp[j][e] = ok ? mp.rowpos[qr[j][e] + gi] : -1;
        }
#pragma unroll
      for (int j = 0; j < FN; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e)
          if (rp[j][e] >= 0) red_add_f64<SYS>(qb[j][e] + 8 * ((i64)rp[j][e] * ql[j][e]), -acc[i][j][e]);
    }
  }
}

// ------------------------------------------------------------------------------ persistent TMA tile kernel
// Warp-specialised, persistent version of the 128 x 128 tile update for launches that fill
// the machine.  One CTA per SM, three warpgroups (register file re-partitioned with
// setmaxnreg): the producer warp draws tile indices from a global counter (tiles are sorted by
// decreasing work: a dynamic LPT schedule) and streams the A / B row panels of each K chunk
// into a ring of shared-memory stages with 2-D tensor-map TMA loads (cp.async.bulk.tensor,
// SASS UTMALDG; one 16 k x 128 row box = 16 KB per instruction, 128-byte swizzle) that
// complete on the stage's mbarrier.  The 8 consumer warps wait on the stage, issue DMMA.8x8x4
// from it and release it; while they run the scatter / RMW epilogue of a tile the producer is
// already prefetching the next tile.  (1-D bulk copies of 256-byte rows were measured first:
// ~64 cycles per row copy, slower than cp.async.)
// The 128B swizzle XORs the 16-byte chunk index with (row % 8); fragment row rho of an 8-row
// group is read from shared row sigma(rho) = {0,2,4,6,1,3,5,7}, which makes the 8-byte fragment
// loads of a half-warp hit 16 distinct bank pairs.
constexpr int TM_BOXK = 16;                              // doubles per box row (128 bytes)
constexpr int TM_BOX = 128 * TM_BOXK;                    // doubles per A box (16 KB)
// stage = A lo, A hi (128 rows each) + B lo, B hi (BN rows each), KC = 32
__host__ __device__ constexpr int tm_stage(int bn) { return 2 * TM_BOX + 2 * bn * TM_BOXK; }
__host__ __device__ constexpr int tm_smem(int bn, int nst) { return nst * tm_stage(bn) * 8 + 1024 + 512; }
constexpr int SMEM_TILE_BG = 116 * 1024;   // > 227 KB / 2: one background CTA per SM
static_assert(KC == 2 * TM_BOXK, "stage = two boxes per operand");

// 2-D tiled TMA load: box (c0 .. c0+16, c1 .. c1+128) of the tensor described by `tmap`
__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, int c0, int c1, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

struct StageHdr {   // what a filled stage contains
  int tile;         // task index, -1 = no more work
  int chunk;        // K chunk index inside the tile
  int last;         // 1 if this is the tile's last chunk
  int valid;        // valid k in this chunk (<= KC)
};

// BN = 128: 8 consumer warps + producer warpgroup = 384 threads, 3 stages, one CTA per SM.
// BN = 64 : 4 consumer warps + producer warpgroup = 256 threads, 2 stages, TWO CTAs per SM: while
//           one CTA runs the (DMMA-idle) scatter epilogue of a tile the other keeps the pipe busy.
template <int BN, int TM_ST, bool SYS>
__global__ void __launch_bounds__(BN == 128 ? 384 : 256, BN == 128 ? 1 : 2)
    k_tile_tma(const TileTask* __restrict__ tasks, int ntasks, int* __restrict__ counter, double* __restrict__ arena,
               DevMaps mp, const unsigned char* __restrict__ tmaps, const unsigned char* __restrict__ tmaps_b) {
  constexpr int WM = 64, WN = 32, FM = WM / 8, FN = WN / 8;
  constexpr int NCW = (128 / WM) * (BN / WN);              // consumer warps
  constexpr int TM_STAGE = tm_stage(BN);
  constexpr int TM_BBOX = BN * TM_BOXK;
  extern __shared__ __align__(16) double sm_raw[];
  // the swizzle pattern is a function of the shared address: boxes must start 1024-byte aligned
  double* sm = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~uintptr_t(1023));
  unsigned long long* full = reinterpret_cast<unsigned long long*>(sm + TM_ST * TM_STAGE);
  unsigned long long* empty = full + TM_ST;
  StageHdr* hdr = reinterpret_cast<StageHdr*>(empty + TM_ST);
  TileTask* ttab = reinterpret_cast<TileTask*>(hdr + TM_ST);   // task of the tile a stage belongs to
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < TM_ST; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, NCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // 3 warpgroups: two of consumers, one for the producer (only its first warp works).  The
  // register file is re-partitioned between them: 168 regs/thread at launch (384 threads).
  if (warp >= NCW) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp != NCW) return;
    // ------------------------------------------------ producer
    int it = 0;  // stages filled so far
    bool first_tile = true;
    for (;;) {
      int ti = 0;
      if (counter) {
        if (lane == 0) ti = atomicAdd(counter, 1);
        ti = __shfl_sync(FULL, ti, 0);
      } else {   // non-persistent mode (background launches): one tile per CTA
        ti = first_tile ? (int)blockIdx.x : ntasks;
        first_tile = false;
      }
      if (ti >= ntasks) {
        int s = it % TM_ST;
        if (it >= TM_ST) mbar_wait(empty + s, ((it / TM_ST) - 1) & 1);
        if (lane == 0) {
          hdr[s] = StageHdr{-1, 0, 1, 0};
          mbar_arrive(full + s);
        }
        break;
      }
      const TileTask t = tasks[ti];
      const unsigned char* tm = tmaps + (size_t)t.node * 128;
      const unsigned char* tmb = tmaps_b + (size_t)t.node * 128;
      const int nch = (t.kk + KC - 1) / KC;
      for (int ch = 0; ch < nch; ++ch, ++it) {
        int s = it % TM_ST;
        if (it >= TM_ST) mbar_wait(empty + s, ((it / TM_ST) - 1) & 1);
        if (lane == 0) {
          hdr[s] = StageHdr{ti, ch, ch == nch - 1, min(KC, t.kk - ch * KC)};
          if (ch == 0) ttab[s] = t;
          mbar_expect_tx(full + s, TM_STAGE * 8);
          double* st = sm + s * TM_STAGE;
          int kc = t.k0 + ch * KC;
          tma_load_2d(st, tm, kc, t.i0, full + s);
          tma_load_2d(st + TM_BOX, tm, kc + TM_BOXK, t.i0, full + s);
          tma_load_2d(st + 2 * TM_BOX, tmb, kc, t.j0, full + s);
          tma_load_2d(st + 2 * TM_BOX + TM_BBOX, tmb, kc + TM_BOXK, t.j0, full + s);
        }
        __syncwarp();
      }
    }
    return;
  }

  // -------------------------------------------------- consumers
  if (BN == 128)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  else
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
  const int wm0 = (warp / (BN / WN)) * WM, wn0 = (warp % (BN / WN)) * WN;
  const int rho = lane >> 2, lk = lane & 3;
  // fragment row / column rho of an 8-group is read from shared row sigma(rho) (see header):
  // this lane's accumulator rows are 8 i + sg(rho), its two columns 8 j + sg(2 lk + e)
  auto sg = [](int x) { return ((x & 3) << 1) | (x >> 2); };
  const int sig = sg(rho), li = sig;
  double acc[FM][FN][2];
  TileTask t = {};
  bool active = false;
  int it = 0;
  for (;;) {
    int s = it % TM_ST;
    mbar_wait(full + s, (it / TM_ST) & 1);
    const StageHdr h = hdr[s];
    if (h.tile < 0) break;
    // the task descriptor is handed over by the producer through shared memory, once per tile
    if (h.chunk == 0) {
      t = ttab[s];
      active = (t.i0 + wm0 + WM - 1 >= t.j0 + wn0) && (wm0 < t.mt) && (wn0 < t.nt);
#pragma unroll
      for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < FN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    }
    if (active) {
      const double* a = sm + s * TM_STAGE + (wm0 + sig) * TM_BOXK + (lk & 1);
      const double* b = sm + s * TM_STAGE + 2 * TM_BOX + (wn0 + sig) * TM_BOXK + (lk & 1);
      const int hsel = lk >> 1;
      if (h.valid == KC) {
#pragma unroll
        for (int k4 = 0; k4 < KC; k4 += 4) {
          const int xs_ = ((((k4 & 15) >> 1) | hsel) ^ sig) << 1;
          const int x = (k4 >> 4) * TM_BOX + xs_, xb = (k4 >> 4) * TM_BBOX + xs_;
          double af[FM], bf[FN];
#pragma unroll
          for (int i = 0; i < FM; ++i) af[i] = a[i * 8 * TM_BOXK + x];
#pragma unroll
          for (int j = 0; j < FN; ++j) bf[j] = b[j * 8 * TM_BOXK + xb];
#pragma unroll
          for (int i = 0; i < FM; ++i)
#pragma unroll
            for (int j = 0; j < FN; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
      } else {
        for (int k4 = 0; k4 < h.valid; k4 += 4) {
          const bool kok = k4 + lk < h.valid;
          const int xs_ = ((((k4 & 15) >> 1) | hsel) ^ sig) << 1;
          const int x = (k4 >> 4) * TM_BOX + xs_, xb = (k4 >> 4) * TM_BBOX + xs_;
          double af[FM], bf[FN];
#pragma unroll
          for (int i = 0; i < FM; ++i) af[i] = kok ? a[i * 8 * TM_BOXK + x] : 0.0;
#pragma unroll
          for (int j = 0; j < FN; ++j) bf[j] = kok ? b[j * 8 * TM_BOXK + xb] : 0.0;
#pragma unroll
          for (int i = 0; i < FM; ++i)
#pragma unroll
            for (int j = 0; j < FN; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);
    ++it;
    if (!h.last || !active) continue;
    if (mp.rowpos == nullptr) continue;   // timing experiment only (SPLLT_B200_DEBUG_NOEPI): no epilogue

    // Lane lk holds columns {0,4,1,5}[lk] (e = 0) and {2,6,3,7}[lk] (e = 1) of each 8-column
    // group.  One exchange between lane pairs (xor 1) re-deals them to {0,2,1,3}[lk] / {4,6,5,7}[lk],
    // so that the four lanes of a row touch ONE 32-byte sector per store / RED instead of two.
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
      for (int j = 0; j < FN; ++j) {
        double send = (lk & 1) ? acc[i][j][0] : acc[i][j][1];
        double recv = __shfl_xor_sync(FULL, send, 1);
        if (lk & 1) acc[i][j][0] = recv;
        else acc[i][j][1] = recv;
      }
    const int cj0 = (lk & 1) ? sg(2 * (lk ^ 1) + 1) : sg(2 * lk);
    const int cj1 = (lk & 1) ? sg(2 * lk + 1) : sg(2 * (lk ^ 1));

    // epilogue (see k_tile): all loads of a batch before its stores / atomics
    if (t.src < 0) {
      double* cbase = arena + t.off;
#pragma unroll
      for (int i = 0; i < FM; i += 2) {
        double cv[2][FN][2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          int ii = wm0 + (i + u) * 8 + li, gi = t.i0 + ii;
#pragma unroll
          for (int j = 0; j < FN; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              int jj = wn0 + j * 8 + (e ? cj1 : cj0), gj = t.j0 + jj;
              bool ok = ii < t.mt && jj < t.nt && gi >= gj;
              cv[u][j][e] = ok ? cbase[(i64)gi * t.ld + gj] : 0.0;
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          int ii = wm0 + (i + u) * 8 + li, gi = t.i0 + ii;
#pragma unroll
          for (int j = 0; j < FN; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              int jj = wn0 + j * 8 + (e ? cj1 : cj0), gj = t.j0 + jj;
              bool ok = ii < t.mt && jj < t.nt && gi >= gj;
              if (ok) cbase[(i64)gi * t.ld + gj] = cv[u][j][e] - acc[i + u][j][e];
            }
        }
      }
    } else {
      i64 qb[FN][2], qr[FN][2];
      int ql[FN][2];
#pragma unroll
      for (int j = 0; j < FN; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          int jj = wn0 + j * 8 + (e ? cj1 : cj0);
          i64 q = t.qoff + t.j0 + min(jj, t.nt - 1);
          qb[j][e] = mp.q_base[q];
          qr[j][e] = mp.q_rp[q];
          ql[j][e] = mp.q_ld[q];
        }
      // Destination rows: rowpos depends on (ancestor segment of the column, source row).  When
      // all 8 columns of this thread land in the same ancestor (the common case) one index per
      // accumulator row suffices: the 8 loads are issued together (one L2 round trip instead of
      // one per row block), then the 64 atomics go out back to back.
      bool same = true;
#pragma unroll
      for (int j = 0; j < FN; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) same = same && (qr[j][e] == qr[0][0]);
      if (t.pad & 1) {
        // Exclusive launch: every tile of this launch comes from ONE source node, so no two
        // CTAs touch the same destination entry -> plain read-modify-write instead of RED.
        // (Measured: the RED epilogues of a 64^3 factorization sustain ~180 G FP64 atomics/s,
        // which is what the L2 atomic units deliver, and cost 28 % of the tile kernel's time.)
#pragma unroll
        for (int i = 0; i < FM; ++i) {
          int ii = wm0 + i * 8 + li, gi = t.i0 + min(ii, t.mt - 1);
          double* dst[FN][2];
          double cv[FN][2];
#pragma unroll
          for (int j = 0; j < FN; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              int jj = wn0 + j * 8 + (e ? cj1 : cj0), gj = t.j0 + jj;
              bool ok = ii < t.mt && jj < t.nt && gi >= gj;
              dst[j][e] = ok ? reinterpret_cast<double*>(qb[j][e]) + (i64)mp.rowpos[qr[j][e] + gi] * ql[j][e] : nullptr;
            }
#pragma unroll
          for (int j = 0; j < FN; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) cv[j][e] = dst[j][e] ? *dst[j][e] : 0.0;
#pragma unroll
          for (int j = 0; j < FN; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e)
              if (dst[j][e]) *dst[j][e] = cv[j][e] - acc[i][j][e];
        }
      } else if (__all_sync(FULL, same)) {
        int rp1[FM];
#pragma unroll
        for (int i = 0; i < FM; ++i) {
          int ii = wm0 + i * 8 + li, gi = t.i0 + min(ii, t.mt - 1);
          rp1[i] = mp.rowpos[qr[0][0] + gi];
        }
#pragma unroll
        for (int i = 0; i < FM; ++i) {
          int ii = wm0 + i * 8 + li, gi = t.i0 + ii;
#pragma unroll
          for (int j = 0; j < FN; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              int jj = wn0 + j * 8 + (e ? cj1 : cj0), gj = t.j0 + jj;
              if (ii < t.mt && jj < t.nt && gi >= gj)
                red_add_f64<SYS>(qb[j][e] + 8 * ((i64)rp1[i] * ql[j][e]), -acc[i][j][e]);
            }
        }
      } else {
#pragma unroll
        for (int i = 0; i < FM; ++i) {
          int ii = wm0 + i * 8 + li, gi = t.i0 + min(ii, t.mt - 1);
          int rp[FN][2];
#pragma unroll
          for (int j = 0; j < FN; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              int jj = wn0 + j * 8 + (e ? cj1 : cj0), gj = t.j0 + jj;
              bool ok = ii < t.mt && jj < t.nt && gi >= gj;
              rp[j][e] = ok ? mp.rowpos[qr[j][e] + gi] : -1;
            }
#pragma unroll
          for (int j = 0; j < FN; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e)
              if (rp[j][e] >= 0) red_add_f64<SYS>(qb[j][e] + 8 * ((i64)rp[j][e] * ql[j][e]), -acc[i][j][e]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------ peak probes
__global__ void __launch_bounds__(256) k_dmma_peak(int iters, double* sink) {
  double acc[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma884(acc[i][0], acc[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
  if (s == 123.456) sink[0] = s;
}
__global__ void __launch_bounds__(256) k_dfma_peak(int iters, double* sink) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 123.456) sink[0] = s;
}

// ------------------------------------------------------------------------------ solve
// Work vector xw is in pivot order, row-major n x nrhs (the nrhs values of one row are
// contiguous), so gathers through node%index move whole rows.
__global__ void k_permute_in(const double* __restrict__ x, int ldx, const int* __restrict__ porder,
                             double* __restrict__ xw, int n, int nrhs) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (i64)n * nrhs) return;
  int p = (int)(i / nrhs), r = (int)(i % nrhs);
  xw[i] = x[porder[p] + (i64)r * ldx];
}
__global__ void k_permute_out(double* __restrict__ x, int ldx, const int* __restrict__ porder,
                              const double* __restrict__ xw, int n, int nrhs) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (i64)n * nrhs) return;
  int p = (int)(i / nrhs), r = (int)(i % nrhs);
  x[porder[p] + (i64)r * ldx] = xw[i];
}

constexpr int SP = 64;         // panel width of the in-CTA triangular solves
constexpr int SPL = SP + 1;
constexpr int DIAG_THREADS = 512;

// Forward: solve L_cc x = b for one block column (w x w lower triangle, row-major), 64 columns
// at a time, left-looking: (a) b_p -= L[p rows, 0..p0) x[0..p0): 16 warps x 4 rows, every
// row's loads (coalesced 256-byte segments) in flight together, while the 64 x 64 diagonal
// block is fetched with cp.async; (b) triangular solve of the diagonal block -- one warp per
// right-hand side, each lane owns rows l and l + 32, pivot broadcast by shuffle, reciprocals
// precomputed so the per-column chain is multiply + shuffle + FMA.
template <int RC>
__global__ void __launch_bounds__(DIAG_THREADS) k_fwd_diag(const SolveBcol* __restrict__ bcs,
                                                           const double* __restrict__ arena, double* __restrict__ xw,
                                                           int nrhs) {
  extern __shared__ __align__(16) double sm[];
  const SolveBcol b = bcs[blockIdx.x];
  const int w = b.w, wp = w + 1, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NW = DIAG_THREADS / 32;
  const int rc0 = blockIdx.y * RC, nr = min(RC, nrhs - rc0);
  double* Ls = sm;                 // [SP][SPL]   (first: 8-byte cp.async destinations stay aligned)
  double* xs = sm + SP * SPL;      // [RC][wp]
  const double* L = arena + b.off + (i64)b.r0 * b.ld + b.r0;
  double* xg = xw + (i64)(b.sa + b.r0) * nrhs + rc0;
  for (int idx = tid; idx < w * RC; idx += DIAG_THREADS) {
    int k = idx / RC, q = idx - k * RC;
    xs[q * wp + k] = (q < nr) ? xg[(i64)k * nrhs + q] : 0.0;
  }
  __syncthreads();
  for (int p0 = 0; p0 < w; p0 += SP) {
    const int pw = min(SP, w - p0);
    {  // diagonal block -> Ls (asynchronous; consumed after the barrier below)
      const int cc = tid & (SP - 1), rr = tid >> 6;
      for (int r = rr; r < pw; r += DIAG_THREADS / SP) {
        bool ok = cc <= r && cc < pw;
        cp_async8(Ls + r * SPL + cc, ok ? L + (i64)(p0 + r) * b.ld + p0 + cc : L, ok ? 8 : 0);
      }
      cp_commit();
    }
    // (a) left-looking update of the panel's rows
    for (int rb = warp * 4; rb < pw; rb += 4 * NW) {
      double acc[4][RC];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int q = 0; q < RC; ++q) acc[u][q] = 0.0;
      for (int kb = 0; kb < p0; kb += 256) {
        double l[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double* lr = L + (i64)(p0 + min(rb + u, pw - 1)) * b.ld;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            int k = kb + j * 32 + lane;
            l[u][j] = (k < p0 && rb + u < pw) ? lr[k] : 0.0;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          int k = min(kb + j * 32 + lane, w - 1);
#pragma unroll
          for (int q = 0; q < RC; ++q) {
            double xv = xs[q * wp + k];
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u][q] += l[u][j] * xv;
          }
        }
      }
      if (p0 > 0) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int q = 0; q < RC; ++q) {
            double v = acc[u][q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
            if (lane == 0 && rb + u < pw && q < nr) xs[q * wp + p0 + rb + u] -= v;
          }
      }
    }
    cp_wait<0>();
    __syncthreads();
    // (b) triangular solve of the diagonal block
    for (int q = warp; q < nr; q += NW) {
      const int i0 = lane, i1 = lane + 32;
      double x0 = (i0 < pw) ? xs[q * wp + p0 + i0] : 0.0, x1 = (i1 < pw) ? xs[q * wp + p0 + i1] : 0.0;
      const double d0 = (i0 < pw) ? 1.0 / Ls[i0 * SPL + i0] : 0.0, d1 = (i1 < pw) ? 1.0 / Ls[i1 * SPL + i1] : 0.0;
      for (int k = 0; k < min(pw, 32); ++k) {
        double xk = __shfl_sync(FULL, x0 * d0, k);
        if (lane == k) x0 = xk;
        if (i0 > k) x0 -= Ls[i0 * SPL + k] * xk;
        if (i1 < pw) x1 -= Ls[i1 * SPL + k] * xk;
      }
      for (int k = 32; k < pw; ++k) {
        double xk = __shfl_sync(FULL, x1 * d1, k - 32);
        if (lane == k - 32) x1 = xk;
        if (i1 > k && i1 < pw) x1 -= Ls[i1 * SPL + k] * xk;
      }
      if (i0 < pw) xs[q * wp + p0 + i0] = x0;
      if (i1 < pw) xs[q * wp + p0 + i1] = x1;
    }
    __syncthreads();
  }
  for (int idx = tid; idx < w * nr; idx += DIAG_THREADS) {
    int k = idx / nr, q = idx - k * nr;
    xg[(i64)k * nrhs + q] = xs[q * wp + k];
  }
}

// Forward update: xw[index[r]] -= L[r, bcol] * x_bcol for a chunk of rows below the block column.
template <int RC>
__global__ void __launch_bounds__(256) k_fwd_upd(const SolveUpd* __restrict__ ups, const SolveBcol* __restrict__ bcs,
                                                 const double* __restrict__ arena, const int* __restrict__ index,
                                                 double* __restrict__ xw, int nrhs) {
  extern __shared__ double sm[];
  const SolveUpd u = ups[blockIdx.x];
  const SolveBcol b = bcs[u.bc];
  const int w = b.w, wp = w + 1, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rc0 = blockIdx.y * RC, nr = min(RC, nrhs - rc0);
  double* xs = sm;  // [RC][wp]
  const double* xg = xw + (i64)(b.sa + b.r0) * nrhs + rc0;
  for (int idx = tid; idx < w * RC; idx += 256) {
    int k = idx / RC, q = idx - k * RC;
    xs[q * wp + k] = (q < nr) ? xg[(i64)k * nrhs + q] : 0.0;
  }
  __syncthreads();
  int g = 32;  // lanes cooperating on one row
  while (g > 4 && (g >> 1) >= w) g >>= 1;
  const int rpw = 32 / g, sub = lane / g, lg = lane % g;
  const double* L = arena + b.off + (i64)u.r * b.ld + b.r0;
  const int* idx = index + b.idx_off + u.r;
  if (g == 32) {
    // wide block columns: two rows per warp in flight, 8 loads per row per lane issued together
    for (int base = warp * 2; base < u.nrows; base += 16) {
      const bool ok0 = base < u.nrows, ok1 = base + 1 < u.nrows;
      const double* lr0 = L + (i64)base * b.ld;
      const double* lr1 = L + (i64)min(base + 1, u.nrows - 1) * b.ld;
      double s0[RC], s1[RC];
#pragma unroll
      for (int q = 0; q < RC; ++q) s0[q] = s1[q] = 0.0;
      for (int kb = 0; kb < w; kb += 256) {
        double v0[8], v1[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          int k = kb + j * 32 + lane;
          v0[j] = (k < w && ok0) ? lr0[k] : 0.0;
          v1[j] = (k < w && ok1) ? lr1[k] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          int k = min(kb + j * 32 + lane, w - 1);
#pragma unroll
          for (int q = 0; q < RC; ++q) {
            double xv = xs[q * wp + k];
            s0[q] += v0[j] * xv;
            s1[q] += v1[j] * xv;
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q = 0; q < RC; ++q) {
          s0[q] += __shfl_xor_sync(FULL, s0[q], o);
          s1[q] += __shfl_xor_sync(FULL, s1[q], o);
        }
      }
      if (lane == 0) {
        if (ok0) {
          double* dst = xw + (i64)idx[base] * nrhs + rc0;
#pragma unroll
          for (int q = 0; q < RC; ++q)
            if (q < nr) atomicAdd(dst + q, -s0[q]);
        }
        if (ok1) {
          double* dst = xw + (i64)idx[base + 1] * nrhs + rc0;
#pragma unroll
          for (int q = 0; q < RC; ++q)
            if (q < nr) atomicAdd(dst + q, -s1[q]);
        }
      }
    }
    return;
  }
  for (int base = warp * rpw; base < u.nrows; base += 8 * rpw) {
    int row = base + sub;
    bool ok = row < u.nrows;
    double s[RC];
#pragma unroll
    for (int q = 0; q < RC; ++q) s[q] = 0.0;
    if (ok) {
      const double* lr = L + (i64)row * b.ld;
      for (int k = lg; k < w; k += g) {
        double l = lr[k];
#pragma unroll
        for (int q = 0; q < RC; ++q) s[q] += l * xs[q * wp + k];
      }
    }
    for (int o = g >> 1; o > 0; o >>= 1) {
#pragma unroll
      for (int q = 0; q < RC; ++q) s[q] += __shfl_xor_sync(FULL, s[q], o);
    }
    if (ok && lg == 0) {
      double* dst = xw + (i64)idx[row] * nrhs + rc0;
#pragma unroll
      for (int q = 0; q < RC; ++q)
        if (q < nr) atomicAdd(dst + q, -s[q]);
    }
  }
}

// Backward update: x_bcol -= L[rows, bcol]^T * xw[index[rows]].
template <int RC>
__global__ void __launch_bounds__(256) k_bwd_upd(const SolveUpd* __restrict__ ups, const SolveBcol* __restrict__ bcs,
                                                 const double* __restrict__ arena, const int* __restrict__ index,
                                                 double* __restrict__ xw, int nrhs) {
  __shared__ double ys[RC][SOLVE_ROWS];
  const SolveUpd u = ups[blockIdx.x];
  const SolveBcol b = bcs[u.bc];
  const int w = b.w, tid = threadIdx.x;
  const int rc0 = blockIdx.y * RC, nr = min(RC, nrhs - rc0);
  const int* idx = index + b.idx_off + u.r;
  for (int i = tid; i < u.nrows * RC; i += 256) {
    int row = i / RC, q = i - row * RC;
    ys[q][row] = (q < nr) ? xw[(i64)idx[row] * nrhs + rc0 + q] : 0.0;
  }
  __syncthreads();
  int tw = 256;  // threads across columns
  while (tw > 4 && (tw >> 1) >= w) tw >>= 1;
  const int tx = tid % tw, ty = tid / tw, ny = 256 / tw;
  const double* L = arena + b.off + (i64)u.r * b.ld + b.r0;
  double* xg = xw + (i64)(b.sa + b.r0) * nrhs + rc0;
  for (int k = tx; k < w; k += tw) {
    double s[RC];
#pragma unroll
    for (int q = 0; q < RC; ++q) s[q] = 0.0;
#pragma unroll 16
    for (int row = ty; row < u.nrows; row += ny) {
      double l = L[(i64)row * b.ld + k];
#pragma unroll
      for (int q = 0; q < RC; ++q) s[q] += l * ys[q][row];
    }
#pragma unroll
    for (int q = 0; q < RC; ++q)
      if (q < nr) atomicAdd(xg + (i64)k * nrhs + q, -s[q]);
  }
}

// ------------------------------------------------------------------------------ many right-hand sides: DMMA tiles
// With nrhs >= 16 the updates below a block column are GEMMs (the reference switches to dgemm there,
// src/spllt_solve_kernels_mod.F90:108-130, :185-202): L tile x X block on the FP64 tensor pipe
// (mma.sync.m8n8k4, SASS DMMA.8x8x4) instead of 8 right-hand sides per pass on scalar FMAs.
// Work vector: row-major n x nrhs, so a row of X / Y is one contiguous run of nrhs doubles.
constexpr int MQ = 64;          // right-hand sides per CTA pass
constexpr int MKC = 32;         // contraction chunk per pipeline stage
constexpr int MLD_A = MKC + 4;  // [64 rows][MKC]: (row * 36 + k) mod 16 distinct for a half-warp
constexpr int MLD_B = MQ + 4;   // [MKC][MQ]:      (k * 68 + q)  mod 16 distinct for a half-warp
constexpr int MST = 3;          // stages
constexpr int SMEM_FWD_MMA = MST * (64 * MLD_A + MKC * MLD_B) * 8;
constexpr int SMEM_BWD_MMA = MST * (2 * MKC * MLD_B) * 8;

// Forward: xw[index[r], :] -= L[r, bcol] * X_bcol for one 64-row chunk, MQ right-hand sides.
// 4 warps, warp tile 32 rows x 32 right-hand sides (16 DMMA per 8 fragment loads and 4-k step).
__global__ void __launch_bounds__(128) k_fwd_upd_mma(const SolveUpd* __restrict__ ups, const SolveBcol* __restrict__ bcs,
                                                     const double* __restrict__ arena, const int* __restrict__ index,
                                                     double* __restrict__ xw, int nrhs) {
  extern __shared__ __align__(16) double sm[];
  double* As = sm;                          // [MST][64][MLD_A]   L rows, k contiguous
  double* Bs = sm + MST * 64 * MLD_A;       // [MST][MKC][MLD_B]  X rows (k), right-hand sides contiguous
  const SolveUpd u = ups[blockIdx.x];
  const SolveBcol b = bcs[u.bc];
  const int w = b.w, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rc0 = blockIdx.y * MQ, nr = min(MQ, nrhs - rc0);
  const int wm0 = (warp >> 1) * 32, wn0 = (warp & 1) * 32;
  const double* L = arena + b.off + (i64)u.r * b.ld + b.r0;
  const double* X = xw + (i64)(b.sa + b.r0) * nrhs + rc0;
  const bool al = (b.r0 & 1) == 0;
  const int nch = (w + MKC - 1) / MKC;
  auto load_stage = [&](int ch, int st) {
    const int kb = ch * MKC;
    // A: 64 rows x MKC, pairs of doubles
    for (int c = tid; c < 64 * (MKC / 2); c += 128) {
      const int r = c / (MKC / 2), q = (c % (MKC / 2)) * 2;
      double* d = As + (st * 64 + r) * MLD_A + q;
      const int k = kb + q;
      const int valid = r < u.nrows ? min(max(w - k, 0), 2) : 0;
      const double* g = L + (i64)min(r, u.nrows - 1) * b.ld + k;
      if (al) {
        cp_async16(d, valid ? g : L, valid * 8);
      } else {
        cp_async8(d, valid > 0 ? g : L, valid > 0 ? 8 : 0);
        cp_async8(d + 1, valid > 1 ? g + 1 : L, valid > 1 ? 8 : 0);
      }
    }
    // B: MKC rows of X x MQ right-hand sides (nrhs is even: 16-byte pieces)
    for (int c = tid; c < MKC * (MQ / 2); c += 128) {
      const int k = c / (MQ / 2), q = (c % (MQ / 2)) * 2;
      double* d = Bs + (st * MKC + k) * MLD_B + q;
      const int valid = (kb + k < w) ? min(max(nr - q, 0), 2) : 0;
      cp_async16(d, valid ? X + (i64)(kb + k) * nrhs + q : X, valid * 8);
    }
  };
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
  for (int s = 0; s < MST - 1; ++s) {
    if (s < nch) load_stage(s, s);
    cp_commit();
  }
  for (int ch = 0; ch < nch; ++ch) {
    cp_wait<MST - 2>();
    __syncthreads();
    const int nx = ch + MST - 1;
    if (nx < nch) load_stage(nx, nx % MST);
    cp_commit();
    const double* a = As + ((ch % MST) * 64 + wm0 + (lane >> 2)) * MLD_A + (lane & 3);
    const double* bq = Bs + ((ch % MST) * MKC + (lane & 3)) * MLD_B + wn0 + (lane >> 2);
#pragma unroll
    for (int k4 = 0; k4 < MKC; k4 += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = a[i * 8 * MLD_A + k4];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = bq[k4 * MLD_B + j * 8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_wait<0>();
  // scatter: accumulator (row 8 i + lane / 4, right-hand sides 8 j + 2 (lane % 4) + {0, 1})
  const int* idx = index + b.idx_off + u.r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = wm0 + i * 8 + (lane >> 2);
    if (row >= u.nrows) continue;
    double* dst = xw + (i64)idx[row] * nrhs + rc0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int q = wn0 + j * 8 + 2 * (lane & 3) + e;
        if (q < nr) atomicAdd(dst + q, -acc[i][j][e]);
      }
  }
}

// Backward: X_bcol[k0 .. k0+64, :] -= L[rows, k0 .. k0+64]^T * xw[index[rows], :] for up to 512
// rows of a block column (the sums stay in registers over all of them: one reduction per entry
// and task).  The contraction runs over ROWS, so both operands are staged row by row.
__global__ void __launch_bounds__(128) k_bwd_upd_mma(const SolveUpdT* __restrict__ ups, const SolveBcol* __restrict__ bcs,
                                                     const double* __restrict__ arena, const int* __restrict__ index,
                                                     double* __restrict__ xw, int nrhs) {
  extern __shared__ __align__(16) double sm[];
  double* Ls = sm;                          // [MST][MKC rows][MLD_B]  L rows, columns k0.. contiguous
  double* Ys = sm + MST * MKC * MLD_B;      // [MST][MKC rows][MLD_B]  gathered rows of the work vector
  const SolveUpdT u = ups[blockIdx.x];
  const SolveBcol b = bcs[u.bc];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rc0 = blockIdx.y * MQ, nr = min(MQ, nrhs - rc0);
  const int kw = min(64, b.w - u.k0);       // valid columns of this tile
  const int wm0 = (warp >> 1) * 32, wn0 = (warp & 1) * 32;
  const double* L = arena + b.off + (i64)u.r * b.ld + b.r0 + u.k0;
  const int* idx = index + b.idx_off + u.r;
  const bool al = ((b.r0 + u.k0) & 1) == 0;
  const int nch = (u.nrows + MKC - 1) / MKC;
  auto load_stage = [&](int ch, int st) {
    const int rb = ch * MKC;
    for (int c = tid; c < MKC * 32; c += 128) {
      const int r = c >> 5, q = (c & 31) * 2;
      const bool rok = rb + r < u.nrows;
      const int row = min(rb + r, u.nrows - 1);
      {
        double* d = Ls + (st * MKC + r) * MLD_B + q;
        const int valid = rok ? min(max(kw - q, 0), 2) : 0;
        const double* g = L + (i64)row * b.ld + q;
        if (al) {
          cp_async16(d, valid ? g : L, valid * 8);
        } else {
          cp_async8(d, valid > 0 ? g : L, valid > 0 ? 8 : 0);
          cp_async8(d + 1, valid > 1 ? g + 1 : L, valid > 1 ? 8 : 0);
        }
      }
      {
        double* d = Ys + (st * MKC + r) * MLD_B + q;
        const int valid = rok ? min(max(nr - q, 0), 2) : 0;
        const double* g = xw + (i64)idx[row] * nrhs + rc0 + q;
        cp_async16(d, valid ? g : xw, valid * 8);
      }
    }
  };
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
  for (int s = 0; s < MST - 1; ++s) {
    if (s < nch) load_stage(s, s);
    cp_commit();
  }
  for (int ch = 0; ch < nch; ++ch) {
    cp_wait<MST - 2>();
    __syncthreads();
    const int nx = ch + MST - 1;
    if (nx < nch) load_stage(nx, nx % MST);
    cp_commit();
    // A fragment (m = column k, kk = row r) = L[r][k]; B fragment (kk = row r, n = q) = Y[r][q]
    const double* a = Ls + ((ch % MST) * MKC + (lane & 3)) * MLD_B + wm0 + (lane >> 2);
    const double* bq = Ys + ((ch % MST) * MKC + (lane & 3)) * MLD_B + wn0 + (lane >> 2);
#pragma unroll
    for (int r4 = 0; r4 < MKC; r4 += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = a[r4 * MLD_B + i * 8];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = bq[r4 * MLD_B + j * 8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_wait<0>();
  double* xg = xw + (i64)(b.sa + b.r0 + u.k0) * nrhs + rc0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = wm0 + i * 8 + (lane >> 2);
    if (k >= kw) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int q = wn0 + j * 8 + 2 * (lane & 3) + e;
        if (q < nr) atomicAdd(xg + (i64)k * nrhs + q, -acc[i][j][e]);
      }
  }
}

// Backward: solve L_cc^T x = b, panels from last to first, right-looking: (a) transposed
// triangular solve of the 64 x 64 diagonal block (fetched with cp.async during the previous
// panel's update); (b) x[0..p0) -= L[p rows, 0..p0)^T x_p: thread per column, coalesced across
// threads, 32 row loads in flight per thread.
template <int RC>
__global__ void __launch_bounds__(DIAG_THREADS) k_bwd_diag(const SolveBcol* __restrict__ bcs,
                                                           const double* __restrict__ arena, double* __restrict__ xw,
                                                           int nrhs) {
  extern __shared__ __align__(16) double sm[];
  const SolveBcol b = bcs[blockIdx.x];
  const int w = b.w, wp = w + 1, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NW = DIAG_THREADS / 32;
  const int rc0 = blockIdx.y * RC, nr = min(RC, nrhs - rc0);
  double* Ls = sm;                    // [SP][SPL]
  double* xs = sm + SP * SPL;         // [RC][wp]
  const double* L = arena + b.off + (i64)b.r0 * b.ld + b.r0;
  double* xg = xw + (i64)(b.sa + b.r0) * nrhs + rc0;
  for (int idx = tid; idx < w * RC; idx += DIAG_THREADS) {
    int k = idx / RC, q = idx - k * RC;
    xs[q * wp + k] = (q < nr) ? xg[(i64)k * nrhs + q] : 0.0;
  }
  const int npan = (w + SP - 1) / SP;
  for (int ip = npan - 1; ip >= 0; --ip) {
    const int p0 = ip * SP, pw = min(SP, w - p0);
    {
      const int cc = tid & (SP - 1), rr = tid >> 6;
      for (int r = rr; r < pw; r += DIAG_THREADS / SP) {
        bool ok = cc <= r && cc < pw;
        cp_async8(Ls + r * SPL + cc, ok ? L + (i64)(p0 + r) * b.ld + p0 + cc : L, ok ? 8 : 0);
      }
      cp_commit();
      cp_wait<0>();
    }
    __syncthreads();
    // (a) transposed solve: x_k = x_k / L_kk, then x_i -= L[k][i] x_k for i < k
    for (int q = warp; q < nr; q += NW) {
      const int i0 = lane, i1 = lane + 32;
      double x0 = (i0 < pw) ? xs[q * wp + p0 + i0] : 0.0, x1 = (i1 < pw) ? xs[q * wp + p0 + i1] : 0.0;
      const double d0 = (i0 < pw) ? 1.0 / Ls[i0 * SPL + i0] : 0.0, d1 = (i1 < pw) ? 1.0 / Ls[i1 * SPL + i1] : 0.0;
      for (int k = pw - 1; k >= 32; --k) {
        double xk = __shfl_sync(FULL, x1 * d1, k - 32);
        if (lane == k - 32) x1 = xk;
        if (i1 < k) x1 -= Ls[k * SPL + i1] * xk;
        x0 -= Ls[k * SPL + i0] * xk;
      }
      for (int k = min(pw, 32) - 1; k >= 0; --k) {
        double xk = __shfl_sync(FULL, x0 * d0, k);
        if (lane == k) x0 = xk;
        if (i0 < k) x0 -= Ls[k * SPL + i0] * xk;
      }
      if (i0 < pw) xs[q * wp + p0 + i0] = x0;
      if (i1 < pw) xs[q * wp + p0 + i1] = x1;
    }
    __syncthreads();
    // (b) columns left of the panel
    for (int k = tid; k < p0; k += DIAG_THREADS) {
      double acc[RC];
#pragma unroll
      for (int q = 0; q < RC; ++q) acc[q] = 0.0;
      const double* lc = L + (i64)p0 * b.ld + k;
      for (int r0 = 0; r0 < pw; r0 += 16) {
        double l[16];
        // unconditional (clamped) loads so that all 16 are issued back to back
#pragma unroll
        for (int r = 0; r < 16; ++r) l[r] = lc[(i64)min(r0 + r, pw - 1) * b.ld];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
          for (int q = 0; q < RC; ++q) {
            double xv = (r0 + r < pw) ? xs[q * wp + p0 + r0 + r] : 0.0;
            acc[q] += l[r] * xv;
          }
        }
      }
#pragma unroll
      for (int q = 0; q < RC; ++q)
        if (q < nr) xs[q * wp + k] -= acc[q];
    }
    __syncthreads();
  }
  for (int idx = tid; idx < w * nr; idx += DIAG_THREADS) {
    int k = idx / nr, q = idx - k * nr;
    xg[(i64)k * nrhs + q] = xs[q * wp + k];
  }
}

// ------------------------------------------------------------------------------ multi-GPU: peer memory
// One rank per GPU; every rank maps every other rank's arena and flag block (CUDA IPC), so a
// finished upper-tree block column is COPIED INTO THE PEERS' ARENAS by the owner's SMs over
// NVLink (st.global on mapped peer pointers) and announced by a flag -- no staging buffer, no
// pack / unpack, no library collective.  Flags carry the factorization's epoch (monotone), so
// nothing is ever reset and captured graphs replay unchanged.
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
constexpr unsigned long long PEER_TIMEOUT_NS = 120ull * 1000000000ull;   // a peer that never shows up: trap, do not hang

__global__ void k_epoch_inc(int* flags) { flags[F_EPOCH] += 1; }

// rows x cols sub-matrix at `off` (leading dimension ld) of this rank's arena -> the same place in
// every peer's arena, then flag `bc` := epoch on every peer (by the CTA that finishes last).
__global__ void __launch_bounds__(256) k_push_bcol(PeerSet ps, unsigned mask, i64 off, int ld, int rows, int cols, int bc,
                                                   int* __restrict__ done) {
  const double* src = ps.arena[ps.rank] + off;
  const bool vec = ((cols | ld) & 1) == 0 && ((off & 1) == 0);
  if (vec) {
    const int c2 = cols >> 1;
    const i64 n2 = (i64)rows * c2;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (i64)gridDim.x * blockDim.x) {
      const i64 r = i / c2;
      const int c = (int)(i - r * c2) * 2;
      const double2 v = *reinterpret_cast<const double2*>(src + r * ld + c);
      for (int p = 0; p < ps.world; ++p)
        if (mask >> p & 1u) *reinterpret_cast<double2*>(ps.arena[p] + off + r * ld + c) = v;
    }
  } else {
    const i64 n = (i64)rows * cols;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
      const i64 r = i / cols;
      const int c = (int)(i - r * cols);
      const double v = src[r * ld + c];
      for (int p = 0; p < ps.world; ++p)
        if (mask >> p & 1u) ps.arena[p][off + r * ld + c] = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int prev = atomicAdd(done, 1);
    if (prev == (int)gridDim.x - 1) {
      __threadfence_system();
      const int epoch = ps.flags[ps.rank][F_EPOCH];
      for (int p = 0; p < ps.world; ++p)
        if (mask >> p & 1u) st_release_sys(ps.flags[p] + F_BCOL + bc, epoch);
    }
  }
}

// the owner of block column `bc` has delivered it for the current factorization
__global__ void k_wait_bcol(const int* flags, int bc) {
  const int epoch = flags[F_EPOCH];
  const unsigned long long t0 = global_ns();
  while (ld_acquire_sys(flags + F_BCOL + bc) < epoch) {
    __nanosleep(200);
    if (global_ns() - t0 > PEER_TIMEOUT_NS) {
      printf("spllt_b200: timed out waiting for block column %d from its owner\n", bc);
      __trap();
    }
  }
}

// Barrier over the ranks (id = 1, 2, ... within a factorization).  what: 0 = arrive + wait,
// 1 = arrive only, 2 = wait only (the single-GPU emulation of several ranks splits it, since kernels
// of one GPU must never wait for kernels that are queued behind them).
__global__ void k_rank_barrier(PeerSet ps, int id, int what) {
  const int p = threadIdx.x;
  if (p >= ps.world) return;
  const int target = ps.flags[ps.rank][F_EPOCH] * 8 + id;
  if (what != 2) {
    __threadfence_system();
    st_release_sys(ps.flags[p] + F_BAR + ps.rank, target);
  }
  if (what != 1) {
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(ps.flags[ps.rank] + F_BAR + p) < target) {
      __nanosleep(200);
      if (global_ns() - t0 > PEER_TIMEOUT_NS) {
        printf("spllt_b200: rank %d timed out in barrier %d waiting for rank %d\n", ps.rank, id, p);
        __trap();
      }
    }
  }
}

// a9: spllt_subtree_apply_buffer / spllt_scatter_block (src/spllt_factorization_mod.F90:39-191,
// src/spllt_kernels_mod.F90:1122-1160).  G: b x b generated element of one subtree (lower triangle,
// row-major, already negated: the update epilogues accumulated -L L^T into it).  Entry (i, j) goes
// to column gq_base[j] (absolute address in the OWNER's arena, possibly a peer mapping), row
// rowpos[gq_rp[j] + i] -- one system-scope reduction per entry, 64 consecutive columns per warp pass.
__global__ void __launch_bounds__(256) k_apply_gen(const double* __restrict__ G, int b, const i64* __restrict__ gq_base,
                                                   const int* __restrict__ gq_ld, const i64* __restrict__ gq_rp,
                                                   const int* __restrict__ rowpos) {
  const int j = blockIdx.x * 64 + (threadIdx.x & 63);        // column of the element
  const int i0 = blockIdx.y * 64;                             // 64-row block
  if (i0 + 63 < blockIdx.x * 64 || j >= b) return;            // block above the diagonal / beyond the edge
  const i64 base = gq_base[j], rp = gq_rp[j];
  const int ld = gq_ld[j];
  for (int i = i0 + (threadIdx.x >> 6); i < min(i0 + 64, b); i += 4) {
    if (i < j) continue;
    const double v = G[(i64)i * b + j];
    if (v != 0.0) red_add_f64<true>(base + 8 * ((i64)rowpos[rp + i] * ld), v);
  }
}

// max |a - b| and max |b| over the lower trapezoids of a list of nodes held in two arenas with
// different layouts (multi-GPU factor against a single-GPU factor): out[0], out[1] as ordered bits
struct CmpNode {
  i64 off_a, off_b;
  int m, n, ld;
  int pad;
};
__global__ void k_compare_nodes(const CmpNode* __restrict__ nodes, const double* __restrict__ a,
                                const double* __restrict__ b, unsigned long long* out) {
  const CmpNode nd = nodes[blockIdx.x];
  double dmax = 0.0, bmax = 0.0;
  const i64 tot = (i64)nd.m * nd.n;
  for (i64 i = threadIdx.x; i < tot; i += blockDim.x) {
    const int r = (int)(i / nd.n), c = (int)(i - (i64)r * nd.n);
    if (c > r) continue;
    const double x = a[nd.off_a + (i64)r * nd.ld + c], y = b[nd.off_b + (i64)r * nd.ld + c];
    const double d = fabs(x - y);
    dmax = (d > dmax || d != d) ? d : dmax;
    bmax = fmax(bmax, fabs(y));
  }
  if (dmax != dmax) dmax = __longlong_as_double(0x7ff0000000000000LL);   // NaN -> +inf
  atomicMax(out, (unsigned long long)__double_as_longlong(dmax));
  atomicMax(out + 1, (unsigned long long)__double_as_longlong(bmax));
}

// ------------------------------------------------------------------------------ launchers
constexpr int SMEM_TILE_S = 2 * (64 + 64) * SLD * 8;
constexpr int SMEM_TILE_L = 3 * (128 + 128) * SLD * 8;
constexpr int SMEM_SOLVE_MAX = 200 * 1024;


void kernels_init() {
  CK(cudaFuncSetAttribute(k_panel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_PANEL));
  CK(cudaFuncSetAttribute(k_panel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_PANEL));
  CK(cudaFuncSetAttribute(k_tile<64, 64, 32, 32, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TILE_S));
  CK(cudaFuncSetAttribute(k_tile<128, 128, 64, 32, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TILE_L));
  CK(cudaFuncSetAttribute(k_tile_tma<128, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tm_smem(128, 3)));
  CK(cudaFuncSetAttribute(k_tile_tma<64, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TILE_BG));
  CK(cudaFuncSetAttribute(k_tile<64, 64, 32, 32, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TILE_S));
  CK(cudaFuncSetAttribute(k_tile<128, 128, 64, 32, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TILE_L));
  CK(cudaFuncSetAttribute(k_tile_tma<128, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tm_smem(128, 3)));
  CK(cudaFuncSetAttribute(k_tile_tma<64, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TILE_BG));
  CK(cudaFuncSetAttribute(k_fwd_diag<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_SOLVE_MAX));
  CK(cudaFuncSetAttribute(k_fwd_diag<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_SOLVE_MAX));
  CK(cudaFuncSetAttribute(k_bwd_diag<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_SOLVE_MAX));
  CK(cudaFuncSetAttribute(k_bwd_diag<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_SOLVE_MAX));
  CK(cudaFuncSetAttribute(k_fwd_upd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_SOLVE_MAX));
  CK(cudaFuncSetAttribute(k_fwd_upd<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_SOLVE_MAX));
  CK(cudaFuncSetAttribute(k_fwd_upd_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_FWD_MMA));
  CK(cudaFuncSetAttribute(k_bwd_upd_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BWD_MMA));
}

void launch_assemble(double* arena, const i64* dst, const i64* src, const double* val, i64 cnt, cudaStream_t st) {
  if (cnt <= 0) return;
  int blocks = (int)std::min<i64>((cnt + 255) / 256, 148 * 16);
  k_assemble<<<blocks, 256, 0, st>>>(arena, dst, src, val, cnt);
}
void launch_panel(const PanelTask* tasks, i64 count, double* arena, int* info, int* pcount, cudaStream_t st) {
  if (count > 0) k_panel<false><<<(unsigned)count, PANEL_THREADS, SMEM_PANEL, st>>>(tasks, arena, info, pcount, nullptr);
}
// diagnostic: runs the panel kernel with clock64() stamps at its phase boundaries (8 per CTA)
void launch_panel_dbg(const PanelTask* tasks, i64 count, double* arena, int* info, int* pcount, long long* dbg,
                      cudaStream_t st) {
  if (count > 0) k_panel<true><<<(unsigned)count, PANEL_THREADS, SMEM_PANEL, st>>>(tasks, arena, info, pcount, dbg);
}
void launch_tiles_tma(const TileTask* tasks, i64 count, int* counter, double* arena, DevMaps maps,
                      const void* tmaps, const void* tmaps_b, int bn, cudaStream_t st) {
  if (count <= 0) return;
  const unsigned char *tm = (const unsigned char*)tmaps, *tb = (const unsigned char*)tmaps_b;
  if (bn == 128) {
    unsigned grid = (unsigned)std::min<i64>(count, 148);
    if (maps.sys)
      k_tile_tma<128, 3, true><<<grid, 384, tm_smem(128, 3), st>>>(tasks, (int)count, counter, arena, maps, tm, tb);
    else
      k_tile_tma<128, 3, false><<<grid, 384, tm_smem(128, 3), st>>>(tasks, (int)count, counter, arena, maps, tm, tb);
  } else {
    unsigned grid = (unsigned)std::min<i64>(count, 296);
    if (maps.sys)
      k_tile_tma<64, 2, true><<<grid, 256, tm_smem(64, 2), st>>>(tasks, (int)count, counter, arena, maps, tm, tb);
    else
      k_tile_tma<64, 2, false><<<grid, 256, tm_smem(64, 2), st>>>(tasks, (int)count, counter, arena, maps, tm, tb);
  }
}
// Background variant (deferred inter-node updates on the low-priority stream): one tile per
// CTA (two CTAs per SM), so that SM slots are handed back every few tens of microseconds and
// the higher-priority panel / small-tile kernels of the main stream (<= 97 KB of shared memory
// each, i.e. one freed slot) get them first.
void launch_tiles_tma_bg(const TileTask* tasks, i64 count, double* arena, DevMaps maps, const void* tmaps,
                         const void* tmaps_b, cudaStream_t st) {
  if (count <= 0) return;
  const unsigned char *tm = (const unsigned char*)tmaps, *tb = (const unsigned char*)tmaps_b;
  if (maps.sys)
    k_tile_tma<64, 2, true><<<(unsigned)count, 256, tm_smem(64, 2), st>>>(tasks, (int)count, nullptr, arena, maps, tm, tb);
  else
    k_tile_tma<64, 2, false><<<(unsigned)count, 256, tm_smem(64, 2), st>>>(tasks, (int)count, nullptr, arena, maps, tm, tb);
}
void launch_tiles(const TileTask* tasks, i64 count, bool large, double* arena, DevMaps maps, cudaStream_t st) {
  if (count <= 0) return;
  if (large) {
    if (maps.sys)
      k_tile<128, 128, 64, 32, 3, true><<<(unsigned)count, 256, SMEM_TILE_L, st>>>(tasks, arena, maps);
    else
      k_tile<128, 128, 64, 32, 3, false><<<(unsigned)count, 256, SMEM_TILE_L, st>>>(tasks, arena, maps);
  } else {
    if (maps.sys)
      k_tile<64, 64, 32, 32, 2, true><<<(unsigned)count, 128, SMEM_TILE_S, st>>>(tasks, arena, maps);
    else
      k_tile<64, 64, 32, 32, 2, false><<<(unsigned)count, 128, SMEM_TILE_S, st>>>(tasks, arena, maps);
  }
}

void launch_epoch_inc(int* flags, cudaStream_t st) { k_epoch_inc<<<1, 1, 0, st>>>(flags); }
void launch_push_bcol(const PeerSet& ps, unsigned mask, i64 off, int ld, int rows, int cols, int bc, int* done,
                      cudaStream_t st) {
  const i64 n = (i64)rows * cols / 2;
  const unsigned grid = (unsigned)std::max<i64>(1, std::min<i64>((n + 255) / 256, 148 * 4));
  k_push_bcol<<<grid, 256, 0, st>>>(ps, mask, off, ld, rows, cols, bc, done);
}
void launch_wait_bcol(const int* flags, int bc, cudaStream_t st) { k_wait_bcol<<<1, 1, 0, st>>>(flags, bc); }
void launch_rank_barrier(const PeerSet& ps, int id, int what, cudaStream_t st) {
  k_rank_barrier<<<1, 32, 0, st>>>(ps, id, what);
}
void launch_apply_gen(const double* G, int b, const i64* gq_base, const int* gq_ld, const i64* gq_rp, const int* rowpos,
                      cudaStream_t st) {
  if (b <= 0) return;
  const unsigned nb64 = (unsigned)((b + 63) / 64);
  k_apply_gen<<<dim3(nb64, nb64), 256, 0, st>>>(G, b, gq_base, gq_ld, gq_rp, rowpos);
}
void launch_compare_nodes(const void* nodes, int count, const double* a, const double* b, unsigned long long* out,
                          cudaStream_t st) {
  if (count > 0) k_compare_nodes<<<count, 256, 0, st>>>((const CmpNode*)nodes, a, b, out);
}

void launch_permute_in(const double* x, int ldx, const int* porder, double* xw, int n, int nrhs, cudaStream_t st) {
  i64 tot = (i64)n * nrhs;
  if (tot > 0) k_permute_in<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(x, ldx, porder, xw, n, nrhs);
}
void launch_permute_out(double* x, int ldx, const int* porder, const double* xw, int n, int nrhs, cudaStream_t st) {
  i64 tot = (i64)n * nrhs;
  if (tot > 0) k_permute_out<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(x, ldx, porder, xw, n, nrhs);
}

static inline int solve_smem(int maxw, int rc, bool tri) {
  return (rc * (maxw + 1) + (tri ? SP * SPL : 0)) * 8;
}
static int g_maxw = 1024;
void set_solve_maxw(int w) { g_maxw = w; }

void launch_fwd_diag(const SolveBcol* bc, i64 count, const double* arena, double* xw, int nrhs, cudaStream_t st) {
  if (count <= 0) return;
  if (nrhs == 1)
    k_fwd_diag<1><<<dim3((unsigned)count, 1), DIAG_THREADS, solve_smem(g_maxw, 1, true), st>>>(bc, arena, xw, nrhs);
  else
    k_fwd_diag<8><<<dim3((unsigned)count, (nrhs + 7) / 8), DIAG_THREADS, solve_smem(g_maxw, 8, true), st>>>(bc, arena, xw, nrhs);
}
void launch_bwd_diag(const SolveBcol* bc, i64 count, const double* arena, double* xw, int nrhs, cudaStream_t st) {
  if (count <= 0) return;
  if (nrhs == 1)
    k_bwd_diag<1><<<dim3((unsigned)count, 1), DIAG_THREADS, solve_smem(g_maxw, 1, true), st>>>(bc, arena, xw, nrhs);
  else
    k_bwd_diag<8><<<dim3((unsigned)count, (nrhs + 7) / 8), DIAG_THREADS, solve_smem(g_maxw, 8, true), st>>>(bc, arena, xw, nrhs);
}
void launch_fwd_upd(const SolveUpd* up, i64 count, const SolveBcol* bc, const double* arena, const int* index,
                    double* xw, int nrhs, cudaStream_t st) {
  if (count <= 0) return;
  if (solve_use_mma(nrhs, false)) {
    k_fwd_upd_mma<<<dim3((unsigned)count, (nrhs + MQ - 1) / MQ), 128, SMEM_FWD_MMA, st>>>(up, bc, arena, index, xw, nrhs);
    return;
  }
  if (nrhs == 1)
    k_fwd_upd<1><<<dim3((unsigned)count, 1), 256, solve_smem(g_maxw, 1, false), st>>>(up, bc, arena, index, xw, nrhs);
  else
    k_fwd_upd<8><<<dim3((unsigned)count, (nrhs + 7) / 8), 256, solve_smem(g_maxw, 8, false), st>>>(up, bc, arena, index,
                                                                                                 xw, nrhs);
}
// Forward updates pay off from 16 right-hand sides; the backward tiles are 64 right-hand sides wide
// and only pay off when at least half of a tile is used (measured on Poisson 80^3: nrhs = 16:
// forward 6.0 -> 3.3 ms, backward 4.5 -> 5.2 ms; nrhs = 64: 20.4 -> 3.5 ms and 11.8 -> 5.2 ms).
bool solve_use_mma(int nrhs, bool backward) {
  static const bool off = getenv("SPLLT_B200_SOLVE_NO_MMA") != nullptr;
  return !off && nrhs >= (backward ? 32 : 16) && (nrhs & 1) == 0;
}
void launch_bwd_upd_mma(const SolveUpdT* up, i64 count, const SolveBcol* bc, const double* arena, const int* index,
                        double* xw, int nrhs, cudaStream_t st) {
  if (count <= 0) return;
  k_bwd_upd_mma<<<dim3((unsigned)count, (nrhs + MQ - 1) / MQ), 128, SMEM_BWD_MMA, st>>>(up, bc, arena, index, xw, nrhs);
}
void launch_bwd_upd(const SolveUpd* up, i64 count, const SolveBcol* bc, const double* arena, const int* index,
                    double* xw, int nrhs, cudaStream_t st) {
  if (count <= 0) return;
  if (nrhs == 1)
    k_bwd_upd<1><<<dim3((unsigned)count, 1), 256, 0, st>>>(up, bc, arena, index, xw, nrhs);
  else
    k_bwd_upd<8><<<dim3((unsigned)count, (nrhs + 7) / 8), 256, 0, st>>>(up, bc, arena, index, xw, nrhs);
}

static double* g_sink = nullptr;
// kind 0: DMMA, full occupancy (8 CTAs x 8 warps per SM); kind 1: DFMA, full occupancy;
// kind 10 + w: DMMA with exactly w warps per SM (one CTA per SM) -- how many resident warps
// the FP64 tensor pipe needs to stay busy.
double launch_dmma_peak(int iters, cudaStream_t st) {
  if (!g_sink) CK(cudaMalloc(&g_sink, 64));
  int blocks = 148 * 8;
  k_dmma_peak<<<blocks, 256, 0, st>>>(iters, g_sink);
  return (double)blocks * 8 * (double)iters * 16 * 512.0;
}
double launch_dmma_warps(int iters, int warps, cudaStream_t st) {
  if (!g_sink) CK(cudaMalloc(&g_sink, 64));
  // warps per SM = CTAs of 4 warps (one warp per SM sub-partition each)
  int per_sm = (warps + 3) / 4;
  k_dmma_peak<<<148 * per_sm, 128, 0, st>>>(iters, g_sink);
  return 148.0 * per_sm * 4 * (double)iters * 16 * 512.0;
}
double launch_dfma_peak(int iters, cudaStream_t st) {
  if (!g_sink) CK(cudaMalloc(&g_sink, 64));
  int blocks = 148 * 8;
  k_dfma_peak<<<blocks, 256, 0, st>>>(iters, g_sink);
  return (double)blocks * 256 * (double)iters * 16 * 2.0;
}

}  // namespace spllt
