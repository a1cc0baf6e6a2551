// Launch wrappers of the sm_100a kernels (kernels.cu).  All pointers are device pointers.
#pragma once
#include <cuda_runtime.h>

#include "model.h"

namespace spllt {

struct DevMaps {          // inter-node update maps (device copies of Analysis::q_*)
  const i64* q_base;      // ABSOLUTE device address of the destination column (own or peer-mapped arena)
  const int* q_ld;
  const i64* q_rp;
  const int* rowpos;
  int sys;                // 1: destinations may be written by several GPUs (system-scope reductions)
};

void kernels_init();      // opt-in shared memory sizes; call once per process
void set_solve_maxw(int w);  // widest block column the solve kernels will see (shared memory size)
void launch_assemble(double* arena, const i64* dst, const i64* src, const double* val, i64 cnt, cudaStream_t st);
// pcount: one counter per panel (PanelTask::group), zero at the start of a factorization
void launch_panel(const PanelTask* tasks, i64 count, double* arena, int* info, int* pcount, cudaStream_t st);
void launch_panel_dbg(const PanelTask* tasks, i64 count, double* arena, int* info, int* pcount, long long* dbg,
                      cudaStream_t st);
void launch_tiles(const TileTask* tasks, i64 count, bool large, double* arena, DevMaps maps, cudaStream_t st);
void launch_tiles_tma_bg(const TileTask* tasks, i64 count, double* arena, DevMaps maps, const void* tmaps,
                         const void* tmaps_b, cudaStream_t st);
// persistent warp-specialised TMA variant of the 128 x 128 tiles; *counter must be 0 at launch
// bn = 128: 128 x 128 tiles, one CTA per SM; bn = 64: 128 x 64 tiles, two CTAs per SM.
// tmaps / tmaps_b: per-node tensor maps with 128-row / bn-row boxes.
void launch_tiles_tma(const TileTask* tasks, i64 count, int* counter, double* arena, DevMaps maps,
                      const void* tmaps, const void* tmaps_b, int bn, cudaStream_t st);

// ---- multi-GPU over peer-mapped memory (one rank per GPU)
constexpr int MAX_RANKS = 8;
// flag block of a rank (ints): [F_EPOCH] factorization counter (local), [F_BAR + r] barrier slot
// written by rank r, [F_BCOL + g] epoch of the last delivery of global block column g
constexpr int F_EPOCH = 0, F_BAR = 16, F_BCOL = 64;
struct PeerSet {
  int rank, world;
  double* arena[MAX_RANKS];   // arena[rank] = own arena, the others are IPC mappings of the peers'
  int* flags[MAX_RANKS];
};
void launch_epoch_inc(int* flags, cudaStream_t st);
// mask: destination ranks; done: this push's own completion counter (zero at launch)
void launch_push_bcol(const PeerSet& ps, unsigned mask, i64 off, int ld, int rows, int cols, int bc, int* done,
                      cudaStream_t st);
void launch_wait_bcol(const int* flags, int bc, cudaStream_t st);
void launch_rank_barrier(const PeerSet& ps, int id, int what, cudaStream_t st);
// generated element (b x b, lower triangle) of a subtree -> upper-tree block columns of any rank
void launch_apply_gen(const double* G, int b, const i64* gq_base, const int* gq_ld, const i64* gq_rp, const int* rowpos,
                      cudaStream_t st);
// nodes: device array of {i64 off_a, off_b; int m, n, ld, pad}; out[0] = max |a - b|, out[1] = max |b| (bits)
void launch_compare_nodes(const void* nodes, int count, const double* a, const double* b, unsigned long long* out,
                          cudaStream_t st);

// solve
void launch_permute_in(const double* x, int ldx, const int* porder, double* xw, int n, int nrhs, cudaStream_t st);
void launch_permute_out(double* x, int ldx, const int* porder, const double* xw, int n, int nrhs, cudaStream_t st);
void launch_fwd_diag(const SolveBcol* bc, i64 count, const double* arena, double* xw, int nrhs, cudaStream_t st);
void launch_fwd_upd(const SolveUpd* up, i64 count, const SolveBcol* bc, const double* arena, const int* index,
                    double* xw, int nrhs, cudaStream_t st);
void launch_bwd_upd(const SolveUpd* up, i64 count, const SolveBcol* bc, const double* arena, const int* index,
                    double* xw, int nrhs, cudaStream_t st);
void launch_bwd_diag(const SolveBcol* bc, i64 count, const double* arena, double* xw, int nrhs, cudaStream_t st);
// many right-hand sides (nrhs >= 16, even): the updates run on DMMA tiles.  launch_fwd_upd switches by
// itself (same 64-row chunks); the backward update has its own task list (up to 512 rows x 64 columns)
bool solve_use_mma(int nrhs, bool backward);
void launch_bwd_upd_mma(const SolveUpdT* up, i64 count, const SolveBcol* bc, const double* arena, const int* index,
                        double* xw, int nrhs, cudaStream_t st);

// pipelined solve (solve_pipe.cu): one persistent kernel per sweep
constexpr int PIPE_RC = 4;   // right-hand sides per pass when nrhs > 1
inline i64 pipe_sync_stride(int nstrips, int nnodes) { return (2 * (i64)nstrips + nnodes + 31) / 32 * 32; }
void pipe_init();            // opt-in shared memory + resident grid sizes; call once per process
int pipe_chunks(int nrhs);
i64 pipe_sync_ints(int nstrips, int nnodes, int nrhs);   // ints of flag / counter storage for one sweep
// sync is zeroed (stream-ordered) before the kernel starts
void launch_solve_pipe(bool fwd, const PTaskD* tasks, int ntasks, const int* dest, const int* expect,
                       const double* arena, const double* dinv, const int* index, double* xw, double* xm, int nrhs,
                       int nstrips, int nnodes, int n, int* sync, cudaStream_t st, unsigned long long* trace = nullptr,
                       bool keep_flags = false);
void launch_mask_rows(double* xw, const char* keep, int n, int nrhs, cudaStream_t st);
// inverses of the 64 x 64 diagonal blocks of every strip, dinv[strip][64][64] (once per factorization)
void launch_invert_diag(const PNode* nodes, const int* strip_node, int nstrips, const double* arena, double* dinv,
                        cudaStream_t st);

// FP64 tensor-pipe peak probe: every warp of every SM runs register-resident DMMA.
// Returns flops issued; time it with events.
double launch_dmma_peak(int iters, cudaStream_t st);
double launch_dfma_peak(int iters, cudaStream_t st);
double launch_dmma_warps(int iters, int warps, cudaStream_t st);

}  // namespace spllt
