// Engine: device-side state of one analysed matrix (see engine.cu).
#pragma once
#include <cuda_runtime.h>

#include <memory>
#include <utility>
#include <vector>

#include "kernels.cuh"
#include "model.h"

namespace spllt {

void require_gpu();  // aborts with a clear message when no CUDA device exists (no CPU fallback)

struct SolveGraphKey {
  double* x;
  int ldx, nrhs, job;
  cudaStream_t st;
  double* xw;
  bool operator==(const SolveGraphKey& o) const {
    return x == o.x && ldx == o.ldx && nrhs == o.nrhs && job == o.job && st == o.st && xw == o.xw;
  }
};

struct Engine {
  std::shared_ptr<Analysis> A;
  int device = 0;
  bool uploaded = false, factored = false, use_graph = true;
  cudaStream_t stream = nullptr, own = nullptr;
  bool own_stream = false;
  cudaStream_t side = nullptr;  // second stream: small-tile launches overlap the large-tile launch of their slot
  cudaStream_t bg = nullptr;    // low-priority stream: deferred inter-node updates
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  i64 lmap_count = 0;             // A -> L entries this rank assembles
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaStream_t ahead = nullptr;   // chain look-ahead: updates that run while the next panel is factorized
  cudaEvent_t ev_fork2 = nullptr, ev_ahead = nullptr;
  cudaStream_t comm = nullptr;    // deliveries of finished block columns to the peers (multi-GPU)
  cudaEvent_t ev_fork3 = nullptr, ev_comm = nullptr;
  bool overlap_tiles = true;

  double* arena = nullptr;
  i64* d_lmap_dst = nullptr;
  i64* d_lmap_src = nullptr;
  double* d_val = nullptr;
  PanelTask* d_panel = nullptr;
  TileTask* d_tile = nullptr;
  i64* d_qbase = nullptr;
  int* d_qld = nullptr;
  i64* d_qrp = nullptr;
  int* d_rowpos = nullptr;
  int* d_info = nullptr;
  int* d_counters = nullptr;  // one tile counter per launch (persistent TMA tile kernel)
  bool use_tma = true;
  void* d_tmaps = nullptr;    // CUtensorMap[nnodes] in device memory (128 bytes each), 128-row boxes
  void* d_tmaps_b = nullptr;  // same with tile_n-row boxes (B operand)
  SolveBcol* d_sb = nullptr;
  SolveUpd* d_su = nullptr;
  SolveUpdT* d_sut = nullptr;     // backward update tasks of the DMMA path (many right-hand sides)
  PNode* d_pnodes = nullptr;      // pipelined solve tables (solve_pipe.cu)
  PTaskD* d_ptask_f = nullptr;    // task + node records, forward / backward order
  PTaskD* d_ptask_b = nullptr;
  int* d_pdest = nullptr;
  int* d_strip_node = nullptr;
  int* d_pexpect = nullptr;
  PTaskD* d_ptask_ft = nullptr;   // multi-GPU: upper-tree lists
  PTaskD* d_ptask_bt = nullptr;
  int* d_pexpect_top = nullptr;
  char* d_col_keep = nullptr;
  double* d_dinv = nullptr;       // [nstrips][64][64] inverses of the diagonal blocks
  bool dinv_valid = false;
  int* d_psync = nullptr;         // flags + counters: forward region, then backward region
  i64 psync_ints = 0;             // ints per region
  double* d_xm = nullptr;         // mailbox copies of x (self-validating words): forward, then backward
  i64 xm_doubles = 0;
  int* d_index = nullptr;
  int* d_porder = nullptr;
  double* d_xw = nullptr;  // pivot-order work vector, n x nrhs row-major (persists between job 1 and job 2)
  double* d_x = nullptr;   // staging copy of the caller's host x
  int xw_nrhs = 0;

  // ---- multi-GPU (one rank per GPU): peer-mapped arenas / flag blocks, see kernels.cuh PeerSet
  PeerSet peers{};
  bool comm_ready = false;        // world == 1, or attach_peers() has run
  bool maps_ready = false;        // q_base (absolute addresses) uploaded
  int* d_flags = nullptr;         // this rank's flag block (F_EPOCH / F_BAR / F_BCOL)
  int* d_pushcnt = nullptr;       // [nbcol][MAX_RANKS] CTAs of a push that have finished (zeroed per factorization)
  double* d_gen = nullptr;        // generated elements of the subtrees this rank owns (Analysis::gen)
  i64* d_gqbase = nullptr;        // their destination maps (absolute addresses, owners' arenas)
  int* d_gqld = nullptr;
  i64* d_gqrp = nullptr;
  void apply_generated(cudaStream_t st);
  std::vector<void*> ipc_open;    // mappings to close on release
  struct StepRange { i64 begin, after_push, end; };
  std::vector<StepRange> step_ranges;   // launch index ranges of the upper-tree steps
  i64 phase0_end = 0;                   // launches [0, phase0_end) belong to phase 0
  std::vector<char> launch_sys;         // per launch: 1 = its extend-add needs system-scope reductions

  void export_handles(void* out128);                       // arena + flag IPC handles (2 x 64 bytes)
  void attach_peers(int rank, int world, const void* all_handles, void* const* same_process);
  void upload_maps();
  void factor_begin(const double* dval, cudaStream_t st);  // epoch, zero, assemble
  void factor_barrier(int id, int what, cudaStream_t st);
  void enqueue_range(i64 first, i64 last, cudaStream_t st);
  void factor_end(cudaStream_t st);                         // inverses of the diagonal blocks

  cudaGraphExec_t factor_graph = nullptr;
  const double* graph_val = nullptr;
  cudaStream_t graph_stream = nullptr;
  std::vector<std::pair<SolveGraphKey, cudaGraphExec_t>> solve_graphs;

  // host mirrors requested through spllt_set_mem_solve
  double* host_y = nullptr;
  int prep_nb = 0, prep_nrhs = 0;

  void upload_tables();
  void ensure_solve_buffers(int nrhs);
  void enqueue_factor(const double* dval, cudaStream_t st);
  void factor(const double* dval);
  void factor_host(const double* val);
  void profile_factor(const double* dval, double* ms5, const char* csv);
  void launch_one(const Launch& L, cudaStream_t st, bool background);
  void enqueue_solve(int nrhs, int job, cudaStream_t st);
  void ensure_dinv();
  bool use_pipe(int nrhs) const;   // persistent pipelined kernels (few right-hand sides) or level-set launches
  void solve(double* dx, int ldx, int nrhs, int job);
  void solve_phase(double* dx, int ldx, int nrhs, int phase);
  void profile_solve(double* dx, int ldx, int nrhs, double* ms6, const char* csv);
  void trace_solve(double* dx, int ldx, int nrhs, unsigned long long* out_f, unsigned long long* out_b);
  void solve_host(double* x, int nrhs, int job);
  void sync();
  int pivot_flag();
  void get_lcol(int g, double* out);
  void get_fwd(int nrhs, double* out);
  void release();
  ~Engine() { release(); }
};

}  // namespace spllt
