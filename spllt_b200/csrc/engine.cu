// Device-side state of one analysed matrix: HBM arena, uploaded work lists, captured CUDA
// graphs.  Replaces spllt_fkeep's lfact(:) / workspaces (src/spllt_data_mod.F90:330-388,
// src/spllt_factorization_mod.F90:347-472) and the STF / task-manager drivers
// (src/spllt_stf_mod.F90:18-192, src/spllt_solve_mod.F90:244-411).
#include "engine.h"

#include "cuda_check.h"

#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace spllt {


template <class T>
static T* upload(const std::vector<T>& v) {
  T* d = nullptr;
  size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
  CK(cudaMalloc(&d, bytes));
  if (!v.empty()) CK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

static unsigned long long g_kernels_ready = 0;   // bit d: kernel attributes set on device d

void require_gpu() {
  int cnt = 0;
  cudaError_t e = cudaGetDeviceCount(&cnt);
  if (e != cudaSuccess || cnt == 0) {
    fprintf(stderr,
            "spllt_b200: no CUDA device available -- the numerical phase has no CPU fallback "
            "(cudaGetDeviceCount: %s)\n",
            cudaGetErrorString(e));
    throw NoDevice{};
  }
}

void Engine::upload_tables() {
  if (uploaded) return;
  require_gpu();
  CK(cudaGetDevice(&device));
  if (!(g_kernels_ready >> (device & 63) & 1ull)) {   // function attributes are per device
    kernels_init();
    pipe_init();
    g_kernels_ready |= 1ull << (device & 63);
  }
  const Analysis& S = *A;
  if (!own_stream) {
    CK(cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking));
    own_stream = true;
  }
  if (!stream) stream = own;
  if (!side) {
    CK(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&bg, cudaStreamNonBlocking, lo));
    CK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&ahead, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&comm, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ev_fork3, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ev_comm, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ev_fork2, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ev_ahead, cudaEventDisableTiming));
  }
  overlap_tiles = !getenv("SPLLT_B200_NO_OVERLAP");
  CK(cudaMalloc(&arena, std::max<i64>(S.arena, 1) * sizeof(double)));
  CK(cudaMemset(arena, 0, std::max<i64>(S.arena, 1) * sizeof(double)));
  CK(cudaMalloc(&d_flags, (F_BCOL + std::max(S.nbcol, 1)) * sizeof(int)));
  CK(cudaMemset(d_flags, 0, (F_BCOL + std::max(S.nbcol, 1)) * sizeof(int)));
  CK(cudaMalloc(&d_pushcnt, MAX_RANKS * std::max(S.nbcol, 1) * sizeof(int)));
  if (S.gen_doubles > 0) CK(cudaMalloc(&d_gen, S.gen_doubles * sizeof(double)));   // generated elements of my subtrees
  if (S.world <= 1) {   // single GPU: the "peer set" is this arena
    peers = PeerSet{};
    peers.rank = 0;
    peers.world = 1;
    peers.arena[0] = arena;
    peers.flags[0] = d_flags;
    comm_ready = true;
  }
  // A -> L map with arena addresses.  Multi-GPU: a rank assembles the entries of the block columns
  // it owns (its subtrees, and its share of the upper tree).
  {
    std::vector<i64> dst, src;
    dst.reserve(S.nnz);
    src.reserve(S.nnz);
    for (int g = 0; g < S.nbcol; ++g) {
      if (S.world > 1 && S.bcol_owner[g] != S.rank) continue;
      const HNode& nd = S.nodes[S.bcol_node[g]];
      for (i64 e = S.lmap_ptr[g]; e < S.lmap_ptr[g + 1]; ++e) {
        dst.push_back(nd.off + (i64)S.lmap_row[e] * nd.ld + S.lmap_col[e]);
        src.push_back(S.lmap_src[e]);
      }
    }
    lmap_count = (i64)dst.size();
    d_lmap_dst = upload(dst);
    d_lmap_src = upload(src);
  }
  CK(cudaMalloc(&d_val, std::max<i64>(S.nnz, 1) * sizeof(double)));
  d_panel = upload(S.panel_tasks);
  d_tile = upload(S.tile_tasks);
  d_qld = upload(S.q_ld);
  d_qrp = upload(S.q_rp);
  d_rowpos = upload(S.rowpos);
  CK(cudaMalloc(&d_info, sizeof(int)));
  CK(cudaMalloc(&d_counters, (S.launches.size() + S.npanel_groups + 1) * sizeof(int)));
  use_tma = (S.nb % 2 == 0) && !getenv("SPLLT_B200_NO_TMA");
  if (use_tma) {
    // one 2-D tensor map per supernode: the node's m x ld row-major matrix, box = 16 k x 128 rows,
    // 128-byte swizzle.  The encoder lives in libcuda; fetch it through the runtime so that the
    // library still loads (for host-only analysis) on machines without a driver.
    typedef CUresult (*encode_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
      fprintf(stderr, "spllt_b200: cuTensorMapEncodeTiled not available\n");
      throw CudaFailure{cudaErrorNotSupported, __FILE__, __LINE__};
    }
    encode_t encode = (encode_t)fn;
    std::vector<CUtensorMap> maps(std::max(S.nnodes, 1));
    for (int pass = 0; pass < 2; ++pass) {   // pass 0: 128-row boxes (A operand), pass 1: tile_n-row boxes (B operand)
      for (int k = 0; k < S.nnodes; ++k) {
        const HNode& nd = S.nodes[k];
        cuuint64_t dims[2] = {(cuuint64_t)nd.ld, (cuuint64_t)nd.m};
        cuuint64_t strides[1] = {(cuuint64_t)nd.ld * 8};
        cuuint32_t box[2] = {16, (cuuint32_t)(pass == 0 ? 128 : S.tile_n)};
        cuuint32_t es[2] = {1, 1};
        CUresult r = encode(&maps[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, arena + nd.off, dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
          fprintf(stderr, "spllt_b200: cuTensorMapEncodeTiled failed (%d) for node %d (m=%d ld=%d)\n", (int)r, k, nd.m,
                  nd.ld);
          throw CudaFailure{cudaErrorInvalidValue, __FILE__, __LINE__};
        }
      }
      void** dst = pass == 0 ? &d_tmaps : &d_tmaps_b;
      CK(cudaMalloc(dst, maps.size() * sizeof(CUtensorMap)));
      CK(cudaMemcpy(*dst, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    }
  }
  d_sb = upload(S.sbcols);
  d_su = upload(S.supds);
  d_sut = upload(S.supds_t);
  d_pnodes = upload(S.pnodes);
  {
    std::vector<PTaskD> td;
    for (int dir = 0; dir < 2; ++dir) {
      const std::vector<PTask>& src = dir == 0 ? S.ptasks_f : S.ptasks_b;
      td.clear();
      td.reserve(src.size());
      for (const PTask& t : src) td.push_back(PTaskD{t, S.pnodes[t.node]});
      (dir == 0 ? d_ptask_f : d_ptask_b) = upload(td);
    }
    for (int dir = 0; dir < 2; ++dir) {   // multi-GPU: upper-tree lists
      const std::vector<PTask>& src = dir == 0 ? S.ptasks_ft : S.ptasks_bt;
      td.clear();
      for (const PTask& t : src) td.push_back(PTaskD{t, S.pnodes[t.node]});
      (dir == 0 ? d_ptask_ft : d_ptask_bt) = upload(td);
    }
    d_pexpect_top = upload(S.pexpect_top);
    d_col_keep = upload(S.col_keep);
  }
  d_pdest = upload(S.pipe_dest);
  d_strip_node = upload(S.strip_node);
  d_pexpect = upload(S.pexpect);
  CK(cudaMalloc(&d_dinv, std::max<i64>((i64)S.nstrips * PS * PS, 1) * sizeof(double)));
  d_index = upload(S.index);
  d_porder = upload(S.porder);
  int maxw = 1;
  for (const SolveBcol& b : S.sbcols) maxw = std::max(maxw, b.w);
  set_solve_maxw(maxw);
  // launch index ranges: phase 0, then the upper-tree steps (multi-GPU)
  phase0_end = 0;
  while (phase0_end < (i64)S.launches.size() && S.launches[phase0_end].phase == 0) ++phase0_end;
  step_ranges.assign(S.top_steps.size(), StepRange{0, 0, 0});
  {
    i64 i = phase0_end;
    for (size_t t = 0; t < S.top_steps.size(); ++t) {
      StepRange& r = step_ranges[t];
      r.begin = r.after_push = i;
      while (i < (i64)S.launches.size() && S.launches[i].depth == (int)t) {
        if (S.launches[i].kind == L_PUSH) r.after_push = i + 1;
        ++i;
      }
      r.end = i;
    }
  }
  // system-scope reductions: only the apply kernel of the generated elements adds into memory that
  // other GPUs add into at the same time; every tile launch scatters into this rank's own HBM
  launch_sys.assign(S.launches.size(), 0);
  uploaded = true;
  if (comm_ready) upload_maps();
}

// q_base as ABSOLUTE addresses: destination column of every below-diagonal row, in the arena of the
// rank that owns the destination block column (own arena, or a peer's mapping)
void Engine::upload_maps() {
  if (maps_ready) return;
  const Analysis& S = *A;
  std::vector<i64> qa(S.q_base.size());
  const int nb = S.nb;
  for (int k = 0; k < S.nnodes; ++k) {
    const HNode& nd = S.nodes[k];
    const int* idx = S.index.data() + nd.idx_off;
    for (int r = nd.n; r < nd.m; ++r) {
      const i64 g = nd.row_base + (r - nd.n);
      int owner = 0;
      if (S.world > 1) {
        const int a = S.col2node[idx[r]];
        owner = S.bcol_owner[S.nodes[a].bcol0 + (idx[r] - S.nodes[a].sa) / nb];
      }
      if (S.q_base[g] < 0)   // into the generated element of the node's subtree (local)
        qa[g] = (i64)(uintptr_t)(d_gen + (-S.q_base[g] - 1));
      else
        qa[g] = (i64)(uintptr_t)(peers.arena[owner] + S.q_base[g]);
    }
  }
  if (d_qbase) cudaFree(d_qbase);
  d_qbase = upload(qa);
  // generated elements -> upper tree: absolute destination columns in their owners' arenas
  if (!S.gen.empty()) {
    std::vector<i64> ga(S.gq_base.size());
    for (size_t k = 0; k < ga.size(); ++k)
      ga[k] = (i64)(uintptr_t)(peers.arena[S.bcol_owner[S.gq_bcol[k]]] + S.gq_base[k]);
    if (d_gqbase) cudaFree(d_gqbase);
    d_gqbase = upload(ga);
    if (!d_gqld) d_gqld = upload(S.gq_ld);
    if (!d_gqrp) d_gqrp = upload(S.gq_rp);
  }
  maps_ready = true;
}

// a9 (spllt_subtree_apply_buffer): the generated elements of this rank's subtrees are added into the
// upper-tree block columns, wherever they live
void Engine::apply_generated(cudaStream_t st) {
  const Analysis& S = *A;
  for (const GenElem& g : S.gen)
    launch_apply_gen(d_gen + g.off, g.b, d_gqbase + g.map0, d_gqld + g.map0, d_gqrp + g.map0, d_rowpos, st);
}

// ---- multi-GPU bootstrap.  export_handles: CUDA IPC handles of this rank's arena and flag block;
// the caller gathers them from all ranks (torch.distributed / MPI all-gather of 128 bytes per rank)
// and passes the table to attach_peers.  same_process (optional): raw device pointers instead --
// {arena, flags} per rank -- for several ranks emulated inside one process on one GPU.
void Engine::export_handles(void* out128) {
  upload_tables();
  cudaIpcMemHandle_t h[2];
  CK(cudaIpcGetMemHandle(&h[0], arena));
  CK(cudaIpcGetMemHandle(&h[1], d_flags));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(out128, h, 128);
}

void Engine::attach_peers(int rank, int world, const void* all_handles, void* const* same_process) {
  upload_tables();
  if (world > MAX_RANKS) {
    fprintf(stderr, "spllt_b200: at most %d ranks\n", MAX_RANKS);
    throw CudaFailure{cudaErrorInvalidValue, __FILE__, __LINE__};
  }
  peers = PeerSet{};
  peers.rank = rank;
  peers.world = world;
  for (int p = 0; p < world; ++p) {
    if (p == rank) {
      peers.arena[p] = arena;
      peers.flags[p] = d_flags;
    } else if (same_process) {
      peers.arena[p] = (double*)same_process[2 * p];
      peers.flags[p] = (int*)same_process[2 * p + 1];
    } else {
      cudaIpcMemHandle_t h[2];
      memcpy(h, (const char*)all_handles + 128 * (size_t)p, 128);
      void *pa = nullptr, *pf = nullptr;
      CK(cudaIpcOpenMemHandle(&pa, h[0], cudaIpcMemLazyEnablePeerAccess));
      CK(cudaIpcOpenMemHandle(&pf, h[1], cudaIpcMemLazyEnablePeerAccess));
      ipc_open.push_back(pa);
      ipc_open.push_back(pf);
      peers.arena[p] = (double*)pa;
      peers.flags[p] = (int*)pf;
    }
  }
  comm_ready = true;
  maps_ready = false;
  upload_maps();
}

void Engine::ensure_solve_buffers(int nrhs) {
  if (nrhs <= xw_nrhs) return;
  // the captured solve graphs point into the buffers that are about to be replaced
  for (auto& e : solve_graphs) cudaGraphExecDestroy(e.second);
  solve_graphs.clear();
  if (d_xw) CK(cudaFree(d_xw));
  if (d_x) CK(cudaFree(d_x));
  CK(cudaMalloc(&d_xw, std::max<i64>((i64)A->n * nrhs, 1) * sizeof(double)));
  CK(cudaMalloc(&d_x, std::max<i64>((i64)A->n * nrhs, 1) * sizeof(double)));
  CK(cudaMemset(d_xw, 0, std::max<i64>((i64)A->n * nrhs, 1) * sizeof(double)));
  if (d_xm) CK(cudaFree(d_xm));
  xm_doubles = std::max<i64>((i64)A->n * nrhs, 1);
  CK(cudaMalloc(&d_xm, 2 * xm_doubles * sizeof(double)));   // mailboxes of the forward / backward sweep
  if (d_psync) CK(cudaFree(d_psync));
  psync_ints = pipe_sync_ints(A->nstrips, A->nnodes, nrhs);
  CK(cudaMalloc(&d_psync, 2 * psync_ints * sizeof(int)));
  xw_nrhs = nrhs;
}

void Engine::launch_one(const Launch& L, cudaStream_t st, bool background) {
  // system-scope reductions only where another GPU may add into the same entries at the same time:
  // phase-0 launches that scatter into the upper tree (peer-owned or shared with peers' scatters)
  const size_t li = (size_t)(&L - A->launches.data());
  DevMaps mp{d_qbase, d_qld, d_qrp, d_rowpos, li < launch_sys.size() ? (int)launch_sys[li] : 0};
  switch (L.kind) {
    case L_PANEL: launch_panel(d_panel + L.begin, L.count, arena, d_info, d_counters + A->launches.size(), st); break;
    case L_PUSH: {
      const HNode& nd = A->nodes[A->bcol_node[L.begin]];
      const int r0 = A->bcol_c[L.begin] * A->nb;
      launch_push_bcol(peers, (unsigned)L.deadline, nd.off + (i64)r0 * nd.ld + r0, nd.ld, nd.m - r0,
                       std::min(A->nb, nd.n - r0), (int)L.begin, d_pushcnt + MAX_RANKS * L.begin + L.count, st);
      break;
    }
    case L_WAIT: launch_wait_bcol(d_flags, (int)L.begin, st); break;
    case L_TILE_S: launch_tiles(d_tile + L.begin, L.count, false, arena, mp, st); break;
    case L_TILE_L:
      if (use_tma && background && A->tile_n == 64)
        launch_tiles_tma_bg(d_tile + L.begin, L.count, arena, mp, d_tmaps, d_tmaps_b, st);
      else if (use_tma) {  // d_counters[i] belongs to launch i and is zeroed at the start of every factorization
        if (getenv("SPLLT_B200_DEBUG_NOEPI")) mp.rowpos = nullptr;
        launch_tiles_tma(d_tile + L.begin, L.count, d_counters + (&L - A->launches.data()), arena, mp, d_tmaps,
                         d_tmaps_b, A->tile_n, st);
      } else
        launch_tiles(d_tile + L.begin, L.count, true, arena, mp, st);
      break;
  }
}

// One factorization = prologue (epoch, zero L, scatter A), the launches of phase 0 (everything on
// one GPU; this rank's subtrees on several), and -- multi-GPU -- the upper-tree steps between two
// barriers over the ranks: after the first every rank has zeroed the block columns it accumulates,
// so peers may scatter into them; after the second every contribution of a subtree has landed.
void Engine::factor_begin(const double* dval, cudaStream_t st) {
  const Analysis& S = *A;
  launch_epoch_inc(d_flags, st);
  if (S.world > 1) {
    if (S.own_end > S.own_begin) CK(cudaMemsetAsync(arena + S.own_begin, 0, (S.own_end - S.own_begin) * sizeof(double), st));
    if (S.arena > S.top_begin) CK(cudaMemsetAsync(arena + S.top_begin, 0, (S.arena - S.top_begin) * sizeof(double), st));
    CK(cudaMemsetAsync(d_pushcnt, 0, MAX_RANKS * std::max(S.nbcol, 1) * sizeof(int), st));
    if (S.gen_doubles > 0) CK(cudaMemsetAsync(d_gen, 0, S.gen_doubles * sizeof(double), st));
  } else {
    CK(cudaMemsetAsync(arena, 0, S.arena * sizeof(double), st));
  }
  CK(cudaMemsetAsync(d_info, 0x7f, sizeof(int), st));
  CK(cudaMemsetAsync(d_counters, 0, (S.launches.size() + S.npanel_groups + 1) * sizeof(int), st));
  launch_assemble(arena, d_lmap_dst, d_lmap_src, dval, lmap_count, st);
}

void Engine::factor_barrier(int id, int what, cudaStream_t st) {
  if (A->world > 1) launch_rank_barrier(peers, id, what, st);
}

void Engine::factor_end(cudaStream_t st) {
  launch_invert_diag(d_pnodes, d_strip_node, A->nstrips, arena, d_dinv, st);
}

// Launches [first, last) of the schedule.
void Engine::enqueue_range(i64 first, i64 last, cudaStream_t st) {
  const Analysis& S = *A;
  // Streams (all of it is captured into one CUDA graph):
  //   st   : panel launch of every slot + the updates on the critical path;
  //   side : the small-tile launch of a slot, forked so that it overlaps the large-tile one
  //          (disjoint destinations or atomics);
  //   bg   : low priority; deferred inter-node updates (Launch::stream == 1).  A background
  //          launch starts after the panel of its slot and is joined right before the panel
  //          of its deadline slot.
  const bool fork = overlap_tiles && side != nullptr;
  struct Pending {
    int deadline;
    cudaEvent_t ev;
  };
  std::vector<Pending> pending;
  auto next_event = [&]() {
    if (ev_used == ev_pool.size()) {
      cudaEvent_t e;
      CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ev_pool.push_back(e);
    }
    return ev_pool[ev_used++];
  };
  auto join_bg = [&](int slot) {   // wait for every background launch whose deadline is <= slot
    int lastp = -1;
    for (size_t k = 0; k < pending.size(); ++k)
      if (pending[k].deadline <= slot) lastp = (int)k;
    if (lastp < 0) return;
    CK(cudaStreamWaitEvent(st, pending[lastp].ev, 0));   // bg is in-order: covers the earlier ones
    pending.erase(pending.begin(), pending.begin() + lastp + 1);
  };
  cudaEvent_t ev_panel = nullptr;
  // Look-ahead (tag 3): updates that may run while the NEXT panel launch is executing.  They are
  // held back until that panel launch has been issued -- issued first, they would occupy every SM
  // (persistent tile kernel) and the panel would queue behind them -- then forked onto `ahead`, and
  // joined before the next tile launch / push of the main stream.
  bool ahead_pending = false;
  std::vector<const Launch*> stash;
  auto flush_stash = [&]() {
    if (stash.empty()) return;
    CK(cudaStreamWaitEvent(ahead, ev_fork2, 0));
    for (const Launch* q : stash) launch_one(*q, ahead, false);
    stash.clear();
    CK(cudaEventRecord(ev_ahead, ahead));
    ahead_pending = true;
  };
  auto join_ahead = [&]() {
    flush_stash();
    if (!ahead_pending) return;
    CK(cudaStreamWaitEvent(st, ev_ahead, 0));
    ahead_pending = false;
  };
  // Experiment (opt-in, SPLLT_B200_PUSH_ASYNC=1): deliveries (L_PUSH) on their own stream, so that
  // the owner's main stream does not stand behind world - 1 NVLink copies.  Measured SLOWER on Poisson
  // 100^3 at 4 GPUs (78.9 vs 73.9 ms): the copy CTAs take SM slots from the persistent tile kernels.
  const bool push_async = fork && comm != nullptr && getenv("SPLLT_B200_PUSH_ASYNC");
  bool comm_pending = false;
  int comm_step = -1;
  for (i64 i = first; i < last; ++i) {
    const Launch& L = S.launches[i];
    if (L.kind == L_PUSH && push_async) {
      if (comm_step != L.depth) {       // first delivery of this step: after everything of the chain
        join_ahead();
        CK(cudaEventRecord(ev_fork3, st));
        CK(cudaStreamWaitEvent(comm, ev_fork3, 0));
        comm_step = L.depth;
      }
      launch_one(L, comm, false);
      comm_pending = true;
      continue;
    }
    if (L.tag == 3 && fork && ahead) {
      if (stash.empty()) CK(cudaEventRecord(ev_fork2, st));   // depends on everything issued so far
      stash.push_back(&L);
      continue;
    }
    if (L.kind != L_PANEL) join_ahead();
    if (L.stream == 1 && fork && ev_panel) {
      CK(cudaStreamWaitEvent(bg, ev_panel, 0));
      launch_one(L, bg, true);
      cudaEvent_t e = next_event();
      CK(cudaEventRecord(e, bg));
      pending.push_back({L.deadline, e});
      continue;
    }
    if (L.kind == L_PANEL) {
      join_bg(L.depth);
      launch_one(L, st, false);
      flush_stash();
      if (fork && L.phase == 0) {
        ev_panel = next_event();
        CK(cudaEventRecord(ev_panel, st));
      }
      continue;
    }
    if (fork && L.kind == L_TILE_S && i + 1 < last && S.launches[i + 1].kind == L_TILE_L &&
        S.launches[i + 1].depth == L.depth && S.launches[i + 1].phase == L.phase && S.launches[i + 1].stream == 0 &&
        S.launches[i + 1].tag == L.tag) {
      CK(cudaEventRecord(ev_fork, st));
      CK(cudaStreamWaitEvent(side, ev_fork, 0));
      launch_one(L, side, false);
      launch_one(S.launches[i + 1], st, false);
      CK(cudaEventRecord(ev_join, side));
      CK(cudaStreamWaitEvent(st, ev_join, 0));
      ++i;
      continue;
    }
    launch_one(L, st, false);
  }
  join_ahead();
  if (comm_pending) {
    CK(cudaEventRecord(ev_comm, comm));
    CK(cudaStreamWaitEvent(st, ev_comm, 0));
  }
  join_bg(1 << 30);
}

void Engine::enqueue_factor(const double* dval, cudaStream_t st) {
  const Analysis& S = *A;
  if (!comm_ready) {
    fprintf(stderr, "spllt_b200: %d ranks but the peers are not attached (spllt_b200_comm_attach)\n", S.world);
    throw CudaFailure{cudaErrorNotReady, __FILE__, __LINE__};
  }
  ev_used = 0;
  factor_begin(dval, st);
  factor_barrier(1, 0, st);
  enqueue_range(0, phase0_end, st);
  apply_generated(st);
  factor_barrier(2, 0, st);
  enqueue_range(phase0_end, (i64)S.launches.size(), st);
  factor_end(st);
}

void Engine::factor(const double* dval) {
  upload_tables();
  if (A->n == 0) return;
  // The launch sequence is value independent: capture it once, replay it afterwards.
  if (use_graph && (!factor_graph || graph_val != dval || graph_stream != stream)) {
    if (factor_graph) {
      CK(cudaGraphExecDestroy(factor_graph));
      factor_graph = nullptr;
    }
    cudaGraph_t g;
    CK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
    enqueue_factor(dval, stream);
    CK(cudaStreamEndCapture(stream, &g));
    CK(cudaGraphInstantiate(&factor_graph, g, 0));
    CK(cudaGraphDestroy(g));
    graph_val = dval;
    graph_stream = stream;
  }
  if (use_graph)
    CK(cudaGraphLaunch(factor_graph, stream));
  else
    enqueue_factor(dval, stream);
  factored = true;
  dinv_valid = true;   // enqueue_factor(phase -1) ends with launch_invert_diag
}

// Un-graphed factorization with one event pair per launch: milliseconds per kernel kind
// (assemble+memset, panel, tile_s, tile_l) and, optionally, one CSV line per launch.  Multi-GPU:
// collective -- every rank must call it; pushes / waits / barriers are executed but not attributed.
void Engine::profile_factor(const double* dval, double* ms4, const char* csv) {
  upload_tables();
  dinv_valid = false;
  for (int i = 0; i < 4; ++i) ms4[i] = 0;
  if (A->n == 0) return;
  const Analysis& S = *A;
  std::vector<cudaEvent_t> ev(S.launches.size() + 2);
  for (auto& e : ev) CK(cudaEventCreate(&e));
  cudaStream_t st = stream;
  CK(cudaEventRecord(ev[0], st));
  factor_begin(dval, st);
  CK(cudaEventRecord(ev[1], st));
  factor_barrier(1, 0, st);
  long long* dbg = nullptr;
  const char* dbg_env = getenv("SPLLT_B200_PANEL_DBG");
  int dbg_launch = dbg_env ? atoi(dbg_env) : -1;
  i64 dbg_count = 0;
  std::vector<cudaEvent_t> ev0(S.launches.size());   // start of every launch (barriers / waits excluded)
  for (auto& e : ev0) CK(cudaEventCreate(&e));
  for (size_t i = 0; i < S.launches.size(); ++i) {
    const Launch& L = S.launches[i];
    if ((i64)i == phase0_end) {
      apply_generated(st);
      factor_barrier(2, 0, st);
    }
    CK(cudaEventRecord(ev0[i], st));
    if ((int)i == dbg_launch && L.kind == L_PANEL) {
      dbg_count = L.count;
      CK(cudaMalloc(&dbg, dbg_count * 8 * sizeof(long long)));
      launch_panel_dbg(d_panel + L.begin, L.count, arena, d_info, d_counters + S.launches.size(), dbg, st);
    } else {
      launch_one(L, st, false);
    }
    CK(cudaEventRecord(ev[i + 2], st));
  }
  if (phase0_end == (i64)S.launches.size()) {
    apply_generated(st);
    factor_barrier(2, 0, st);
  }
  CK(cudaStreamSynchronize(st));
  if (dbg) {
    std::vector<long long> h(dbg_count * 8);
    CK(cudaMemcpy(h.data(), dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    for (i64 c = 0; c < std::min<i64>(dbg_count, 4); ++c)
      fprintf(stderr, "panel dbg cta %lld (pw %d nrows %d): load_diag %lld potrf %lld | solve_done@ %lld store %lld\n",
              (long long)c, S.panel_tasks[S.launches[dbg_launch].begin + c].pw,
              S.panel_tasks[S.launches[dbg_launch].begin + c].nrows, h[c * 8 + 1] - h[c * 8 + 0],
              h[c * 8 + 2] - h[c * 8 + 1], h[c * 8 + 3] - h[c * 8 + 1], h[c * 8 + 4] - h[c * 8 + 3]);
    CK(cudaFree(dbg));
  }
  float ms;
  CK(cudaEventElapsedTime(&ms, ev[0], ev[1]));
  ms4[0] = ms;
  FILE* f = csv ? fopen(csv, "w") : nullptr;
  if (f) fprintf(f, "launch,kind,tag,depth,ctas,ms,flops_issued,flops_algo\n");
  for (size_t i = 0; i < S.launches.size(); ++i) {
    const Launch& L = S.launches[i];
    CK(cudaEventElapsedTime(&ms, ev0[i], ev[i + 2]));
    if (L.kind <= L_TILE_L) ms4[1 + L.kind] += ms;
    if (f) {
      double fl = 0, fa = 0;
      if (L.kind == L_TILE_S || L.kind == L_TILE_L) {
        double T = L.kind == L_TILE_L ? 128.0 : 64.0, TN = L.kind == L_TILE_L ? (double)S.tile_n : 64.0;
        for (i64 k = L.begin; k < L.begin + L.count; ++k) {
          fl += 2.0 * T * TN * S.tile_tasks[k].kk;
          fa += tile_algo_flops(S.tile_tasks[k]);
        }
      }
      fprintf(f, "%zu,%d,%d,%d,%lld,%.6f,%.0f,%.0f\n", i, L.kind, L.tag, L.depth, (long long)L.count, ms, fl, fa);
    }
  }
  if (f) fclose(f);
  for (auto& e : ev) CK(cudaEventDestroy(e));
  for (auto& e : ev0) CK(cudaEventDestroy(e));
  factored = true;
}

void Engine::factor_host(const double* val) {
  upload_tables();
  if (A->n == 0) return;
  CK(cudaMemcpyAsync(d_val, val, A->nnz * sizeof(double), cudaMemcpyHostToDevice, stream));
  factor(d_val);
}

// The sweeps work on the internal pivot-order vector d_xw only, so the captured graph does
// not depend on the caller's x: the two permutation kernels are launched around it.
bool Engine::use_pipe(int nrhs) const { return !A->ptasks_f.empty() && nrhs <= A->pipe_max_nrhs; }

void Engine::enqueue_solve(int nrhs, int job, cudaStream_t st) {
  const Analysis& S = *A;
  const bool pipe = use_pipe(nrhs);
  const std::vector<SolveLaunch>& SL = pipe ? S.slaunch : S.slaunch_full;
  if (job == 0 || job == 1) {
    for (int d = 0; d < S.ndepth; ++d) {
      const SolveLaunch& L = SL[d];
      launch_fwd_diag(d_sb + L.diag_begin, L.diag_count, arena, d_xw, nrhs, st);
      launch_fwd_upd(d_su + L.upd_begin, L.upd_count, d_sb, arena, d_index, d_xw, nrhs, st);
    }
    if (pipe)
      launch_solve_pipe(true, d_ptask_f, (int)S.ptasks_f.size(), d_pdest, d_pexpect, arena, d_dinv, d_index, d_xw, d_xm, nrhs,
                        S.nstrips, S.nnodes, S.n, d_psync, st);
  }
  if (job == 0 || job == 2) {
    if (pipe)
      launch_solve_pipe(false, d_ptask_b, (int)S.ptasks_b.size(), d_pdest, d_pexpect, arena, d_dinv, d_index, d_xw, d_xm + xm_doubles, nrhs,
                        S.nstrips, S.nnodes, S.n, d_psync + psync_ints, st);
    for (int d = S.ndepth - 1; d >= 0; --d) {
      const SolveLaunch& L = SL[d];
      if (solve_use_mma(nrhs, true))
        launch_bwd_upd_mma(d_sut + L.updt_begin, L.updt_count, d_sb, arena, d_index, d_xw, nrhs, st);
      else
        launch_bwd_upd(d_su + L.upd_begin, L.upd_count, d_sb, arena, d_index, d_xw, nrhs, st);
      launch_bwd_diag(d_sb + L.diag_begin, L.diag_count, arena, d_xw, nrhs, st);
    }
  }
}

// The pipelined solve multiplies by the inverses of the 64 x 64 diagonal blocks.  A single-GPU
// factorization computes them as its last launch; after any other way of producing the factor
// (multi-GPU phases, diagnostics) they are computed before the first solve.
void Engine::ensure_dinv() {
  if (dinv_valid) return;
  launch_invert_diag(d_pnodes, d_strip_node, A->nstrips, arena, d_dinv, stream);
  dinv_valid = true;
}

// Multi-GPU solve, one call per phase; the caller all-reduces the work vector (xw_ptr) over the
// ranks after phases 1 and 4.  0: permute the right-hand side in and drop the entries this rank
// does not own; 1: forward sweep of the rank's subtrees; 2: forward sweep of the upper tree;
// 3: backward sweep of the upper tree; 4: backward sweep of the rank's subtrees, then drop the
// entries it does not own; 5: permute the solution out.
void Engine::solve_phase(double* dx, int ldx, int nrhs, int phase) {
  upload_tables();
  if (A->n == 0 || nrhs <= 0) return;
  ensure_solve_buffers(nrhs);
  ensure_dinv();
  const Analysis& S = *A;
  cudaStream_t st = stream;
  switch (phase) {
    case 0:
      launch_permute_in(dx, ldx, d_porder, d_xw, S.n, nrhs, st);
      launch_mask_rows(d_xw, d_col_keep, S.n, nrhs, st);
      break;
    case 1:
      launch_solve_pipe(true, d_ptask_f, (int)S.ptasks_f.size(), d_pdest, d_pexpect, arena, d_dinv, d_index, d_xw, d_xm, nrhs,
                        S.nstrips, S.nnodes, S.n, d_psync, st);
      break;
    case 2:
      launch_solve_pipe(true, d_ptask_ft, (int)S.ptasks_ft.size(), d_pdest, d_pexpect_top, arena, d_dinv, d_index, d_xw, d_xm, nrhs, S.nstrips, S.nnodes, S.n, d_psync, st);
      break;
    case 3:
      launch_solve_pipe(false, d_ptask_bt, (int)S.ptasks_bt.size(), d_pdest, d_pexpect_top, arena, d_dinv, d_index, d_xw, d_xm + xm_doubles, nrhs, S.nstrips, S.nnodes, S.n, d_psync + psync_ints, st);
      break;
    case 4:   // same flag region as phase 3: the subtrees wait on the flags of the upper tree
      launch_solve_pipe(false, d_ptask_b, (int)S.ptasks_b.size(), d_pdest, d_pexpect, arena, d_dinv, d_index, d_xw, d_xm + xm_doubles, nrhs,
                        S.nstrips, S.nnodes, S.n, d_psync + psync_ints, st, nullptr, true);
      launch_mask_rows(d_xw, d_col_keep, S.n, nrhs, st);
      break;
    case 5:
      launch_permute_out(dx, ldx, d_porder, d_xw, S.n, nrhs, st);
      break;
  }
}

void Engine::solve(double* dx, int ldx, int nrhs, int job) {
  upload_tables();
  if (A->n == 0 || nrhs <= 0) return;
  ensure_solve_buffers(nrhs);
  ensure_dinv();
  if (job == 0 || job == 1) launch_permute_in(dx, ldx, d_porder, d_xw, A->n, nrhs, stream);
  if (use_graph) {
    SolveGraphKey key{nullptr, 0, nrhs, job, stream, d_xw};
    cudaGraphExec_t ex = nullptr;
    for (auto& e : solve_graphs)
      if (e.first == key) ex = e.second;
    if (!ex) {
      cudaGraph_t g;
      CK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
      enqueue_solve(nrhs, job, stream);
      CK(cudaStreamEndCapture(stream, &g));
      CK(cudaGraphInstantiate(&ex, g, 0));
      CK(cudaGraphDestroy(g));
      if (solve_graphs.size() >= 16) {
        CK(cudaGraphExecDestroy(solve_graphs.front().second));
        solve_graphs.erase(solve_graphs.begin());
      }
      solve_graphs.push_back({key, ex});
    }
    CK(cudaGraphLaunch(ex, stream));
  } else {
    enqueue_solve(nrhs, job, stream);
  }
  if (job == 0 || job == 2) launch_permute_out(dx, ldx, d_porder, d_xw, A->n, nrhs, stream);
}

// Un-graphed forward + backward solve with one event pair per launch: ms6 = {fwd_diag, fwd_upd,
// bwd_upd, bwd_diag, fwd_pipe, bwd_pipe}; csv (optional) gets one line per launch.  Diagnostic only.
void Engine::profile_solve(double* dx, int ldx, int nrhs, double* ms6, const char* csv) {
  upload_tables();
  for (int i = 0; i < 6; ++i) ms6[i] = 0;
  if (A->n == 0) return;
  ensure_solve_buffers(nrhs);
  ensure_dinv();
  const Analysis& S = *A;
  const bool pipe = use_pipe(nrhs);
  const std::vector<SolveLaunch>& SL = pipe ? S.slaunch : S.slaunch_full;
  cudaStream_t st = stream;
  struct Rec {
    int kind, depth;
    long long ctas;
  };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> ev;
  auto mark = [&]() {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    CK(cudaEventRecord(e, st));
    ev.push_back(e);
  };
  launch_permute_in(dx, ldx, d_porder, d_xw, S.n, nrhs, st);
  mark();
  for (int d = 0; d < S.ndepth; ++d) {
    const SolveLaunch& L = SL[d];
    if (L.diag_count == 0 && L.upd_count == 0) continue;
    launch_fwd_diag(d_sb + L.diag_begin, L.diag_count, arena, d_xw, nrhs, st);
    mark();
    recs.push_back({0, d, L.diag_count});
    launch_fwd_upd(d_su + L.upd_begin, L.upd_count, d_sb, arena, d_index, d_xw, nrhs, st);
    mark();
    recs.push_back({1, d, L.upd_count});
  }
  if (pipe) {
  launch_solve_pipe(true, d_ptask_f, (int)S.ptasks_f.size(), d_pdest, d_pexpect, arena, d_dinv, d_index, d_xw, d_xm, nrhs, S.nstrips,
                    S.nnodes, S.n, d_psync, st);
  mark();
  recs.push_back({4, -1, (long long)S.ptasks_f.size()});
  launch_solve_pipe(false, d_ptask_b, (int)S.ptasks_b.size(), d_pdest, d_pexpect, arena, d_dinv, d_index, d_xw, d_xm + xm_doubles, nrhs,
                    S.nstrips, S.nnodes, S.n, d_psync + psync_ints, st);
  mark();
  recs.push_back({5, -1, (long long)S.ptasks_b.size()});
  }
  for (int d = S.ndepth - 1; d >= 0; --d) {
    const SolveLaunch& L = SL[d];
    if (L.diag_count == 0 && L.upd_count == 0) continue;
    if (solve_use_mma(nrhs, true))
        launch_bwd_upd_mma(d_sut + L.updt_begin, L.updt_count, d_sb, arena, d_index, d_xw, nrhs, st);
      else
        launch_bwd_upd(d_su + L.upd_begin, L.upd_count, d_sb, arena, d_index, d_xw, nrhs, st);
    mark();
    recs.push_back({2, d, L.upd_count});
    launch_bwd_diag(d_sb + L.diag_begin, L.diag_count, arena, d_xw, nrhs, st);
    mark();
    recs.push_back({3, d, L.diag_count});
  }
  launch_permute_out(dx, ldx, d_porder, d_xw, S.n, nrhs, st);
  CK(cudaStreamSynchronize(st));
  FILE* f = csv ? fopen(csv, "w") : nullptr;
  if (f) fprintf(f, "kind,depth,ctas,ms\n");
  for (size_t i = 0; i < recs.size(); ++i) {
    float ms;
    CK(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
    ms6[recs[i].kind] += ms;
    if (f) fprintf(f, "%d,%d,%lld,%.6f\n", recs[i].kind, recs[i].depth, recs[i].ctas, ms);
  }
  if (f) fclose(f);
  for (auto& x : ev) CK(cudaEventDestroy(x));
}

// One un-graphed forward + backward solve with per-task time stamps (globaltimer, ns) from the
// persistent kernels: out_f / out_b get 8 stamps per claimed task (claim order x rhs chunk):
// start, waits satisfied, solved, end, + 4 strip-internal stamps.  Diagnostic only.
void Engine::trace_solve(double* dx, int ldx, int nrhs, unsigned long long* out_f, unsigned long long* out_b) {
  upload_tables();
  if (A->n == 0) return;
  ensure_solve_buffers(nrhs);
  ensure_dinv();
  const Analysis& S = *A;
  const size_t nf = S.ptasks_f.size() * (size_t)pipe_chunks(nrhs) * 8, nbk = S.ptasks_b.size() * (size_t)pipe_chunks(nrhs) * 8;
  unsigned long long *tf = nullptr, *tb = nullptr;
  CK(cudaMalloc(&tf, std::max<size_t>(nf, 1) * 8));
  CK(cudaMalloc(&tb, std::max<size_t>(nbk, 1) * 8));
  CK(cudaMemset(tf, 0, std::max<size_t>(nf, 1) * 8));
  CK(cudaMemset(tb, 0, std::max<size_t>(nbk, 1) * 8));
  launch_permute_in(dx, ldx, d_porder, d_xw, S.n, nrhs, stream);
  for (int d = 0; d < S.ndepth; ++d) {
    const SolveLaunch& L = S.slaunch[d];
    launch_fwd_diag(d_sb + L.diag_begin, L.diag_count, arena, d_xw, nrhs, stream);
    launch_fwd_upd(d_su + L.upd_begin, L.upd_count, d_sb, arena, d_index, d_xw, nrhs, stream);
  }
  launch_solve_pipe(true, d_ptask_f, (int)S.ptasks_f.size(), d_pdest, d_pexpect, arena, d_dinv, d_index, d_xw, d_xm, nrhs, S.nstrips,
                    S.nnodes, S.n, d_psync, stream, tf);
  launch_solve_pipe(false, d_ptask_b, (int)S.ptasks_b.size(), d_pdest, d_pexpect, arena, d_dinv, d_index, d_xw, d_xm + xm_doubles, nrhs,
                    S.nstrips, S.nnodes, S.n, d_psync + psync_ints, stream, tb);
  for (int d = S.ndepth - 1; d >= 0; --d) {
    const SolveLaunch& L = S.slaunch[d];
    if (solve_use_mma(nrhs, true))
      launch_bwd_upd_mma(d_sut + L.updt_begin, L.updt_count, d_sb, arena, d_index, d_xw, nrhs, stream);
    else
      launch_bwd_upd(d_su + L.upd_begin, L.upd_count, d_sb, arena, d_index, d_xw, nrhs, stream);
    launch_bwd_diag(d_sb + L.diag_begin, L.diag_count, arena, d_xw, nrhs, stream);
  }
  launch_permute_out(dx, ldx, d_porder, d_xw, S.n, nrhs, stream);
  CK(cudaStreamSynchronize(stream));
  CK(cudaMemcpy(out_f, tf, nf * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(out_b, tb, nbk * 8, cudaMemcpyDeviceToHost));
  CK(cudaFree(tf));
  CK(cudaFree(tb));
}

void Engine::solve_host(double* x, int nrhs, int job) {
  upload_tables();
  if (A->n == 0 || nrhs <= 0) return;
  ensure_solve_buffers(nrhs);
  size_t bytes = (size_t)A->n * nrhs * sizeof(double);
  // job 2 continues from the device-resident forward result; x is only an output then
  if (job != 2) CK(cudaMemcpyAsync(d_x, x, bytes, cudaMemcpyHostToDevice, stream));
  solve(d_x, A->n, nrhs, job);
  if (job != 1) CK(cudaMemcpyAsync(x, d_x, bytes, cudaMemcpyDeviceToHost, stream));
}

void Engine::sync() {
  if (stream) CK(cudaStreamSynchronize(stream));
}

int Engine::pivot_flag() {
  if (!uploaded || !factored) return 0;
  int v = 0;
  sync();
  CK(cudaMemcpy(&v, d_info, sizeof(int), cudaMemcpyDeviceToHost));
  return v >= 0x7f000000 ? 0 : v;
}

void Engine::get_lcol(int g, double* out) {
  // block column g (0-based) in the reference layout: rows [c*nb, m), cols [c*nb, c*nb+w), ld = w
  const Analysis& S = *A;
  const HNode& nd = S.nodes[S.bcol_node[g]];
  int r0 = S.bcol_c[g] * S.nb;
  int w = std::min(S.nb, nd.n - r0), h = nd.m - r0;
  sync();
  CK(cudaMemcpy2D(out, (size_t)w * sizeof(double), arena + nd.off + (i64)r0 * nd.ld + r0, (size_t)nd.ld * sizeof(double),
                  (size_t)w * sizeof(double), h, cudaMemcpyDeviceToHost));
}

void Engine::get_fwd(int nrhs, double* out) {
  sync();
  CK(cudaMemcpy(out, d_xw, (size_t)A->n * nrhs * sizeof(double), cudaMemcpyDeviceToHost));
}

void Engine::release() {
  if (!uploaded) return;
  cudaDeviceSynchronize();
  if (factor_graph) cudaGraphExecDestroy(factor_graph);
  for (auto& e : solve_graphs) cudaGraphExecDestroy(e.second);
  solve_graphs.clear();
  factor_graph = nullptr;
  cudaFree(arena);
  cudaFree(d_lmap_dst);
  cudaFree(d_lmap_src);
  cudaFree(d_val);
  cudaFree(d_panel);
  cudaFree(d_tile);
  if (d_qbase) cudaFree(d_qbase);
  cudaFree(d_qld);
  cudaFree(d_qrp);
  cudaFree(d_rowpos);
  cudaFree(d_info);
  cudaFree(d_counters);
  for (void* p : ipc_open) cudaIpcCloseMemHandle(p);
  ipc_open.clear();
  cudaFree(d_flags);
  cudaFree(d_pushcnt);
  if (d_gen) cudaFree(d_gen);
  if (d_gqbase) cudaFree(d_gqbase);
  if (d_gqld) cudaFree(d_gqld);
  if (d_gqrp) cudaFree(d_gqrp);
  d_gen = nullptr;
  d_gqbase = nullptr;
  d_gqld = nullptr;
  d_gqrp = nullptr;
  d_flags = d_pushcnt = nullptr;
  d_qbase = nullptr;
  comm_ready = maps_ready = false;
  peers = PeerSet{};
  if (d_tmaps) cudaFree(d_tmaps);
  if (d_tmaps_b) cudaFree(d_tmaps_b);
  d_tmaps = d_tmaps_b = nullptr;
  cudaFree(d_sb);
  cudaFree(d_su);
  cudaFree(d_sut);
  cudaFree(d_pnodes);
  cudaFree(d_ptask_f);
  cudaFree(d_ptask_b);
  cudaFree(d_pdest);
  cudaFree(d_strip_node);
  cudaFree(d_pexpect);
  cudaFree(d_ptask_ft);
  cudaFree(d_ptask_bt);
  cudaFree(d_pexpect_top);
  cudaFree(d_col_keep);
  cudaFree(d_dinv);
  if (d_psync) cudaFree(d_psync);
  d_psync = nullptr;
  if (d_xm) cudaFree(d_xm);
  d_xm = nullptr;
  cudaFree(d_index);
  cudaFree(d_porder);
  if (d_xw) cudaFree(d_xw);
  if (d_x) cudaFree(d_x);
  // nothing of the old analysis may survive: a later spllt_analyse on the same handles starts clean
  d_xw = d_x = nullptr;
  xw_nrhs = 0;
  xm_doubles = psync_ints = 0;
  dinv_valid = factored = false;
  graph_val = nullptr;
  graph_stream = nullptr;
  host_y = nullptr;
  if (stream == own) stream = nullptr;
  if (own_stream) cudaStreamDestroy(own);
  own_stream = false;
  own = nullptr;
  if (side) {
    cudaStreamDestroy(side);
    cudaStreamDestroy(bg);
    for (cudaEvent_t e : ev_pool) cudaEventDestroy(e);
    ev_pool.clear();
    bg = nullptr;
    cudaEventDestroy(ev_fork);
    cudaEventDestroy(ev_join);
    if (comm) {
      cudaStreamDestroy(comm);
      cudaEventDestroy(ev_fork3);
      cudaEventDestroy(ev_comm);
      comm = nullptr;
    }
    if (ahead) {
      cudaStreamDestroy(ahead);
      cudaEventDestroy(ev_fork2);
      cudaEventDestroy(ev_ahead);
      ahead = nullptr;
    }
    side = nullptr;
  }
  uploaded = false;
}

}  // namespace spllt
