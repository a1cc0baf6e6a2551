// Flat (SoA) data model of the B200 build: what spllt_data_mod's derived types
// (src/spllt_data_mod.F90:83-388: spllt_block, spllt_node, lfactor, lmap_type,
// spllt_sblock_t, spllt_akeep, spllt_fkeep) become when node / block storage moves to HBM.
//
// Host side keeps the symbolic content (value independent, produced by spllt_analyse);
// the device side holds one HBM arena for all supernodes plus the work lists that
// replace the OpenMP / StarPU task DAG (src/spllt_factorization_task_mod.F90).
//
// HBM layout: every supernode is ONE row-major m x ld matrix (ld = n rounded up to 4,
// padding columns stay zero).  Block column c of the reference (lfact(bcol)%lcol, row-major
// tiles with ld = blkn, src/spllt_analyse_mod.F90:1160-1165) is the sub-matrix
// rows [c*nb, m) x cols [c*nb, c*nb+blkn) of that matrix, so the reference layout is
// recovered by a strided copy (capi: spllt_b200_get_factor).
//
// Internal indices are 0-based; getters convert to the reference's 1-based tables so they
// can be compared bit-for-bit with the oracle.
#pragma once
#include <cstdint>
#include <vector>

#include "symbolic.h"

namespace spllt {

typedef int64_t i64;

constexpr int IB = 64;      // inner panel width of the block-column factorization
constexpr int LDPAD = 4;    // node leading dimension is a multiple of 4 doubles (32 B)
constexpr int TRSM_ROWS = 128;  // rows per CTA of the panel solve (multiple of 32)
constexpr int SOLVE_ROWS = 64;  // rows per CTA of the solve update kernels

// ------------------------------------------------------------------ host symbolic tables
struct HNode {
  int sa, en;        // first / last column (0-based, pivot order)
  int m, n;          // rows, columns
  i64 idx_off;       // node index list = index[idx_off .. idx_off+m)
  int parent;        // 0-based, -1 for roots (the reference's virtual root nnodes+1)
  int nchild;
  int least_desc;    // first node of the subtree (postorder => subtree = [least_desc, self])
  int nc, nr;        // block columns, block rows
  int bcol0;         // first global block column (0-based)
  i64 blk0;          // first tile id (0-based)
  int depth0;        // schedule depth of block column 0 (block column c runs at depth0 + c)
  int small;         // pruning mark (reference semantics: 0, 1, -root(1-based))
  int owner;         // GPU rank that owns the node (multi-GPU); -1 = shared top of the tree
  int ld;            // leading dimension in the arena
  i64 off;           // arena offset (doubles)
  i64 row_base;      // first below-diagonal row of this node in the q_* update maps
};

// ------------------------------------------------------------------ device work lists
// One inner panel of a block column (a1 + a2, src/spllt_kernels_mod.F90:1168-1189, :1217-1229):
// every CTA factorizes the pw x pw diagonal block in shared memory (redundantly -- it is on
// the critical path anyway and this saves a launch) and then solves its own chunk of rows,
//   rows <- rows * L_pp^-T.
// L_pp is stored in place by exactly one CTA of the panel: every CTA draws a ticket from the
// panel's counter once it has read the un-factorized block, and the CTA that draws the last
// ticket stores (nobody waits for anybody).  `store` marks one task per panel for bookkeeping.
struct PanelTask {
  i64 d_off;         // arena offset of the pw x pw diagonal block
  i64 r_off;         // first row of this chunk (below the diagonal block)
  int ld, pw, nrows;
  int col0;          // global pivot column of the panel's first column (error report)
  int store;         // 1: this CTA stores L_pp
  int group;         // panel id (index of the panel's counter)
  int ngroup;        // CTAs working on this panel
  int pad;
};
// One dense tile update (a3 intra-node, a4 inter-node):
//   C[i, j] -= sum_{k in [k0, k0+kk)} L[i, k] * L[j, k],  i in [i0,i0+mt), j in [j0,j0+nt), i >= j
// rows/cols are row numbers of the SOURCE node's matrix.  src < 0: C is the same node
// (column j of the node == row j).  src >= 0: scatter through the q_* maps into ancestors.
struct TileTask {
  i64 off;           // arena offset of the source node
  int ld;
  int i0, j0, k0;
  int mt, nt, kk;
  int src;
  i64 qoff;          // row_base - n of the source (index of row r in the maps = qoff + r)
  int node, pad;     // source node (selects its TMA tensor map)
};

// Multi-GPU (one rank per GPU, peer-mapped arenas, see Engine): the upper tree is walked in STEPS,
// one per upper-tree block column, in the same global order on every rank.
//   L_PUSH : this rank owns the block column of the step and has just factorized it: copy it into
//            the arenas of the peers in the mask `deadline` over NVLink and raise its flag there
//            (begin = global block column, count = 0 .. world-2: ordinal of this push of the step)
//   L_WAIT : another rank owns it: wait for the flag (begin = global block column)
enum LaunchKind { L_PANEL = 0, L_TILE_S = 1, L_TILE_L = 2, L_PUSH = 3, L_WAIT = 4, L_NKIND = 5 };
struct Launch {
  int kind;
  int depth;         // phase 0: panel slot; phase 1 (multi-GPU upper tree): step index
  i64 begin;         // first task in the list of this kind
  i64 count;         // tasks (= CTAs)
  int phase;         // multi-GPU: 0 = subtrees owned by this rank, 1 = upper tree
  int tag;           // 0 panel, 4 updates on the critical path, 3 updates overlapped with the next panel of a
                     // chain (side stream, joined before the next tile launch), 5 deferred updates (previous step),
                     // 7 urgent updates of a step (destinations whose turn comes next), 8 push, 9 wait
  int stream;        // 0 = main; 1 = background stream (deferred inter-node updates)
  int deadline;      // background launches: the slot whose panel launch must wait for them
};
// Multi-GPU: the generated element of one subtree this rank owns (src/spllt_kernels_mod.F90:780-821
// `buffer`, size (m - n)^2 of the subtree root): every update of a subtree node into the upper tree is
// accumulated HERE, in local HBM, and scattered once into the owners of the upper-tree block columns
// when the subtree is complete (a9: spllt_subtree_apply_buffer / spllt_scatter_block).
struct GenElem {
  int root;          // subtree root node
  int b;             // rows below the root's diagonal block = side of the element
  i64 off;           // offset (doubles) in the rank's element buffer; entry (i, j), i >= j, at off + i * b + j
  i64 map0;          // gq_* index of the root's first below-diagonal row
};
struct TopStep {     // one upper-tree block column in the global step order
  int node, c;       // node (0-based), local block column
  int owner;         // rank that factorizes it and accumulates the updates into it
  int slot;          // as-soon-as-possible panel slot of its first panel (sort key)
};

// ------------------------------------------------------------------ solve work lists
// One block column in the supernodal triangular solves (a13/a14).
struct SolveBcol {
  i64 off;           // arena offset of the node
  i64 idx_off;       // node index list
  int ld, m;
  int r0, w;         // first row/col of the block column inside the node, width
  int sa;            // first pivot column of the node
  int pad;
};
struct SolveUpd {    // rows [r, r+nrows) of block column `bc` (index into the SolveBcol list)
  int bc, r, nrows, pad;
};
struct SolveUpdT {   // backward update on DMMA tiles: rows [r, r+nrows) x columns [k0, k0+64) of block column `bc`
  int bc, r, nrows, k0;
};
constexpr int SOLVE_ROWS_T = 512;   // rows per SolveUpdT task
struct SolveLaunch {
  i64 diag_begin, diag_count;   // SolveBcol range
  i64 upd_begin, upd_count;     // SolveUpd range
  i64 updt_begin, updt_count;   // SolveUpdT range (many right-hand sides)
};

// ------------------------------------------------------------------ pipelined solve (solve_pipe.cu)
// The default solve: ONE persistent kernel per sweep.  CTAs claim tasks from a global counter in
// a topological order (forward: children before parents; backward: the reverse) and synchronise
// through flags / counters in HBM instead of kernel boundaries -- the device-resident
// replacement of the solve task DAG (fwd_wdep / bwd_wdep, src/spllt_solve_dep_mod.F90:27-248).
// Granularity: 64-row strips of a node's diagonal block (a13) and 64-row chunks of the rows
// below it (a14/a15), independent of the factorization's nb.
constexpr int PS = 64;           // strip width of the pipelined solve
constexpr int PIPE_SMALL_ROWS = 256;  // nodes with n <= PS and m - n <= this are one fused task
constexpr int PIPE_FAT_NP = 4;        // nodes with at most this many strips: rows below them are streamed in
constexpr int PIPE_FAT_ROWS = 512;    //   chunks of up to this many rows after the whole node is solved
constexpr int PIPE_LEVEL_TASKS = 512; // ... sized so that a tree level yields about this many chunks
constexpr double PIPE_WIDE_FRAC_MAX = 0.6;   // above this share of L in wide nodes the level-set path is kept
constexpr int PIPE_TASK_BYTES = 64 * 1024;  // ... and a chunk streams at least this many bytes of L
enum PipeKind { P_DIAG = 0, P_BELOW = 1, P_SMALL = 2 };
struct PNode {
  i64 off;           // arena offset of the node
  i64 idx_off;       // node index list
  int ld, m, n, sa;
  int strip0;        // flag index of the node's first strip
  int np;            // strips = ceil(n / PS)
  int expect_f;      // forward: (task, strip) contributions into this node (informational; waits are per strip)
  int expect_b;      // backward: below-chunk tasks of this node
  int pflag;         // flag index of the parent's strip 0 (-1: root); informational
  int pad[3];
};
struct PTask {
  int node, kind;
  int r0, nrows;     // P_DIAG: r0 = strip index; P_BELOW: rows [r0, r0 + nrows) of the node (r0 >= n)
  int dest_begin, dest_count;  // ancestor strips its rows map to (pipe_dest): forward = counters it bumps,
                               // backward = flags it waits for
  int pad[2];
};

struct PTaskD {      // what the kernels read: task + its node in ONE 96-byte record (one load per claim)
  PTask t;
  PNode n;
};

// ------------------------------------------------------------------ reference-format tables
struct RefBlock {    // spllt_block, 1-based (src/spllt_data_mod.F90:123-172)
  i64 id;
  int blkm, blkn;
  i64 sa, dblk, last_blk;
  int node, bcol, dep_initial;
};

struct Analysis {
  int n = 0, nb = 0, nemin = 32, ncpu = 1, prune = 1, min_width_blas = 8;
  int rank = 0, world = 1;   // multi-GPU partition (partition_tree)
  int tile_n = 128;          // N extent of the large tiles: 128 (one CTA/SM) or 64 (two CTAs/SM)
  int dist_top = 0;          // multi-GPU: upper tree distributed block-column-cyclically (owner computes)
  std::vector<int> bcol_owner;  // [nbcol] rank that factorizes / receives the updates of a block column
  std::vector<TopStep> top_steps;   // multi-GPU: upper-tree block columns in step order (same on every rank)
  std::vector<int> bcol_step;       // [nbcol] step index of an upper-tree block column, -1 otherwise
  i64 own_begin = 0, own_end = 0;   // arena slice holding the subtrees this rank owns
  i64 nnz = 0;               // entries of the user's lower triangle
  Symbolic sym;
  std::vector<int> porder;   // porder[p] = variable at pivot position p (0-based both)

  int nnodes = 0, nbcol = 0, ndepth = 0, maxmn = 0;
  i64 final_blk = 0;
  std::vector<HNode> nodes;
  std::vector<int> index;          // concatenated row lists (0-based pivot indices)
  std::vector<i64> weight;         // [nnodes+1] flops per subtree (+ virtual root)
  std::vector<int> col2node;       // pivot column -> node

  // A -> L map (spllt_make_map + spllt_lcol_map), grouped by block column
  std::vector<i64> lmap_ptr;       // [nbcol+1]
  std::vector<int> lmap_row;       // row inside the NODE (0-based)
  std::vector<int> lmap_col;       // column inside the NODE (0-based)
  std::vector<i64> lmap_src;       // 0-based index into the user's val
  std::vector<int> bcol_node;      // [nbcol] owning node
  std::vector<int> bcol_c;         // [nbcol] local block-column index

  i64 arena = 0;                   // device arena size (doubles)
  i64 top_begin = 0;               // arena offset where the upper tree (small == 0) starts
  i64 num_factor = 0;              // entries of L (trapezoids, reference count)
  i64 num_flops = 0;               // sum_nodes sum_j (m-n+j)^2  (akeep%weight(nnodes+1))

  // factor schedule
  std::vector<PanelTask> panel_tasks;
  int npanel_groups = 0;
  std::vector<TileTask> tile_tasks;
  std::vector<Launch> launches;
  std::vector<i64> q_base;         // per below-diagonal row: dest address when used as a COLUMN
  std::vector<int> q_ld;           //   leading dimension of that dest node
  std::vector<i64> q_rp;           //   rowpos[q_rp[j] + r] = row position of source row r in dest
  std::vector<int> rowpos;
  // multi-GPU: generated elements of the subtrees this rank owns; q_base < 0 encodes a destination
  // inside the element buffer (-(1 + offset)); gq_*: where an element's columns / rows land in the
  // upper tree (same meaning as q_*, indexed from GenElem::map0)
  std::vector<GenElem> gen;
  i64 gen_doubles = 0;
  std::vector<i64> gq_base;        // arena offset of the destination column
  std::vector<int> gq_ld, gq_bcol; // leading dimension, global block column (-> owner) of the destination
  std::vector<i64> gq_rp;          // rowpos[gq_rp + i] = row position of element row i in that destination node
  double tile_flops = 0;           // flops issued by the tile kernels (incl. masked halves)
  double tile_flops_algo = 0;      // algorithmic flops of the same tiles (2 kk per entry with i >= j)

  // solve schedule (forward order; the backward sweep walks it in reverse)
  std::vector<SolveBcol> sbcols;
  std::vector<SolveUpd> supds;
  std::vector<SolveUpdT> supds_t;
  std::vector<SolveLaunch> slaunch;   // one per depth: nodes below solve_cut (none by default)
  std::vector<SolveLaunch> slaunch_full;  // one per depth: ALL nodes (used when nrhs > pipe_max_nrhs)
  int pipe_max_nrhs = 8;              // more right-hand sides than this: level-set launches (RC = 8 kernels)
  double wide_frac = 0;               // share of L's entries in nodes wider than PIPE_FAT_NP strips
  // pipelined solve: nodes with depth0 >= solve_cut run in the persistent kernels, the rest
  // (none by default) in the level-set launches above
  int solve_cut = 0;
  int nstrips = 0;
  std::vector<PNode> pnodes;          // [nnodes]
  std::vector<PTask> ptasks_f, ptasks_b;
  std::vector<int> pipe_dest;
  std::vector<int> strip_node;        // [nstrips] owning node
  std::vector<int> pexpect;           // [nstrips] forward: tasks that add into the strip's rows
  // multi-GPU: ptasks_f / ptasks_b hold the subtrees this rank owns, these the shared upper tree
  std::vector<PTask> ptasks_ft, ptasks_bt;
  std::vector<int> pexpect_top;       // [nstrips] expected counts within the upper-tree lists
  std::vector<char> col_keep;         // [n] pivot columns kept by this rank when the work vector is all-reduced
};

// analyse.cpp
int build_analysis(int n, const int* ptr, const int* row, int nb, int nemin, int ncpu, int prune,
                   int ordering, const int* user_order, Analysis& A);
void prune_tree(Analysis& A, int nth, std::vector<int>& small);
void partition_tree(Analysis& A, int rank, int world);
void build_factor_schedule(Analysis& A, int tile_l_min);
void build_solve_schedule(Analysis& A);
void ref_blocks(const Analysis& A, std::vector<RefBlock>& out);
double tile_algo_flops(const TileTask& t);

}  // namespace spllt
