/* B200-specific additions to the C ABI (everything the reference header cannot express):
 * device-resident factor / solve entry points, 64-bit counters, table getters used by the
 * parity tests, stream control for event timing, and the multi-GPU hooks.
 *
 * Plain pointers and sizes only.  `d_` arguments are device pointers on the current CUDA
 * device; `stream` is a cudaStream_t passed as void* (NULL = the library's own stream).
 */
#ifndef SPLLT_B200_H
#define SPLLT_B200_H

#include "spllt_iface.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ordering selector for spllt_b200_analyse (the reference always uses METIS,
 * src/spllt_analyse_mod.F90:109): 1 = METIS nested dissection, 0 = natural, 2 = `order` is input */
void spllt_b200_analyse(void **akeep, void **fkeep, spllt_options_t *options, int n, const int *ptr,
                        const int *row, spllt_inform_t *info, int *order, int ordering);

/* ---- 64-bit counters (spllt_inform_t truncates, interfaces/C/spllt_data_ciface.F90:77-78) */
long long spllt_b200_num_factor(void *akeep);   /* entries of L                                */
long long spllt_b200_num_flops(void *akeep);    /* akeep%weight(nnodes+1), analyse_mod:1013-1021 */
long long spllt_b200_arena_doubles(void *akeep); /* HBM arena size                              */
int spllt_b200_num_nodes(void *akeep);
int spllt_b200_num_bcol(void *akeep);
long long spllt_b200_final_blk(void *akeep);
int spllt_b200_maxmn(void *akeep);
int spllt_b200_num_depth(void *akeep);          /* block-column level sets in the schedule     */

/* ---- symbolic tables, reference numbering (1-based), for bit-exact comparison
 * sptr[nnodes+1], sparent[nnodes], rptr[nnodes+1] (64-bit), rlist[rptr[nnodes]-1] */
void spllt_b200_get_symbolic(void *akeep, int *sptr, int *sparent, long long *rptr, int *rlist);
long long spllt_b200_rlist_len(void *akeep);
/* 9 columns per tile: id blkm blkn sa dblk last_blk node bcol dep_initial (spllt_block) */
void spllt_b200_get_blocks(void *akeep, long long *out);
/* 8 columns per node: sa en parent nchild least_desc nb blk_sa blk_en (spllt_node) */
void spllt_b200_get_nodes(void *akeep, long long *out);
void spllt_b200_get_small(void *akeep, int *out);          /* akeep%small(1:nnodes)  */
void spllt_b200_get_weight(void *akeep, long long *out);   /* akeep%weight(1:nnodes+1) */
long long spllt_b200_lmap_len(void *akeep, int bcol);      /* bcol 1-based */
void spllt_b200_get_lmap(void *akeep, int bcol, long long *dst, long long *src); /* lmap(bcol)%map(1:2,:) */
/* solve tiles of get_solve_blocks (src/spllt_solve_dep_mod.F90:1861-2030):
 * 9 columns: id blkm blkn sa dblk last_blk bcol node ldu */
int spllt_b200_num_sblocks(void *akeep, int nb);
void spllt_b200_get_sblocks(void *akeep, int nb, int *out);

/* ---- factor entries in the reference layout (lfact(bcol)%lcol, row-major tiles) */
long long spllt_b200_lcol_size(void *akeep, int bcol);
void spllt_b200_get_lcol(void *fkeep, int bcol, double *out);   /* device -> host, synchronises */
long long spllt_b200_factor_size(void *akeep);
void spllt_b200_get_factor(void *fkeep, double *out);

/* ---- device-resident hot path */
void spllt_b200_set_stream(void *fkeep, void *stream);
/* d_val: the user's val array already in HBM (nnz doubles).  Asynchronous on the stream. */
void spllt_b200_factor_dev(void *akeep, void *fkeep, const double *d_val, spllt_inform_t *info);
/* d_x: n x nrhs column-major (ldx >= n) in HBM, overwritten.  Asynchronous on the stream. */
void spllt_b200_solve_dev(void *fkeep, int nrhs, double *d_x, int ldx, int job, spllt_inform_t *info);
/* forward result of job 1 in pivot order, n x nrhs row-major -> host (synchronises) */
void spllt_b200_get_fwd(void *fkeep, int nrhs, double *out);
/* reads the device pivot flag (synchronises): 0 = ok, else 1-based pivot column that failed */
int spllt_b200_pivot_flag(void *fkeep);

/* spllt_chkerr without printing: err[nrhs] receives the scaled backward errors
 * (src/utils_mod.F90:432-478); returns how many are <= 1e-14 */
int spllt_b200_chkerr(int n, const int *ptr, const int *row, const double *val, int nrhs, const double *x,
                      const double *rhs, double *err);

/* ---- counters for bench.py */
long long spllt_b200_factor_launches(void *fkeep);  /* kernels per spllt_factor             */
long long spllt_b200_solve_launches(void *fkeep, int job);
double spllt_b200_tile_flops(void *akeep);          /* flops issued by the DMMA tile kernels (padded tiles) */
double spllt_b200_tile_flops_algo(void *akeep);     /* algorithmic flops of the same tiles: 2 K per entry i >= j */
/* launch counts of one factorization: panel, tile_s, tile_l, total */
void spllt_b200_launch_breakdown(void *akeep, long long *out4);

/* diagnostic: one un-graphed factorization timed launch by launch; ms4 = milliseconds spent in
 * {memset+assemble, panel (potrf+trsm), 64x64 tiles, 128x128 tiles}; csv (may be NULL) receives
 * one line per launch.  Synchronises. */
void spllt_b200_profile_factor(void *fkeep, const double *d_val, double *ms4, const char *csv);

/* same for one forward + backward solve: ms6 = {fwd_diag, fwd_upd, bwd_upd, bwd_diag (level-set
 * launches, only used below SPLLT_B200_SOLVE_CUT), fwd_pipe, bwd_pipe (persistent kernels)} */
void spllt_b200_profile_solve(void *fkeep, int nrhs, double *d_x, int ldx, double *ms6, const char *csv);

/* pipelined-solve work lists (value independent; tests replay them to prove the claim order is
 * deadlock free).  sizes: out4 = {forward tasks, backward tasks, strips, dest entries};
 * tasks: 6 ints each {node (0-based), kind (0 diag strip, 1 below chunk, 2 fused small node), r0,
 * nrows, dest_begin, dest_count}; nodes: 8 ints each {m, n, sa, strip0, np, expect_f, expect_b,
 * pflag}; dest: global strip ids (forward: counters the task bumps, backward: flags it waits for);
 * expect: per strip, how many forward tasks add into its rows */
void spllt_b200_pipe_sizes(void *akeep, long long *out4);
/* multi-GPU (after spllt_b200_partition[_host]): spllt_b200_get_pipe returns the lists of the subtrees
 * this rank owns, these return the lists of the shared upper tree (same record layout) */
void spllt_b200_pipe_top_sizes(void *akeep, long long *sizes2);
void spllt_b200_get_pipe_top(void *akeep, int *tasks_f, int *tasks_b, int *expect);
/* path selection of the solve: share of L's entries in nodes wider than 256 columns, and the largest
 * nrhs served by the persistent kernels (0: level-set launches for every nrhs) */
double spllt_b200_wide_frac(void *akeep);
int spllt_b200_pipe_max_nrhs(void *akeep);
/* diagnostic: one un-graphed solve with 8 globaltimer stamps (ns) per claimed task of the persistent
 * kernels {start, waits satisfied, solved, end, 4 strip-internal}; out_f / out_b hold 8 * tasks * ceil(nrhs / rc) values
 * (rc = 1 for nrhs == 1, else 4), in claim order */
void spllt_b200_trace_solve(void *fkeep, int nrhs, double *d_x, int ldx, unsigned long long *out_f,
                            unsigned long long *out_b);
void spllt_b200_get_pipe(void *akeep, int *tasks_f, int *tasks_b, int *nodes, int *dest, int *expect);

/* ---- FP64 peak probes (no FP64 figure in MEASURED_PEAKS.json): enqueue a register-resident
 * DMMA (kind 0) or DFMA (kind 1) loop on every SM; returns the flops it will execute */
double spllt_b200_peak_probe(int kind, int iters, void *stream);

/* ---- multi-GPU (one process per GPU, all GPUs of one NVLink / NVSwitch box).
 * Subtrees of the assembly tree are mapped to ranks by proportional mapping; the upper tree is
 * distributed by block column (owner computes) and walked in the same step order on every rank.
 * Every rank maps every peer's factor arena (CUDA IPC): a subtree's contributions into the upper
 * tree are accumulated in its generated element in local HBM (src/spllt_kernels_mod.F90:780-821) and
 * scattered once into the owning ranks' HBM by an apply kernel (RED.ADD.F64 on peer addresses --
 * spllt_subtree_apply_buffer / spllt_scatter_block, src/spllt_factorization_mod.F90:39-191), and a
 * finished upper-tree block column is copied into the peers' arenas by its owner's SMs and
 * announced by a flag.  Usage on every rank r of `world`:
 *   spllt_analyse(...);  spllt_b200_partition(akeep, fkeep, r, world);
 *   spllt_b200_comm_export(fkeep, mine);   all-gather the 128-byte records (MPI / torch.distributed);
 *   spllt_b200_comm_attach(fkeep, r, world, table);
 *   spllt_factor / spllt_b200_factor_dev, spllt_wait -- as on one GPU, called by every rank. */
void *spllt_b200_arena_ptr(void *fkeep);            /* device pointer of the HBM arena        */
void spllt_b200_partition(void *akeep, void *fkeep, int rank, int world);
int spllt_b200_node_owner(void *akeep, int node);  /* node 1-based; -1 = upper tree (distributed by block column) */
/* host-only partition (no device state): used by the CPU multi-process tests */
void spllt_b200_partition_host(void *akeep, int rank, int world);
#define SPLLT_B200_HANDLE_BYTES 128
int spllt_b200_comm_export(void *fkeep, void *out128);                       /* 0 = ok */
int spllt_b200_comm_attach(void *fkeep, int rank, int world, const void *all_handles);
/* out[node] = number of inner panels (64 columns) this rank's schedule holds for each node */
void spllt_b200_panel_coverage(void *akeep, long long *out);
/* this rank's launch records (8 columns: kind depth begin count phase tag stream deadline; kind 0
 * panel, 1 / 2 tile updates, 3 push of global block column `begin` to the peers, 4 wait for it; phase 1 =
 * upper tree, depth = step index there) and tile tasks (10 columns: node i0 j0 k0 mt nt kk src off ld) */
long long spllt_b200_num_launch_records(void *akeep);
void spllt_b200_get_launch_records(void *akeep, long long *out);
long long spllt_b200_num_tile_tasks(void *akeep);
void spllt_b200_get_tile_tasks(void *akeep, long long *out);
/* upper-tree steps, identical on every rank (4 columns: node (0-based), local block column, owner, slot) */
int spllt_b200_num_top_steps(void *akeep);
void spllt_b200_get_top_steps(void *akeep, int *out);
void spllt_b200_get_bcol_owner(void *akeep, int *out);   /* [nbcol] */
int spllt_b200_dist_top(void *akeep);
/* `world` ranks emulated on ONE GPU inside one process (engines partitioned with rank 0..world-1):
 * the same per-rank programs, kernels and peer addressing, enqueued in an order in which no kernel
 * waits for a later one.  For tests on single-GPU machines.  Synchronises. */
int spllt_b200_emulate_ranks_factor(void **fkeeps, int world, const double *d_val);
/* {max |a - b|, max |b|} over the lower trapezoids of the nodes rank A holds: two factorizations of
 * the same matrix with the same ordering on the same device (distributed vs single-GPU) */
int spllt_b200_compare_factor(void *akeep_a, void *fkeep_a, void *akeep_b, void *fkeep_b, double *out2);
/* multi-GPU solve, one call per phase (all asynchronous on the stream).  The caller sums the work
 * vector (spllt_b200_xw_ptr: n x nrhs doubles, row-major) over the ranks after phases 1 and 4.
 * 0: permute the rhs in, drop the entries this rank does not own; 1: forward sweep of the rank's
 * subtrees; 2: forward sweep of the upper tree; 3: backward sweep of the upper tree; 4: backward
 * sweep of the rank's subtrees, drop foreign entries; 5: permute the solution out (d_x, ldx). */
void spllt_b200_solve_phase(void *fkeep, int nrhs, double *d_x, int ldx, int phase);
void *spllt_b200_xw_ptr(void *fkeep, int nrhs);

#ifdef __cplusplus
}
#endif
#endif
