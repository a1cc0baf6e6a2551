/* C ABI of the B200-native SpLLT numerical phase.
 *
 * Drop-in for the reference's C interface: every entry point, struct layout and calling
 * convention below is the one declared in the reference's include/spllt_iface.h:8-148 and
 * implemented by interfaces/C/spllt_data_ciface.F90:89-780, so a program written against the
 * reference (example/C/simple.c) links against libspllt_b200.so unchanged.
 *
 * Conventions kept from the reference:
 *   - ptr/row: lower triangle, CSC, 1-based (example/C/simple.c:38-39);
 *   - order is an output: order[i] = pivot position (1-based) of variable i+1;
 *   - x is n x nrhs column-major and is overwritten by the solution;
 *   - job 0 = forward + backward, 1 = forward only, 2 = backward only, anything else
 *     sets info->flag = -10 and does no work (src/spllt_solve_mod.F90:203-221);
 *   - spllt_factor and spllt_solve_worker are asynchronous; spllt_wait() completes them.
 * What differs: factors live in HBM (owned by fkeep); y / workspace passed to
 * spllt_set_mem_solve are only used to mirror the forward result of job 1.
 */
#ifndef SPLLT_IFACE_H
#define SPLLT_IFACE_H

#ifdef __cplusplus
extern "C" {
#endif

/* reference include/spllt_iface.h:8-12 */
typedef struct {
  void *akeep;
  void *fkeep;
  void *tm;
} spllt_data_t;

/* reference include/spllt_iface.h:14-31 (14 ints) */
typedef struct {
  int print_level;
  int nrhs;
  int ncpu;
  int nb;
  int nemin;
  int prune_tree;
  int min_width_blas;
  int nb_min;
  int nb_max;
  int nrhs_min;
  int nrhs_max;
  int nb_linear_comp;
  int nrhs_linear_comp;
  int chunk;
} spllt_options_t;

/* defaults of reference include/spllt_iface.h:33-47 */
#define SPLLT_OPTIONS_NULL()                                                                  \
  {                                                                                           \
    .print_level = 0, .nrhs = 1, .ncpu = 1, .nb = 16, .nemin = 32, .prune_tree = 1,           \
    .min_width_blas = 8, .nb_min = 32, .nb_max = 32, .nrhs_min = 1, .nrhs_max = 1,            \
    .nb_linear_comp = 0, .nrhs_linear_comp = 0, .chunk = 10                                   \
  }

/* reference include/spllt_iface.h:49-57.  num_factor / num_flops are C ints in the
 * reference and overflow for large problems; they saturate at INT_MAX here and the exact
 * values are available through spllt_b200_num_factor / spllt_b200_num_flops. */
typedef struct {
  int flag;
  int maxdepth;
  int num_factor;
  int num_flops;
  int num_nodes;
  int stat;
} spllt_inform_t;

/* error codes, src/spllt_data_mod.F90:31-35 (+ one addition for a failed pivot, which the
 * reference silently drops at src/spllt_kernels_mod.F90:1179-1181) */
#define SPLLT_SUCCESS 0
#define SPLLT_ERROR_ALLOCATION (-1)
#define SPLLT_WARNING_PARAM_VALUE (-10)
#define SPLLT_ERROR_NOT_POS_DEF (-20)
#define SPLLT_ERROR_UNIMPLEMENTED (-98)
#define SPLLT_ERROR_UNKNOWN (-99)

/* interfaces/C/spllt_data_ciface.F90:137 */
void spllt_analyse(void **akeep, void **fkeep, spllt_options_t *options, int n, int *ptr, int *row,
                   spllt_inform_t *info, int *order);
/* :194 */
void spllt_factor(void *akeep, void *fkeep, spllt_options_t *options, int nnz, double *val,
                  spllt_inform_t *info);
/* :238 */
void spllt_prepare_solve(void *akeep, void *fkeep, int nb, int nrhs, long *worksize, spllt_inform_t *info);
/* :292 */
void spllt_set_mem_solve(void *akeep, void *fkeep, int nb, int nrhs, long worksize, double *y,
                         double *workspace, spllt_inform_t *info);
/* declared in the reference header (:91-94); Fortran body commented out (:343-369) */
void spllt_solve_workspace_size(void *fkeep, int nworker, int nrhs, long *size);
/* :372 */
void spllt_solve(void *fkeep, spllt_options_t *options, int *order, int nrhs, double *x,
                 spllt_inform_t *info, int job);
/* :431 */
void spllt_solve_worker(void *fkeep, spllt_options_t *options, int *order, int nrhs, double *x,
                        spllt_inform_t *info, int job, double *workspace, long worksize, void *tm);
/* :490 */
void spllt_wait(void);
/* :499 */
void spllt_chkerr(int n, int *ptr, int *row, double *val, int nrhs, double *x, double *rhs);
/* :555, :587 */
void spllt_deallocate_fkeep(void **fkeep, int *stat);
void spllt_deallocate_akeep(void **akeep, int *stat);
/* :89, :118 */
void spllt_task_manager_deallocate(void **task_manager, int *stat);
void spllt_task_manager_init(void **task_manager);
/* :656 */
void spllt_all(void **akeep, void **fkeep, spllt_options_t *options, int n, int nnz, int nrhs, int nb,
               int *ptr, int *row, double *val, double *x, double *rhs, spllt_inform_t *info);

#ifdef __cplusplus
}
#endif
#endif
