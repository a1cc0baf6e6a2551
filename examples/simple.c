/* C-ABI smoke program: the 3x3 system of the reference's example/C/simple.c:25-52, written
 * against include/spllt_iface.h only (job = 0; the reference example passes 6, which the
 * current reference library rejects, src/spllt_solve_mod.F90:216-220).
 *   gcc -Iinclude examples/simple.c -Lspllt_b200 -lspllt_b200 -Wl,-rpath,$PWD/spllt_b200 -o simple
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <spllt_iface.h>

int main(void) {
  void *akeep = NULL, *fkeep = NULL;
  int n = 3, nnz = 5, nrhs = 1, nb = 4, stat;
  int ptr[4] = {1, 3, 5, 6}, row[5] = {1, 2, 2, 3, 3}, order[3];
  double val[5] = {2.0, -1.0, 2.0, -1.0, 2.0}, x[3] = {1, 1, 1}, rhs[3] = {1, 1, 1};
  long worksize;
  spllt_inform_t info;
  spllt_options_t options = SPLLT_OPTIONS_NULL();
  options.nb = nb;

  spllt_analyse(&akeep, &fkeep, &options, n, ptr, row, &info, order);
  spllt_factor(akeep, fkeep, &options, nnz, val, &info);
  spllt_wait();
  spllt_prepare_solve(akeep, fkeep, nb, nrhs, &worksize, &info);
  printf("Need a workspace of size %ld\n", worksize);
  double *y = calloc(n * nrhs, sizeof(double)), *w = calloc(worksize + 1, sizeof(double));
  spllt_set_mem_solve(akeep, fkeep, nb, nrhs, worksize, y, w, &info);
  spllt_solve(fkeep, &options, order, nrhs, x, &info, 0);
  spllt_wait();
  spllt_chkerr(n, ptr, row, val, nrhs, x, rhs);
  printf("x = %.15g %.15g %.15g (expected 1.5 2 1.5), flag %d\n", x[0], x[1], x[2], info.flag);
  spllt_deallocate_akeep(&akeep, &stat);
  spllt_deallocate_fkeep(&fkeep, &stat);
  free(y);
  free(w);
  return !(x[0] > 1.49999 && x[0] < 1.50001);
}
