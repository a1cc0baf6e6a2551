// extern "C" boundary: the reference's C ABI (include/spllt_iface.h, implemented in the
// reference by interfaces/C/spllt_data_ciface.F90:89-780) plus the B200 additions of
// include/spllt_b200.h.  No torch types, no exceptions across the boundary.
#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/spllt_b200.h"
#include "cuda_check.h"
#include "engine.h"

using namespace spllt;

namespace {

struct AKeep {
  std::shared_ptr<Analysis> A;
};
struct FKeep {
  Engine eng;
};

std::mutex g_mu;
std::vector<Engine*> g_live;  // spllt_wait() has no handle argument: it drains every engine

void reg(Engine* e) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_live.push_back(e);
}
void unreg(Engine* e) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_live.erase(std::remove(g_live.begin(), g_live.end(), e), g_live.end());
}

int sat(i64 v) { return v > INT_MAX ? INT_MAX : (int)v; }

void fill_info(const Analysis& A, spllt_inform_t* info, int flag) {
  if (!info) return;
  info->flag = flag;
  info->maxdepth = A.ndepth;
  info->num_factor = sat(A.num_factor);
  info->num_flops = sat(A.num_flops);
  info->num_nodes = A.nnodes;
  info->stat = 0;
}

Analysis* AA(void* akeep) { return akeep ? ((AKeep*)akeep)->A.get() : nullptr; }
Engine* EE(void* fkeep) { return fkeep ? &((FKeep*)fkeep)->eng : nullptr; }

int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

// Runs `f`; a CUDA failure / missing device / host allocation failure inside it becomes the
// reference's error code in info%flag (src/spllt_data_mod.F90:31-35) instead of leaving the C ABI
// as an exception or aborting the caller's process.  Returns the flag (0 = no failure).
template <class F>
int guarded(spllt_inform_t* info, F&& f) {
  int flag = SPLLT_SUCCESS, stat = 0;
  try {
    f();
    return SPLLT_SUCCESS;
  } catch (const CudaFailure& e) {
    flag = e.oom() ? SPLLT_ERROR_ALLOCATION : SPLLT_ERROR_UNKNOWN;
    stat = (int)e.err;
  } catch (const NoDevice&) {
    flag = SPLLT_ERROR_UNKNOWN;
  } catch (const std::bad_alloc&) {
    fprintf(stderr, "spllt_b200: host allocation failed\n");
    flag = SPLLT_ERROR_ALLOCATION;
  }
  if (info) {
    info->flag = flag;
    info->stat = stat;
  }
  return flag;
}

// Pivot status of the last factorization once it has completed (the stream is idle): the
// reference leaves info%flag of the asynchronous spllt_factor untouched; here the first call on
// the handle after spllt_wait() reports SPLLT_ERROR_NOT_POS_DEF (-20).
int post_flag(Engine* e) {
  if (!e || !e->uploaded || !e->factored || !e->stream) return SPLLT_SUCCESS;
  if (cudaStreamQuery(e->stream) != cudaSuccess) {
    cudaGetLastError();
    return SPLLT_SUCCESS;   // still running: unknown yet
  }
  return e->pivot_flag() ? SPLLT_ERROR_NOT_POS_DEF : SPLLT_SUCCESS;
}

void analyse_impl(void** akeep, void** fkeep, spllt_options_t* options, int n, const int* ptr, const int* row,
                  spllt_inform_t* info, int* order, int ordering) {
  if (!akeep || !fkeep || !options) {
    fprintf(stderr, "Error, akeep/fkeep/options handle is NULL\n");
    return;
  }
  if (!ptr) fprintf(stderr, "Error, ptr provided by the user is empty\n");
  if (!row) fprintf(stderr, "Error, row provided by the user is empty\n");
  if (!ptr || !row) {
    if (info) info->flag = SPLLT_ERROR_UNKNOWN;
    return;
  }
  // handles are allocated on first use (interfaces/C/spllt_data_ciface.F90:151-163)
  if (!*akeep) *akeep = new AKeep();
  if (!*fkeep) {
    FKeep* f = new FKeep();
    reg(&f->eng);
    *fkeep = f;
  }
  AKeep* ak = (AKeep*)*akeep;
  FKeep* fk = (FKeep*)*fkeep;
  fk->eng.release();
  ak->A = std::make_shared<Analysis>();
  Analysis& A = *ak->A;
  // SPLLT_B200_ANALYSE_TIMING=1: phase timers of the host analysis on stderr (the reference's
  // timer_mod, src/timer_mod.F90:76-547, reduced to what the host still does)
  const bool timing = env_int("SPLLT_B200_ANALYSE_TIMING", 0) != 0;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
  int rc = build_analysis(n, ptr, row, options->nb, options->nemin, options->ncpu, options->prune_tree != 0, ordering,
                          ordering == ORDER_USER ? order : nullptr, A);
  const double t1 = now();
  A.min_width_blas = options->min_width_blas;
  if (rc == 0 && n > 0) {
    build_factor_schedule(A, env_int("SPLLT_B200_TILE_L_MIN", 128));
    const double t2 = now();
    build_solve_schedule(A);
    if (timing)
      fprintf(stderr, "spllt_b200 analyse: ordering + symbolic + tables %.3f s, factor schedule %.3f s, solve schedule %.3f s\n",
              t1 - t0, t2 - t1, now() - t2);
    if (order)
      for (int i = 0; i < n; ++i) order[i] = A.sym.order[i];
  }
  fk->eng.A = ak->A;
  fk->eng.use_graph = env_int("SPLLT_B200_GRAPH", 1) != 0;
  fill_info(A, info, rc == 0 ? SPLLT_SUCCESS : SPLLT_ERROR_UNKNOWN);
}

long worksize_of(const Analysis& A, int nrhs) {
  // get_solve_blocks: worksize = sum over nodes of the L2 row blocks' blkm * nrhs
  // (src/spllt_solve_dep_mod.F90:1990-1996) = sum (m - n) * nrhs
  i64 s = 0;
  for (const HNode& nd : A.nodes) s += (i64)(nd.m - nd.n) * nrhs;
  return (long)s;
}

// mirror of the forward result into the caller's y in the reference's layout:
// L1 row blocks in node order, each blkm x nrhs column-major (sblock_assoc_mem,
// src/spllt_solve_dep_mod.F90:2033-2143)
void mirror_y(Engine& e, int nrhs) {
  if (!e.host_y) return;
  const Analysis& A = *e.A;
  std::vector<double> xw((size_t)A.n * nrhs);
  e.get_fwd(nrhs, xw.data());
  i64 off = 0;
  for (const HNode& nd : A.nodes)
    for (int r0 = 0; r0 < nd.n; r0 += A.nb) {
      int bm = std::min(A.nb, nd.n - r0);
      for (int r = 0; r < nrhs; ++r)
        for (int i = 0; i < bm; ++i) e.host_y[off + i + (i64)r * bm] = xw[(size_t)(nd.sa + r0 + i) * nrhs + r];
      off += (i64)bm * nrhs;
    }
}

void solve_impl(void* fkeep, int nrhs, double* x, spllt_inform_t* info, int job, bool wait) {
  Engine* e = EE(fkeep);
  if (!e || !e->A) {
    fprintf(stderr, "Error, fkeep provided by the user is empty\n");
    return;
  }
  if (!x) {
    fprintf(stderr, "Error, x provided by the user is empty\n");
    return;
  }
  if (job < 0 || job > 2) {
    // src/spllt_solve_mod.F90:216-220
    fprintf(stderr, "Warning, job = %d is not a valid value (0, 1 or 2): nothing done\n", job);
    if (info) fill_info(*e->A, info, SPLLT_WARNING_PARAM_VALUE);
    return;
  }
  if (info) fill_info(*e->A, info, SPLLT_SUCCESS);
  guarded(info, [&] {
    e->solve_host(x, nrhs, job);
    if (wait) {
      e->sync();
      if (job == 1) mirror_y(*e, nrhs);
      if (e->pivot_flag() && info) info->flag = SPLLT_ERROR_NOT_POS_DEF;
    }
  });
}

}  // namespace

// =========================================================================== reference ABI
extern "C" {

void spllt_analyse(void** akeep, void** fkeep, spllt_options_t* options, int n, int* ptr, int* row,
                   spllt_inform_t* info, int* order) {
  guarded(info, [&] { analyse_impl(akeep, fkeep, options, n, ptr, row, info, order, ORDER_METIS); });
}

void spllt_b200_analyse(void** akeep, void** fkeep, spllt_options_t* options, int n, const int* ptr, const int* row,
                        spllt_inform_t* info, int* order, int ordering) {
  guarded(info, [&] { analyse_impl(akeep, fkeep, options, n, ptr, row, info, order, ordering); });
}

void spllt_factor(void* akeep, void* fkeep, spllt_options_t* options, int nnz, double* val, spllt_inform_t* info) {
  (void)options;
  Analysis* A = AA(akeep);
  Engine* e = EE(fkeep);
  if (!A) fprintf(stderr, "Error, akeep provided by the user is empty\n");
  if (!e) fprintf(stderr, "Error, fkeep provided by the user is empty\n");
  if (!val) fprintf(stderr, "Error, val provided by the user is empty\n");
  if (!A || !e || !val) return;
  if ((i64)nnz != A->nnz) fprintf(stderr, "Warning, nnz = %d differs from ptr(n+1)-1 = %lld\n", nnz, (long long)A->nnz);
  fill_info(*A, info, SPLLT_SUCCESS);
  // asynchronous like the reference: a pivot failure of THIS factorization is only known after
  // spllt_wait(); it is reported (flag -20) by the first call on the handle after the wait, and
  // spllt_b200_pivot_flag / spllt_b200_factor_status read it directly.
  guarded(info, [&] { e->factor_host(val); });
}

void spllt_prepare_solve(void* akeep, void* fkeep, int nb, int nrhs, long* worksize, spllt_inform_t* info) {
  Analysis* A = AA(akeep);
  Engine* e = EE(fkeep);
  if (!A || !e) {
    fprintf(stderr, "Error, akeep/fkeep provided by the user is empty\n");
    return;
  }
  int flag = SPLLT_SUCCESS;
  if (nb != A->nb) {
    // solve tiles index the factor's block columns by position (src/spllt_solve_dep_mod.F90:1950,2021)
    fprintf(stderr, "Warning, solve nb = %d differs from the analyse nb = %d; using %d\n", nb, A->nb, A->nb);
    flag = SPLLT_WARNING_PARAM_VALUE;
  }
  e->prep_nb = A->nb;
  e->prep_nrhs = nrhs;
  if (worksize) *worksize = worksize_of(*A, nrhs);
  fill_info(*A, info, flag);
  guarded(info, [&] {
    if (post_flag(e) && info) info->flag = SPLLT_ERROR_NOT_POS_DEF;
  });
}

void spllt_set_mem_solve(void* akeep, void* fkeep, int nb, int nrhs, long worksize, double* y, double* workspace,
                         spllt_inform_t* info) {
  (void)nb;
  (void)nrhs;
  (void)worksize;
  (void)workspace;  // update vectors live in HBM; the host workspace is not needed
  Analysis* A = AA(akeep);
  Engine* e = EE(fkeep);
  if (!A || !e) {
    fprintf(stderr, "Error, akeep/fkeep provided by the user is empty\n");
    return;
  }
  e->host_y = y;
  fill_info(*A, info, SPLLT_SUCCESS);
  guarded(info, [&] {
    if (post_flag(e) && info) info->flag = SPLLT_ERROR_NOT_POS_DEF;
  });
}

void spllt_solve_workspace_size(void* fkeep, int nworker, int nrhs, long* size) {
  Engine* e = EE(fkeep);
  if (!e || !e->A || !size) return;
  // src/spllt_data_mod.F90:655
  i64 n = e->A->n;
  *size = (long)(n * nrhs + ((i64)e->A->maxmn + n) * nrhs * nworker);
}

void spllt_solve(void* fkeep, spllt_options_t* options, int* order, int nrhs, double* x, spllt_inform_t* info,
                 int job) {
  (void)options;
  (void)order;
  solve_impl(fkeep, nrhs, x, info, job, true);
}

void spllt_solve_worker(void* fkeep, spllt_options_t* options, int* order, int nrhs, double* x, spllt_inform_t* info,
                        int job, double* workspace, long worksize, void* tm) {
  (void)options;
  (void)order;
  (void)workspace;
  (void)worksize;
  (void)tm;
  solve_impl(fkeep, nrhs, x, info, job, false);
}

void spllt_wait(void) {
  std::vector<Engine*> live;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    live = g_live;
  }
  guarded(nullptr, [&] {
    for (Engine* e : live)
      if (e->uploaded) e->sync();
  });
}

// src/utils_mod.F90:432-478 (host-side acceptance metric; prints like the reference)
static int chkerr_impl(int n, const int* ptr, const int* row, const double* val, int nrhs, const double* x,
                       const double* rhs, double* err_out, bool print) {
  std::vector<double> ax((size_t)n * nrhs, 0.0);
  double amax = 0.0;
  for (int j = 0; j < n; ++j)
    for (int e = ptr[j] - 1; e < ptr[j + 1] - 1; ++e) {
      int i = row[e] - 1;
      double a = val[e];
      amax = std::max(amax, std::fabs(a));
      for (int r = 0; r < nrhs; ++r) {
        ax[i + (size_t)r * n] += a * x[j + (size_t)r * n];
        if (i != j) ax[j + (size_t)r * n] += a * x[i + (size_t)r * n];
      }
    }
  int ok = 0;
  for (int r = 0; r < nrhs; ++r) {
    double nr = 0, nb = 0, nx = 0;
    for (int i = 0; i < n; ++i) {
      double d = rhs[i + (size_t)r * n] - ax[i + (size_t)r * n];
      nr += d * d;
      nb += rhs[i + (size_t)r * n] * rhs[i + (size_t)r * n];
      nx += x[i + (size_t)r * n] * x[i + (size_t)r * n];
    }
    double err = std::sqrt(nr) / (std::sqrt(nb) + amax * std::sqrt(nx));
    if (err_out) err_out[r] = err;
    if (err != err) {
      if (print) printf("Backward error of rhs %3d is equal to a NAN\n", r + 1);
    } else if (err > 1e-14) {
      if (print) fprintf(stderr, "Wrong Bwd error for %4d/%4d : %10.2E\n", r + 1, nrhs, err);
    } else {
      if (print) fprintf(stderr, "Bwd error for %4d/%4d : %10.2E\n", r + 1, nrhs, err);
      ++ok;
    }
  }
  if (print) fprintf(stderr, "Backward error... ok for %3d/%3d\n", ok, nrhs);
  return ok;
}

void spllt_chkerr(int n, int* ptr, int* row, double* val, int nrhs, double* x, double* rhs) {
  if (!ptr) fprintf(stderr, "Error, ptr provided by the user is empty\n");
  if (!row) fprintf(stderr, "Error, row provided by the user is empty\n");
  if (!val) fprintf(stderr, "Error, val provided by the user is empty\n");
  if (!x) fprintf(stderr, "Error, x provided by the user is empty\n");
  if (!rhs) fprintf(stderr, "Error, rhs provided by the user is empty\n");
  if (!ptr || !row || !val || !x || !rhs) return;
  chkerr_impl(n, ptr, row, val, nrhs, x, rhs, nullptr, true);
}

int spllt_b200_chkerr(int n, const int* ptr, const int* row, const double* val, int nrhs, const double* x,
                      const double* rhs, double* err) {
  return chkerr_impl(n, ptr, row, val, nrhs, x, rhs, err, false);
}

void spllt_deallocate_fkeep(void** fkeep, int* stat) {
  if (stat) *stat = 0;
  if (!fkeep || !*fkeep) return;
  FKeep* f = (FKeep*)*fkeep;
  unreg(&f->eng);
  delete f;
  *fkeep = nullptr;
}

void spllt_deallocate_akeep(void** akeep, int* stat) {
  if (stat) *stat = 0;
  if (!akeep || !*akeep) return;
  delete (AKeep*)*akeep;
  *akeep = nullptr;
}

// The task manager is an OpenMP artefact (src/task_manager_omp.F90); the handle is kept so
// that callers of the reference ABI keep working.
void spllt_task_manager_init(void** task_manager) {
  if (task_manager && !*task_manager) *task_manager = calloc(1, 16);
}
void spllt_task_manager_deallocate(void** task_manager, int* stat) {
  if (stat) *stat = 0;
  if (task_manager && *task_manager) {
    free(*task_manager);
    *task_manager = nullptr;
  }
}

void spllt_all(void** akeep, void** fkeep, spllt_options_t* options, int n, int nnz, int nrhs, int nb, int* ptr,
               int* row, double* val, double* x, double* rhs, spllt_inform_t* info) {
  if (!rhs) fprintf(stderr, "Error, rhs provided by the user is empty\n");
  if (!x) fprintf(stderr, "Error, x/rhs provided by the user is empty\n");
  std::vector<int> order(std::max(n, 1));
  spllt_analyse(akeep, fkeep, options, n, ptr, row, info, order.data());
  spllt_factor(*akeep, *fkeep, options, nnz, val, info);
  spllt_wait();
  long worksize = 0;
  spllt_prepare_solve(*akeep, *fkeep, nb, nrhs, &worksize, info);
  spllt_solve_worker(*fkeep, options, order.data(), nrhs, x, info, 0, nullptr, worksize, nullptr);
  spllt_wait();
  if (ptr && row && val && x && rhs) chkerr_impl(n, ptr, row, val, nrhs, x, rhs, nullptr, true);
}

// =========================================================================== B200 additions
long long spllt_b200_num_factor(void* akeep) { return AA(akeep) ? AA(akeep)->num_factor : 0; }
long long spllt_b200_num_flops(void* akeep) { return AA(akeep) ? AA(akeep)->num_flops : 0; }
long long spllt_b200_arena_doubles(void* akeep) { return AA(akeep) ? AA(akeep)->arena : 0; }
int spllt_b200_num_nodes(void* akeep) { return AA(akeep) ? AA(akeep)->nnodes : 0; }
int spllt_b200_num_bcol(void* akeep) { return AA(akeep) ? AA(akeep)->nbcol : 0; }
long long spllt_b200_final_blk(void* akeep) { return AA(akeep) ? AA(akeep)->final_blk : 0; }
int spllt_b200_maxmn(void* akeep) { return AA(akeep) ? AA(akeep)->maxmn : 0; }
int spllt_b200_num_depth(void* akeep) { return AA(akeep) ? AA(akeep)->ndepth : 0; }

long long spllt_b200_rlist_len(void* akeep) { return AA(akeep) ? (long long)AA(akeep)->sym.rlist.size() : 0; }
void spllt_b200_get_symbolic(void* akeep, int* sptr, int* sparent, long long* rptr, int* rlist) {
  const Symbolic& S = AA(akeep)->sym;
  std::copy(S.sptr.begin(), S.sptr.end(), sptr);
  std::copy(S.sparent.begin(), S.sparent.end(), sparent);
  std::copy(S.rptr.begin(), S.rptr.end(), rptr);
  std::copy(S.rlist.begin(), S.rlist.end(), rlist);
}
void spllt_b200_get_blocks(void* akeep, long long* out) {
  std::vector<RefBlock> b;
  ref_blocks(*AA(akeep), b);
  for (size_t i = 0; i < b.size(); ++i) {
    long long* r = out + 9 * i;
    r[0] = b[i].id; r[1] = b[i].blkm; r[2] = b[i].blkn; r[3] = b[i].sa; r[4] = b[i].dblk;
    r[5] = b[i].last_blk; r[6] = b[i].node; r[7] = b[i].bcol; r[8] = b[i].dep_initial;
  }
}
void spllt_b200_get_nodes(void* akeep, long long* out) {
  const Analysis& A = *AA(akeep);
  for (int s = 0; s < A.nnodes; ++s) {
    const HNode& nd = A.nodes[s];
    long long* r = out + 8 * s;
    i64 ntile = (i64)nd.nc * nd.nr - (i64)nd.nc * (nd.nc - 1) / 2;
    r[0] = nd.sa + 1; r[1] = nd.en + 1; r[2] = nd.parent < 0 ? A.nnodes + 1 : nd.parent + 1; r[3] = nd.nchild;
    r[4] = nd.least_desc + 1; r[5] = A.nb; r[6] = nd.blk0 + 1; r[7] = nd.blk0 + ntile;
  }
}
void spllt_b200_get_small(void* akeep, int* out) {
  const Analysis& A = *AA(akeep);
  for (int s = 0; s < A.nnodes; ++s) out[s] = A.nodes[s].small;
}
void spllt_b200_get_weight(void* akeep, long long* out) {
  const Analysis& A = *AA(akeep);
  for (int s = 0; s <= A.nnodes; ++s) out[s] = A.weight[s];
}
long long spllt_b200_lmap_len(void* akeep, int bcol) {
  const Analysis& A = *AA(akeep);
  return A.lmap_ptr[bcol] - A.lmap_ptr[bcol - 1];
}
void spllt_b200_get_lmap(void* akeep, int bcol, long long* dst, long long* src) {
  const Analysis& A = *AA(akeep);
  int g = bcol - 1;
  const HNode& nd = A.nodes[A.bcol_node[g]];
  int r0 = A.bcol_c[g] * A.nb;
  int w = std::min(A.nb, nd.n - r0);
  i64 k = 0;
  for (i64 e = A.lmap_ptr[g]; e < A.lmap_ptr[g + 1]; ++e, ++k) {
    dst[k] = 1 + (i64)(A.lmap_row[e] - r0) * w + (A.lmap_col[e] - r0);
    src[k] = A.lmap_src[e] + 1;
  }
}

// closed form of get_solve_blocks (src/spllt_solve_dep_mod.F90:1861-2030)
static void sblocks(const Analysis& A, int nb, std::vector<int>* out) {
  int id = 0, bcol = 0;
  for (int s = 0; s < A.nnodes; ++s) {
    const HNode& nd = A.nodes[s];
    int l1 = (nd.n + nb - 1) / nb, l2 = (nd.m - nd.n + nb - 1) / nb, nrow = l1 + l2;
    for (int j = 0; j < l1; ++j) {
      int bn = std::min(nb, nd.n - j * nb);
      int dblk = id + 1, last = dblk + nrow - 1 - j, sa = 1;
      for (int k = j; k < nrow; ++k) {
        int bm = k < l1 ? std::min(nb, nd.n - k * nb) : std::min(nb, nd.m - nd.n - (k - l1) * nb);
        ++id;
        if (out) {
          int row[9] = {id, bm, bn, sa, dblk, last, bcol + j + 1, s + 1, bm};
          out->insert(out->end(), row, row + 9);
        }
        sa += bm * bn;
      }
    }
    bcol += l1;
  }
  if (!out) return;
}
int spllt_b200_num_sblocks(void* akeep, int nb) {
  const Analysis& A = *AA(akeep);
  int cnt = 0;
  for (const HNode& nd : A.nodes) {
    int l1 = (nd.n + nb - 1) / nb, l2 = (nd.m - nd.n + nb - 1) / nb;
    cnt += l1 * (l1 + 1) / 2 + l1 * l2;
  }
  return cnt;
}
void spllt_b200_get_sblocks(void* akeep, int nb, int* out) {
  std::vector<int> v;
  sblocks(*AA(akeep), nb, &v);
  std::copy(v.begin(), v.end(), out);
}

long long spllt_b200_lcol_size(void* akeep, int bcol) {
  const Analysis& A = *AA(akeep);
  const HNode& nd = A.nodes[A.bcol_node[bcol - 1]];
  int r0 = A.bcol_c[bcol - 1] * A.nb;
  return (long long)(nd.m - r0) * std::min(A.nb, nd.n - r0);
}
void spllt_b200_get_lcol(void* fkeep, int bcol, double* out) {
  guarded(nullptr, [&] { EE(fkeep)->get_lcol(bcol - 1, out); });
}
long long spllt_b200_factor_size(void* akeep) {
  const Analysis& A = *AA(akeep);
  long long s = 0;
  for (int g = 1; g <= A.nbcol; ++g) s += spllt_b200_lcol_size(akeep, g);
  return s;
}
void spllt_b200_get_factor(void* fkeep, double* out) {
  Engine* e = EE(fkeep);
  const Analysis& A = *e->A;
  guarded(nullptr, [&] {
    i64 s = 0;
    for (int g = 0; g < A.nbcol; ++g) {
      const HNode& nd = A.nodes[A.bcol_node[g]];
      int r0 = A.bcol_c[g] * A.nb;
      e->get_lcol(g, out + s);
      s += (i64)(nd.m - r0) * std::min(A.nb, nd.n - r0);
    }
  });
}

void spllt_b200_set_stream(void* fkeep, void* stream) {
  Engine* e = EE(fkeep);
  guarded(nullptr, [&] {
    e->upload_tables();
    e->stream = stream ? (cudaStream_t)stream : e->own;
  });
}
void spllt_b200_factor_dev(void* akeep, void* fkeep, const double* d_val, spllt_inform_t* info) {
  if (info) fill_info(*AA(akeep), info, SPLLT_SUCCESS);
  guarded(info, [&] { EE(fkeep)->factor(d_val); });
}
void spllt_b200_solve_dev(void* fkeep, int nrhs, double* d_x, int ldx, int job, spllt_inform_t* info) {
  Engine* e = EE(fkeep);
  if (job < 0 || job > 2) {
    if (info) fill_info(*e->A, info, SPLLT_WARNING_PARAM_VALUE);
    return;
  }
  if (info) fill_info(*e->A, info, SPLLT_SUCCESS);
  guarded(info, [&] { e->solve(d_x, ldx, nrhs, job); });
}
void spllt_b200_get_fwd(void* fkeep, int nrhs, double* out) {
  guarded(nullptr, [&] { EE(fkeep)->get_fwd(nrhs, out); });
}
int spllt_b200_pivot_flag(void* fkeep) {
  int v = 0;
  int rc = guarded(nullptr, [&] { v = EE(fkeep)->pivot_flag(); });
  return rc ? rc : v;
}

long long spllt_b200_factor_launches(void* fkeep) {
  const Analysis& A = *EE(fkeep)->A;
  // + epoch counter + assemble + inversion of the diagonal blocks (the memsets are not kernels of ours);
  // multi-GPU: + two rank barriers + one apply kernel per generated element of this rank
  return (long long)A.launches.size() + 3 + (A.world > 1 ? 2 + (long long)A.gen.size() : 0);
}
long long spllt_b200_solve_launches(void* fkeep, int job) {
  // kernels of one solve with a single right-hand side (the pipelined path when it is enabled)
  const Analysis& A = *EE(fkeep)->A;
  const bool pipe = EE(fkeep)->use_pipe(1);
  long long per = 0;
  for (const SolveLaunch& L : (pipe ? A.slaunch : A.slaunch_full)) per += (L.diag_count > 0) + (L.upd_count > 0);
  long long tot = 0;   // level-set launches + persistent kernel + permutation
  if (job == 0 || job == 1) tot += per + pipe + 1;
  if (job == 0 || job == 2) tot += per + pipe + 1;
  return tot;
}
double spllt_b200_tile_flops(void* akeep) { return AA(akeep)->tile_flops; }
double spllt_b200_tile_flops_algo(void* akeep) { return AA(akeep)->tile_flops_algo; }
void spllt_b200_launch_breakdown(void* akeep, long long* out4) {
  out4[0] = out4[1] = out4[2] = out4[3] = 0;
  for (const Launch& L : AA(akeep)->launches)
    if (L.kind < 3) out4[L.kind]++;
  out4[3] = (long long)AA(akeep)->launches.size();
}

void spllt_b200_profile_factor(void* fkeep, const double* d_val, double* ms4, const char* csv) {
  guarded(nullptr, [&] { EE(fkeep)->profile_factor(d_val, ms4, csv); });
}

void spllt_b200_profile_solve(void* fkeep, int nrhs, double* d_x, int ldx, double* ms6, const char* csv) {
  guarded(nullptr, [&] { EE(fkeep)->profile_solve(d_x, ldx, nrhs, ms6, csv); });
}

void spllt_b200_trace_solve(void* fkeep, int nrhs, double* d_x, int ldx, unsigned long long* out_f,
                            unsigned long long* out_b) {
  guarded(nullptr, [&] { EE(fkeep)->trace_solve(d_x, ldx, nrhs, out_f, out_b); });
}
// multi-GPU: the upper-tree lists (sizes2 = {forward tasks, backward tasks}); same record layout
void spllt_b200_pipe_top_sizes(void* akeep, long long* sizes2) {
  const Analysis& A = *AA(akeep);
  sizes2[0] = (long long)A.ptasks_ft.size();
  sizes2[1] = (long long)A.ptasks_bt.size();
}
void spllt_b200_get_pipe_top(void* akeep, int* tasks_f, int* tasks_b, int* expect) {
  const Analysis& A = *AA(akeep);
  auto put = [](const std::vector<PTask>& v, int* o) {
    for (const PTask& t : v) {
      *o++ = t.node; *o++ = t.kind; *o++ = t.r0; *o++ = t.nrows; *o++ = t.dest_begin; *o++ = t.dest_count;
    }
  };
  put(A.ptasks_ft, tasks_f);
  put(A.ptasks_bt, tasks_b);
  for (int e : A.pexpect_top) *expect++ = e;
}
double spllt_b200_wide_frac(void* akeep) { return AA(akeep)->wide_frac; }
int spllt_b200_pipe_max_nrhs(void* akeep) { return AA(akeep)->pipe_max_nrhs; }
void spllt_b200_pipe_sizes(void* akeep, long long* out4) {
  const Analysis& A = *AA(akeep);
  out4[0] = (long long)A.ptasks_f.size();
  out4[1] = (long long)A.ptasks_b.size();
  out4[2] = A.nstrips;
  out4[3] = (long long)A.pipe_dest.size();
}
void spllt_b200_get_pipe(void* akeep, int* tasks_f, int* tasks_b, int* nodes, int* dest, int* expect) {
  const Analysis& A = *AA(akeep);
  auto put = [](const std::vector<PTask>& v, int* o) {
    for (const PTask& t : v) {
      *o++ = t.node; *o++ = t.kind; *o++ = t.r0; *o++ = t.nrows; *o++ = t.dest_begin; *o++ = t.dest_count;
    }
  };
  put(A.ptasks_f, tasks_f);
  put(A.ptasks_b, tasks_b);
  for (const PNode& p : A.pnodes) {
    *nodes++ = p.m; *nodes++ = p.n; *nodes++ = p.sa; *nodes++ = p.strip0;
    *nodes++ = p.np; *nodes++ = p.expect_f; *nodes++ = p.expect_b; *nodes++ = p.pflag;
  }
  for (int d : A.pipe_dest) *dest++ = d;
  for (int e : A.pexpect) *expect++ = e;
}
double spllt_b200_peak_probe(int kind, int iters, void* stream) {
  double fl = 0;
  guarded(nullptr, [&] {
    require_gpu();
    if (kind >= 10) fl = launch_dmma_warps(iters, kind - 10, (cudaStream_t)stream);
    else fl = kind == 0 ? launch_dmma_peak(iters, (cudaStream_t)stream) : launch_dfma_peak(iters, (cudaStream_t)stream);
  });
  return fl;
}

void* spllt_b200_arena_ptr(void* fkeep) {
  Engine* e = EE(fkeep);
  void* p = nullptr;
  guarded(nullptr, [&] {
    e->upload_tables();
    p = e->arena;
  });
  return p;
}
void spllt_b200_partition(void* akeep, void* fkeep, int rank, int world) {
  Analysis* A = AA(akeep);
  Engine* e = EE(fkeep);
  e->release();   // work lists change: upload again on next use
  partition_tree(*A, rank, world);
  build_factor_schedule(*A, env_int("SPLLT_B200_TILE_L_MIN", 128));
  build_solve_schedule(*A);
}
int spllt_b200_node_owner(void* akeep, int node) { return AA(akeep)->nodes[node - 1].owner; }
// host-only variant of spllt_b200_partition (no device state touched): for CPU tests of the mapping
void spllt_b200_partition_host(void* akeep, int rank, int world) {
  Analysis* A = AA(akeep);
  partition_tree(*A, rank, world);
  build_factor_schedule(*A, env_int("SPLLT_B200_TILE_L_MIN", 128));
  build_solve_schedule(*A);
}
// out[node] = number of inner panels this rank's schedule holds for that node
void spllt_b200_panel_coverage(void* akeep, long long* out) {
  const Analysis& A = *AA(akeep);
  for (int k = 0; k < A.nnodes; ++k) out[k] = 0;
  for (const PanelTask& t : A.panel_tasks)
    if (t.store) out[A.col2node[t.col0]]++;
}
long long spllt_b200_num_launch_records(void* akeep) { return (long long)AA(akeep)->launches.size(); }
// 8 columns per record: kind depth begin count phase tag stream deadline
void spllt_b200_get_launch_records(void* akeep, long long* out) {
  const Analysis& A = *AA(akeep);
  for (size_t i = 0; i < A.launches.size(); ++i) {
    const Launch& L = A.launches[i];
    long long* r = out + 8 * i;
    r[0] = L.kind; r[1] = L.depth; r[2] = L.begin; r[3] = L.count; r[4] = L.phase; r[5] = L.tag; r[6] = L.stream;
    r[7] = L.deadline;
  }
}
// tile tasks (10 columns): node i0 j0 k0 mt nt kk src off ld -- for host-side replays of the schedule
long long spllt_b200_num_tile_tasks(void* akeep) { return (long long)AA(akeep)->tile_tasks.size(); }
void spllt_b200_get_tile_tasks(void* akeep, long long* out) {
  const Analysis& A = *AA(akeep);
  for (size_t i = 0; i < A.tile_tasks.size(); ++i) {
    const TileTask& t = A.tile_tasks[i];
    long long* r = out + 10 * i;
    r[0] = t.node; r[1] = t.i0; r[2] = t.j0; r[3] = t.k0; r[4] = t.mt; r[5] = t.nt; r[6] = t.kk; r[7] = t.src;
    r[8] = t.off; r[9] = t.ld;
  }
}
// upper-tree steps (multi-GPU), 4 columns: node (0-based), local block column, owner rank, slot
int spllt_b200_num_top_steps(void* akeep) { return (int)AA(akeep)->top_steps.size(); }
void spllt_b200_get_top_steps(void* akeep, int* out) {
  const Analysis& A = *AA(akeep);
  for (size_t t = 0; t < A.top_steps.size(); ++t) {
    out[4 * t] = A.top_steps[t].node; out[4 * t + 1] = A.top_steps[t].c;
    out[4 * t + 2] = A.top_steps[t].owner; out[4 * t + 3] = A.top_steps[t].slot;
  }
}
void spllt_b200_get_bcol_owner(void* akeep, int* out) {
  const Analysis& A = *AA(akeep);
  for (int g = 0; g < A.nbcol; ++g) out[g] = A.bcol_owner[g];
}

// ---- multi-GPU bootstrap (one process per GPU).  After spllt_b200_partition on every rank:
//   spllt_b200_comm_export  -> 128 bytes (CUDA IPC handles of the rank's arena and flag block);
//   all-gather them over the ranks (torch.distributed, MPI, a file -- the library does not care);
//   spllt_b200_comm_attach  <- the table of world x 128 bytes: maps the peers' memory.
// From then on spllt_factor / spllt_b200_factor_dev run the distributed factorization: every
// rank calls them, exactly as on one GPU.
int spllt_b200_comm_export(void* fkeep, void* out128) {
  return guarded(nullptr, [&] { EE(fkeep)->export_handles(out128); });
}
int spllt_b200_comm_attach(void* fkeep, int rank, int world, const void* all_handles) {
  return guarded(nullptr, [&] { EE(fkeep)->attach_peers(rank, world, all_handles, nullptr); });
}

// Several ranks EMULATED on one GPU inside one process (tests on a single-GPU box; the engines
// were partitioned with rank = 0 .. world-1 of `world`): the per-rank programs are enqueued on one
// stream, interleaved step by step in an order in which every flag is already raised when its
// waiter starts (owner's chain + push first, then everybody's updates), barriers split into
// arrive / wait rounds.  Same kernels, same peer addressing, no kernel ever waits for a later one.
int spllt_b200_emulate_ranks_factor(void** fkeeps, int world, const double* d_val) {
  return guarded(nullptr, [&] {
    std::vector<Engine*> E(world);
    std::vector<void*> table(2 * world);
    for (int r = 0; r < world; ++r) {
      E[r] = EE(fkeeps[r]);
      E[r]->upload_tables();
      table[2 * r] = E[r]->arena;
      table[2 * r + 1] = E[r]->d_flags;
    }
    cudaStream_t st = E[0]->stream;
    for (int r = 0; r < world; ++r)
      if (!E[r]->comm_ready) E[r]->attach_peers(r, world, nullptr, table.data());
    for (int r = 0; r < world; ++r) {
      E[r]->ev_used = 0;
      E[r]->factor_begin(d_val, st);
      E[r]->factor_barrier(1, 1, st);
    }
    for (int r = 0; r < world; ++r) {
      E[r]->factor_barrier(1, 2, st);
      E[r]->enqueue_range(0, E[r]->phase0_end, st);
      E[r]->apply_generated(st);
      E[r]->factor_barrier(2, 1, st);
    }
    for (int r = 0; r < world; ++r) E[r]->factor_barrier(2, 2, st);
    const Analysis& A0 = *E[0]->A;
    for (size_t t = 0; t < A0.top_steps.size(); ++t) {
      const int o = A0.top_steps[t].owner;
      E[o]->enqueue_range(E[o]->step_ranges[t].begin, E[o]->step_ranges[t].after_push, st);
      for (int r = 0; r < world; ++r)
        E[r]->enqueue_range(r == o ? E[r]->step_ranges[t].after_push : E[r]->step_ranges[t].begin,
                            E[r]->step_ranges[t].end, st);
    }
    for (int r = 0; r < world; ++r) {
      E[r]->factor_end(st);
      E[r]->factored = true;
      E[r]->dinv_valid = true;
    }
    CK(cudaStreamSynchronize(st));
  });
}

// max |a - b| / max |b| over the lower trapezoids of the nodes rank A holds (its subtrees + the
// upper tree; every node on one GPU), A and B being two factorizations of the same matrix with the
// same ordering on the same device (e.g. distributed vs single-GPU).  out2 = {max abs diff, max |b|}
int spllt_b200_compare_factor(void* akeep_a, void* fkeep_a, void* akeep_b, void* fkeep_b, double* out2) {
  return guarded(nullptr, [&] {
    const Analysis &A = *AA(akeep_a), &B = *AA(akeep_b);
    Engine *ea = EE(fkeep_a), *eb = EE(fkeep_b);
    struct CmpNode { i64 off_a, off_b; int m, n, ld, pad; };
    std::vector<CmpNode> v;
    for (int k = 0; k < A.nnodes && k < B.nnodes; ++k) {
      const HNode &a = A.nodes[k], &b = B.nodes[k];
      if (A.world > 1 && a.owner >= 0 && a.owner != A.rank) continue;
      if (a.m != b.m || a.n != b.n || a.ld != b.ld) {
        out2[0] = out2[1] = -1.0;   // different analyses
        return;
      }
      v.push_back({a.off, b.off, a.m, a.n, a.ld, 0});
    }
    ea->sync();
    eb->sync();
    CmpNode* d = nullptr;
    unsigned long long* o = nullptr;
    CK(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(CmpNode)));
    CK(cudaMalloc(&o, 16));
    CK(cudaMemcpy(d, v.data(), v.size() * sizeof(CmpNode), cudaMemcpyHostToDevice));
    CK(cudaMemset(o, 0, 16));
    launch_compare_nodes(d, (int)v.size(), ea->arena, eb->arena, o, nullptr);
    unsigned long long h[2];
    CK(cudaMemcpy(h, o, 16, cudaMemcpyDeviceToHost));
    cudaFree(d);
    cudaFree(o);
    memcpy(out2, h, 16);
  });
}
int spllt_b200_dist_top(void* akeep) { return AA(akeep)->dist_top; }
void spllt_b200_solve_phase(void* fkeep, int nrhs, double* d_x, int ldx, int phase) {
  EE(fkeep)->solve_phase(d_x, ldx, nrhs, phase);
}
void* spllt_b200_xw_ptr(void* fkeep, int nrhs) {
  Engine* e = EE(fkeep);
  e->upload_tables();
  e->ensure_solve_buffers(nrhs);
  return e->d_xw;
}

}  // extern "C"
