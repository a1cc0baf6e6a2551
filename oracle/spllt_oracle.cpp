// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product; nothing under spllt_b200/
// may include, link or call this file.  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py load it, as the checker.
//
// CPU restatement (C++17 + sequential BLAS) of the numerical phase of NLAFET/SpLLT:
//   * the post-SSIDS part of spllt_analyse  (src/spllt_analyse_mod.F90:210-558, 806-1171)
//   * spllt_stf_factorize and every factor kernel (src/spllt_stf_mod.F90:84-165,
//     src/spllt_factorization_mod.F90:39-261,474-751, src/spllt_kernels_mod.F90)
//   * the solve set-up and sweeps (src/spllt_solve_dep_mod.F90:1684-1761,1861-2143,
//     src/spllt_solve_mod.F90:167-411, src/spllt_solve_kernels_mod.F90:11-484,
//     src/include/spllt_solve_{fwd,bwd}_{block,update}_worker.F90.inc)
//   * the acceptance metric check_backward_error (src/utils_mod.F90:191-294,432-478)
// Each function cites the lines it follows.  Index conventions are the reference's
// (1-based ids and offsets) so that integer tables can be compared bit-for-bit.
//
// PARITY STATUS: the reference cannot be compiled here (no Fortran compiler, no SPRAL),
// and ships no golden factors or index tables.  This oracle is pinned by
// (1) the 3x3 known answer of example/C/simple.c:25-52, (2) the backward-error gate
// 1e-14 of src/utils_mod.F90:467, and (3) dense Cholesky / dense solves computed
// independently with LAPACK (scipy) on small matrices.  Symbolic inputs
// (sptr/sparent/rptr/rlist/order) come from the SSIDS stand-in and are "parity unpinned".
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t i64;

// ---------------------------------------------------------------- BLAS (sequential)
typedef void (*dpotrf_t)(const char*, const int*, double*, const int*, int*);
typedef void (*dtrsm_t)(const char*, const char*, const char*, const char*, const int*, const int*,
                        const double*, const double*, const int*, double*, const int*);
typedef void (*dsyrk_t)(const char*, const char*, const int*, const int*, const double*, const double*,
                        const int*, const double*, double*, const int*);
typedef void (*dgemm_t)(const char*, const char*, const int*, const int*, const int*, const double*,
                        const double*, const int*, const double*, const int*, const double*, double*,
                        const int*);
typedef void (*dgemv_t)(const char*, const int*, const int*, const double*, const double*, const int*,
                        const double*, const int*, const double*, double*, const int*);
typedef void (*dtrsv_t)(const char*, const char*, const char*, const int*, const double*, const int*,
                        double*, const int*);
static dpotrf_t p_dpotrf;
static dtrsm_t p_dtrsm;
static dsyrk_t p_dsyrk;
static dgemm_t p_dgemm;
static dgemv_t p_dgemv;
static dtrsv_t p_dtrsv;

extern "C" int orc_init_blas(const char* libpath) {
  void* h = dlopen(libpath, RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    fprintf(stderr, "orc_init_blas: %s\n", dlerror());
    return -1;
  }
  const char* pre[] = {"scipy_", ""};
  for (const char* p : pre) {
    char nm[64];
    auto sym = [&](const char* s) {
      snprintf(nm, sizeof nm, "%s%s", p, s);
      return dlsym(h, nm);
    };
    if (!sym("dpotrf_")) continue;
    p_dpotrf = (dpotrf_t)sym("dpotrf_");
    p_dtrsm = (dtrsm_t)sym("dtrsm_");
    p_dsyrk = (dsyrk_t)sym("dsyrk_");
    p_dgemm = (dgemm_t)sym("dgemm_");
    p_dgemv = (dgemv_t)sym("dgemv_");
    p_dtrsv = (dtrsv_t)sym("dtrsv_");
    typedef void (*setnt_t)(int);
    setnt_t snt = (setnt_t)sym("openblas_set_num_threads");
    if (snt) snt(1);  // "sequential BLAS" as README.md:122 recommends
    return 0;
  }
  return -2;
}

static const double one = 1.0, zero = 0.0, mone = -1.0;

// ---------------------------------------------------------------- data model
// src/spllt_data_mod.F90:123-172 (spllt_block), :212-248 (spllt_node), :104-118 (lfactor, lmap)
struct Block {
  i64 id = 0;
  int blkm = 0, blkn = 0;
  i64 sa = 0;  // 1-based offset in its block column
  i64 dblk = 0, last_blk = 0;
  int node = 0, bcol = 0, dep_initial = 0;
};
struct SBlock {  // src/spllt_data_mod.F90:175-209
  int id = 0, blkm = 0, blkn = 0, sa = 0, dblk = 0, last_blk = 0, bcol = 0, node = 0;
  int ldu = 0;
  i64 upd = -1;        // offset of p_upd: >=0 into y, <0 encoded as -(w offset)-2
  int idx_off = 0;     // p_index => node%index(idx_off+1 : idx_off+blkm)
  std::vector<int> fwd_wdep, bwd_wdep;
};
struct Node {
  int sa = 0, en = 0, parent = 0, nchild = 0, least_desc = 0, nb = 0, num = 0;
  i64 blk_sa = 0, blk_en = 0;
  int sblk_sa = 0, sblk_en = 0, snb = 0;
  std::vector<int> index, child;
  std::vector<double> buffer;  // generated element of a subtree root
};
struct Tree {
  int num, nnode, node_sa, node_en;
};

struct Oracle {
  int n = 0, nnodes = 0, nb = 0, ncpu = 1, prune = 1, min_width_blas = 8;
  i64 nz = 0;
  std::vector<Node> nodes_;  // index node+1  (-1 .. nnodes+1)
  Node& nd(int i) { return nodes_[i + 1]; }
  std::vector<Block> bc;     // 1-based
  i64 final_blk = 0;
  int nbcol = 0, maxmn = 0;
  std::vector<i64> weight;   // 1..nnodes+1
  std::vector<int> small;    // 1..nnodes
  std::vector<std::vector<double>> lcol;        // 1..nbcol
  std::vector<std::vector<i64>> lmap_dst, lmap_src;  // 1..nbcol (1-based values)
  std::vector<int> porder, order;               // 1-based
  // solve
  std::vector<SBlock> sbc;  // 1-based
  std::vector<Tree> trees;
  std::vector<int> assoc_tree;
  i64 worksize = 0;
  int s_nrhs = 0;
  std::vector<double> y, w;
  int info_flag = 0;
};

// ---------------------------------------------------------------- heap sort
// src/spllt_utils_mod.F90:22-140 (spllt_sortl / pushdownl with map)
static void pushdown(int root, int last, i64* a, int* map) {  // 1-based arrays
  i64 rv = a[root];
  int rm = map[root];
  int ins = root, test = 2 * ins;
  while (test <= last) {
    if (test != last && a[test + 1] > a[test]) test++;
    if (a[test] <= rv) break;
    a[ins] = a[test];
    map[ins] = map[test];
    ins = test;
    test = 2 * ins;
  }
  a[ins] = rv;
  map[ins] = rm;
}
static void sort_with_map(i64* a, int n, int* map) {
  if (n <= 1) return;
  for (int r = n / 2; r >= 1; --r) pushdown(r, n, a, map);
  for (int i = n; i >= 2; --i) {
    std::swap(a[1], a[i]);
    std::swap(map[1], map[i]);
    pushdown(1, i - 1, a, map);
  }
}

// ---------------------------------------------------------------- analyse
// src/spllt_analyse_mod.F90:990-1029 (spllt_symbolic)
static void symbolic_weights(Oracle& o, const int* sptr, const int* sparent, const i64* rptr) {
  o.weight.assign(o.nnodes + 2, 0);
  for (int node = 1; node <= o.nnodes; ++node) {
    int parent = sparent[node - 1];
    i64 m = rptr[node] - rptr[node - 1];
    i64 n = sptr[node] - sptr[node - 1];
    i64 mm = m - n, nflops = 0;
    for (i64 j = 1; j <= n; ++j) nflops += (mm + j) * (mm + j);
    o.weight[node] += nflops;
    o.weight[parent] += o.weight[node];
  }
}

// src/spllt_analyse_mod.F90:806-987 (spllt_prune_tree)
static void prune_tree(Oracle& o, int nth) {
  int nn = o.nnodes;
  std::vector<i64> lzero_w(nn + 2);
  std::vector<int> lzero(nn + 2);
  std::vector<i64> proc_w(nth);
  double smallth = (double)0.01f;  // `smallth = 0.01` assigns a default-real literal (:840)
  int nlz = 0, leaves = 0, totleaves = 0;
  i64 totflops = 0;
restart:
  totleaves = 0;
  std::fill(o.small.begin(), o.small.end(), 0);
  totflops = o.weight[nn + 1];
  nlz = 0;
  {
    int node = nn + 1;
    nlz++;
    lzero[nlz] = node;
    lzero_w[nlz] = -o.weight[node];
  }
  for (int node = 1; node <= nn + 1; ++node)
    if (o.nd(node).nchild == 0) totleaves++;
  leaves = 0;
  double lim = (double)nth * std::max(2.0, std::pow(std::log((double)nth) / std::log(2.0), 2));
  for (;;) {  // godown
    if (nlz <= 0) break;
    if ((double)nlz > lim) break;
    std::fill(proc_w.begin(), proc_w.end(), 0);
    sort_with_map(lzero_w.data(), nlz, lzero.data());
    for (int node = 1; node <= nlz; ++node) {
      int p = (int)(std::min_element(proc_w.begin(), proc_w.end()) - proc_w.begin());
      proc_w[p] += std::llabs(lzero_w[node]);
    }
    // the reference divides two default reals here (real(minval)/real(maxval))
    float rm = (float)*std::min_element(proc_w.begin(), proc_w.end()) /
               (float)*std::max_element(proc_w.begin(), proc_w.end());
    if (rm > 0.9f && nlz >= nth) break;
    bool found = false, bottom = false;
    int n = 0;
    for (;;) {  // findn
      if (leaves == totleaves) {
        bottom = true;
        break;
      }
      if (leaves == nlz) {
        if ((double)nlz >= lim) {
          bottom = true;
          break;
        }
        smallth = smallth / 2.0;
        if (smallth < (double)1e-4f) {
          bottom = true;
          break;
        }
        goto restart;
      }
      n = lzero[leaves + 1];
      for (int c : o.nd(n).child) {
        if ((double)o.weight[c] > smallth * (double)totflops) {
          found = true;
          nlz++;
          lzero[nlz] = c;
          lzero_w[nlz] = -o.weight[c];
        } else {
          for (int k = o.nd(c).least_desc; k <= c; ++k) o.small[k] = -c;
          o.small[c] = 1;
        }
      }
      if (found) break;
      leaves++;
    }
    if (bottom) break;
    lzero[leaves + 1] = lzero[nlz];
    lzero_w[leaves + 1] = lzero_w[nlz];
    nlz--;
  }
  for (int i = 1; i <= nlz; ++i) {
    int n = lzero[i];
    for (int c : o.nd(n).child) {
      for (int k = o.nd(c).least_desc; k <= c; ++k) o.small[k] = -c;
      o.small[c] = 1;
    }
  }
}

// src/spllt_analyse_mod.F90:1033-1087 (spllt_make_map)
static void make_map(int n, const int* perm, const int* optr, const int* orow, std::vector<i64>& nptr,
                     std::vector<int>& nrow, std::vector<i64>& map) {
  nptr.assign(n + 4, 0);  // 1-based, n+3 used
  i64 ne = optr[n] - 1;
  nrow.assign(ne + 1, 0);
  map.assign(ne + 1, 0);
  for (int i = 1; i <= n; ++i) {
    int l = perm[i - 1];
    for (i64 j = optr[i - 1]; j <= optr[i] - 1; ++j) {
      int k = perm[orow[j - 1] - 1];
      if (k < l)
        nptr[k + 2]++;
      else
        nptr[l + 2]++;
    }
  }
  nptr[1] = 1;
  nptr[2] = 1;
  for (int i = 2; i <= n; ++i) nptr[i + 1] = nptr[i] + nptr[i + 1];
  for (int i = 1; i <= n; ++i) {
    int l = perm[i - 1];
    for (i64 j = optr[i - 1]; j <= optr[i] - 1; ++j) {
      int k = perm[orow[j - 1] - 1];
      if (k < l) {
        map[nptr[k + 1]] = j;
        nrow[nptr[k + 1]] = l;
        nptr[k + 1]++;
      } else {
        map[nptr[l + 1]] = j;
        nrow[nptr[l + 1]] = k;
        nptr[l + 1]++;
      }
    }
  }
}

// src/spllt_analyse_mod.F90:1093-1171 (spllt_lcol_map)
static void lcol_map(Oracle& o, const std::vector<i64>& aptr, const std::vector<int>& arow,
                     const std::vector<i64>& amap) {
  std::vector<i64> map(o.n + 1, 0);
  o.lmap_dst.assign(o.nbcol + 1, {});
  o.lmap_src.assign(o.nbcol + 1, {});
  for (int snode = 1; snode <= o.nnodes; ++snode) {
    Node& nd = o.nd(snode);
    for (size_t j = 1; j <= nd.index.size(); ++j) map[nd.index[j - 1]] = (i64)j - 1;
    i64 dblk = nd.blk_sa;
    int l_nb = nd.nb, sa = nd.sa, en = nd.en;
    for (int cb = sa; cb <= en; cb += l_nb) {
      int bcol = o.bc[dblk].bcol;
      i64 offset = o.bc[dblk].sa - (i64)(cb - sa) * o.bc[dblk].blkn;
      int swidth = o.bc[dblk].blkn;
      int cend = std::min(cb + l_nb - 1, en);
      i64 len = aptr[cend + 1] - aptr[cb];
      o.lmap_dst[bcol].reserve(len);
      o.lmap_src[bcol].reserve(len);
      for (int col = cb; col <= cend; ++col) {
        for (i64 j = aptr[col]; j <= aptr[col + 1] - 1; ++j) {
          i64 i = map[arow[j]];
          o.lmap_dst[bcol].push_back(offset + i * swidth);
          o.lmap_src[bcol].push_back(amap[j]);
        }
        offset++;
      }
      dblk = o.bc[dblk].last_blk + 1;
    }
  }
}

// src/spllt_analyse_mod.F90:210-558 (everything after the ssids_analyse call)
extern "C" void* orc_analyse(int n, const int* ptr, const int* row, const int* order, int nnodes,
                             const int* sptr, const int* sparent, const i64* rptr, const int* rlist,
                             int nb, int ncpu, int prune, int min_width_blas) {
  Oracle* po = new Oracle();
  Oracle& o = *po;
  o.n = n;
  o.nnodes = nnodes;
  o.ncpu = ncpu;
  o.prune = prune;
  o.min_width_blas = min_width_blas;
  o.nz = ptr[n] - 1;
  if (n == 0) return po;
  symbolic_weights(o, sptr, sparent, rptr);  // :213
  o.nodes_.assign(nnodes + 3, Node());       // nodes(-1:num_nodes+1)  :220
  o.nd(0).blk_en = 0;
  o.nd(1).blk_sa = 1;
  o.nd(1).sa = 1;
  for (int node = 1; node <= nnodes + 1; ++node) {  // :236-251
    if (node <= nnodes) {
      int par = sparent[node - 1];
      o.nd(node).parent = par;
      o.nd(par).nchild++;
    } else {
      o.nd(node).parent = -1;
    }
  }
  for (int node = 1; node <= nnodes + 1; ++node) {  // :258-269
    int par = o.nd(node).parent;
    if (par > 0) o.nd(par).child.push_back(node);
  }
  for (int node = -1; node <= nnodes; ++node) o.nd(node).least_desc = node;  // :273-290
  o.nd(nnodes + 1).least_desc = -1;
  for (int node = 1; node <= nnodes; ++node) {
    int an = o.nd(node).parent;
    if (o.nd(an).least_desc == -1)
      o.nd(an).least_desc = o.nd(node).least_desc;
    else
      o.nd(an).least_desc = std::min(o.nd(node).least_desc, o.nd(an).least_desc);
  }
  o.small.assign(nnodes + 1, 0);  // :293-302
  if (prune) prune_tree(o, ncpu);
  for (int node = 1; node <= nnodes; ++node) {  // :305-358
    Node& nd = o.nd(node);
    nd.sa = sptr[node - 1];
    nd.en = sptr[node] - 1;
    nd.num = node;
    int l_nb = nb;
    if (l_nb < 1) l_nb = 256;  // nb_default, src/spllt_data_mod.F90:39
    nd.nb = l_nb;
    nd.index.assign(rlist + (rptr[node - 1] - 1), rlist + (rptr[node] - 1));
    int sz = (int)((rptr[node] - rptr[node - 1] - 1) / l_nb + 1);
    i64 j = 0;
    for (int i = nd.sa; i <= nd.en; i += l_nb) {
      j += sz;
      sz--;
    }
    nd.blk_en = o.nd(node - 1).blk_en + j;
    if (node < nnodes) o.nd(node + 1).blk_sa = nd.blk_en + 1;
  }
  o.nb = o.nd(1).nb;
  o.final_blk = o.nd(nnodes).blk_en;  // :361
  o.bc.assign(o.final_blk + 1, Block());
  i64 blk = 1;  // :381-469
  o.nbcol = 0;
  o.maxmn = 0;
  for (int node = 1; node <= nnodes; ++node) {
    Node& nd = o.nd(node);
    int sa = nd.sa, en = nd.en, numcol = en - sa + 1, numrow = (int)nd.index.size();
    int l_nb = nd.nb;
    int sz = (numrow - 1) / l_nb + 1;
    int cb = 0, col_used = 0;
    for (int ci = sa; ci <= en; ci += l_nb) {
      i64 k = 1;
      o.nbcol++;
      cb++;
      int blkn = std::min(l_nb, numcol - col_used);
      col_used += blkn;
      i64 dblk = blk;
      int row_used = 0;
      for (blk = dblk; blk <= dblk + sz - 1; ++blk) {
        Block& b = o.bc[blk];
        b.id = blk;
        b.blkm = std::min(l_nb, numrow - row_used);
        row_used += b.blkm;
        b.blkn = blkn;
        o.maxmn = std::max(o.maxmn, std::max(b.blkm, b.blkn));
        b.sa = k;
        b.dblk = dblk;
        b.last_blk = dblk + sz - 1;
        b.node = node;
        b.dep_initial = cb;
        b.bcol = o.nbcol;
        k += (i64)b.blkm * b.blkn;
      }
      o.bc[dblk].dep_initial = cb - 1;
      sz--;
      numrow -= l_nb;
    }
  }
  std::vector<i64> aptr, amap;  // :546-552
  std::vector<int> arow;
  make_map(n, order, ptr, row, aptr, arow, amap);
  lcol_map(o, aptr, arow, amap);
  o.order.assign(order, order + n);
  o.porder.assign(n + 1, 0);  // :555-558
  for (int i = 1; i <= n; ++i) o.porder[order[i - 1]] = i;
  return po;
}

extern "C" void orc_free(void* h) { delete (Oracle*)h; }

// ------------------------------------------------------------ table getters (for tests)
extern "C" i64 orc_final_blk(void* h) { return ((Oracle*)h)->final_blk; }
extern "C" int orc_nbcol(void* h) { return ((Oracle*)h)->nbcol; }
extern "C" int orc_maxmn(void* h) { return ((Oracle*)h)->maxmn; }
// out: 9 columns per block: id blkm blkn sa dblk last_blk node bcol dep_initial
extern "C" void orc_get_blocks(void* h, i64* out) {
  Oracle& o = *(Oracle*)h;
  for (i64 b = 1; b <= o.final_blk; ++b) {
    const Block& k = o.bc[b];
    i64* r = out + 9 * (b - 1);
    r[0] = k.id; r[1] = k.blkm; r[2] = k.blkn; r[3] = k.sa; r[4] = k.dblk; r[5] = k.last_blk;
    r[6] = k.node; r[7] = k.bcol; r[8] = k.dep_initial;
  }
}
// out: 8 columns per node 1..nnodes: sa en parent nchild least_desc nb blk_sa blk_en
extern "C" void orc_get_nodes(void* h, i64* out) {
  Oracle& o = *(Oracle*)h;
  for (int s = 1; s <= o.nnodes; ++s) {
    Node& nd = o.nd(s);
    i64* r = out + 8 * (s - 1);
    r[0] = nd.sa; r[1] = nd.en; r[2] = nd.parent; r[3] = nd.nchild; r[4] = nd.least_desc;
    r[5] = nd.nb; r[6] = nd.blk_sa; r[7] = nd.blk_en;
  }
}
extern "C" void orc_get_small(void* h, int* out) {
  Oracle& o = *(Oracle*)h;
  for (int s = 1; s <= o.nnodes; ++s) out[s - 1] = o.small[s];
}
extern "C" void orc_get_weight(void* h, i64* out) {
  Oracle& o = *(Oracle*)h;
  for (int s = 1; s <= o.nnodes + 1; ++s) out[s - 1] = o.weight[s];
}
extern "C" i64 orc_lmap_len(void* h, int bcol) { return (i64)((Oracle*)h)->lmap_dst[bcol].size(); }
extern "C" void orc_get_lmap(void* h, int bcol, i64* dst, i64* src) {
  Oracle& o = *(Oracle*)h;
  std::copy(o.lmap_dst[bcol].begin(), o.lmap_dst[bcol].end(), dst);
  std::copy(o.lmap_src[bcol].begin(), o.lmap_src[bcol].end(), src);
}
extern "C" i64 orc_lcol_size(void* h, int bcol) {
  Oracle& o = *(Oracle*)h;
  // size as allocated by spllt_activate_node, src/spllt_kernels_mod.F90:2476-2488
  i64 sz = 0;
  for (i64 b = 1; b <= o.final_blk; ++b)
    if (o.bc[b].bcol == bcol) sz += (i64)o.bc[b].blkm * o.bc[b].blkn;
  return sz;
}
extern "C" void orc_get_lcol(void* h, int bcol, double* out) {
  Oracle& o = *(Oracle*)h;
  std::copy(o.lcol[bcol].begin(), o.lcol[bcol].end(), out);
}
// all block columns concatenated in bcol order
extern "C" i64 orc_factor_size(void* h) {
  Oracle& o = *(Oracle*)h;
  i64 s = 0;
  for (int b = 1; b <= o.nbcol; ++b) s += (i64)o.lcol[b].size();
  return s;
}
extern "C" void orc_get_factor(void* h, double* out) {
  Oracle& o = *(Oracle*)h;
  i64 s = 0;
  for (int b = 1; b <= o.nbcol; ++b) {
    std::copy(o.lcol[b].begin(), o.lcol[b].end(), out + s);
    s += (i64)o.lcol[b].size();
  }
}

// ---------------------------------------------------------------- factor kernels
// src/spllt_kernels_mod.F90:1168-1189
static void factor_diag_block(int m, int n, double* dest) {
  int info;
  p_dpotrf("U", &n, dest, &n, &info);
  if (info != 0) return;
  if (m > n) {
    int mn = m - n;
    p_dtrsm("L", "U", "T", "N", &n, &mn, &one, dest, &n, dest + (i64)n * n, &n);
  }
}
// src/spllt_kernels_mod.F90:1217-1229
static void solve_block(int m, int n, double* dest, const double* diag) {
  p_dtrsm("L", "U", "T", "N", &n, &m, &one, diag, &n, dest, &n);
}
// src/spllt_kernels_mod.F90:1261-1292
static void update_block(int m, int n, double* dest, bool diag, int n1, const double* src1,
                         const double* src2) {
  if (diag) {
    p_dsyrk("U", "T", &n, &n1, &mone, src1, &n1, &one, dest, &n);
    if (m > n) {
      int mn = m - n;
      p_dgemm("T", "N", &n, &mn, &n1, &mone, src1, &n1, src2 + (i64)n * n1, &n1, &one,
              dest + (i64)n * n, &n);
    }
  } else {
    p_dgemm("T", "N", &n, &m, &n1, &mone, src1, &n1, src2, &n1, &one, dest, &n);
  }
}
// src/spllt_data_mod.F90:663-683
static i64 get_dest_block(const Block& src1, const Block& src2) {
  i64 sz = src1.last_blk - src1.dblk + 1;
  i64 d = src1.dblk;
  for (i64 i = src1.dblk + 1; i <= src1.id; ++i) {
    d += sz;
    sz--;
  }
  return d + src2.id - src1.id;
}
// src/spllt_kernels_mod.F90:1606-1723.  Lists are 1-based values, stored from [0].
static void update_between_compute_map(const Block& blk, int dcol, const Node& dnode, int scol,
                                       const Node& snode, int* row_list, int* col_list, int& rls,
                                       int& cls, int& s1sa, int& s1en, int& s2sa, int& s2en) {
  cls = 0;
  rls = 0;
  int size_dnode = (int)dnode.index.size(), size_snode = (int)snode.index.size();
  int dcsa = dnode.sa + (dcol - 1) * dnode.nb;
  int dcen = std::min(dnode.sa + dcol * dnode.nb - 1, dnode.en);
  int cptr = 1 + std::min(snode.en - snode.sa + 1, (scol - 1) * snode.nb);
  while (snode.index[cptr - 1] < dcsa) {
    cptr++;
    if (cptr > size_snode) return;
  }
  s1sa = cptr;
  while (snode.index[cptr - 1] <= dcen) {
    col_list[cls++] = snode.index[cptr - 1] - dcsa + 1;
    cptr++;
    if (cptr > size_snode) break;
  }
  s1en = cptr - 1;
  i64 i = dcol + blk.id - blk.dblk;
  int drsa = dnode.index[1 + (i - 1) * dnode.nb - 1];
  int dren = dnode.index[std::min((i64)(1 + i * dnode.nb - 1), (i64)size_dnode) - 1];
  int rptr = s1sa;
  while (snode.index[rptr - 1] < drsa) {
    rptr++;
    if (rptr > size_snode) return;
  }
  s2sa = rptr;
  i = blk.id - blk.dblk + 1;
  i64 dptr_sa = 1 + (dcol - 1 + i - 1) * (i64)dnode.nb;
  i64 dptr = dptr_sa;
  for (rptr = s2sa; rptr <= size_snode; ++rptr) {
    if (snode.index[rptr - 1] > dren) break;
    while (dnode.index[dptr - 1] < snode.index[rptr - 1]) dptr++;
    row_list[rls++] = (int)(dptr - dptr_sa + 1);
  }
  s2en = rptr - 1;
}
// src/spllt_kernels_mod.F90:2010-2053
static void expand_buffer(double* a, int blkn, const int* row_list, int rls, const int* col_list,
                          int cls, int ndiag, const double* buffer) {
  for (int j = 1; j <= rls; ++j) {
    i64 rptr = (i64)(j - 1) * cls;
    i64 cptr = (i64)(row_list[j - 1] - 1) * blkn;
    int imax = cls;
    if (j <= ndiag) imax = j;
    for (int i = 1; i <= imax; ++i) {
      i64 k = cptr + col_list[i - 1];
      a[k - 1] += buffer[rptr + i - 1];
    }
  }
}
// src/spllt_kernels_mod.F90:14-93
static void update_direct(int n, double* dest, int n1, const double* csrc, const double* rsrc,
                          const int* row_list, int rls, const int* col_list, int cls, int ndiag) {
  for (int j = 1; j <= rls; ++j) {
    i64 cptr = (i64)(row_list[j - 1] - 1) * n;
    const double* r = rsrc + (i64)(j - 1) * n1;
    int imax = (j <= ndiag) ? j : cls;
    for (int i = 1; i <= imax; ++i) {
      const double* c = csrc + (i64)(i - 1) * n1;
      double work = 0.0;
      for (int l = 0; l < n1; ++l) work += c[l] * r[l];
      dest[cptr + col_list[i - 1] - 1] -= work;
    }
  }
}
// src/spllt_kernels_mod.F90:2108-2237
static void update_between(int m, int n, const Block& blk, int dcol, const Node& dnode, int n1,
                           int scol, const Node& snode, double* dest, const double* csrc,
                           const double* rsrc, int* row_list, int* col_list, double* buffer,
                           int min_width_blas) {
  (void)m;
  bool diag = (blk.dblk == blk.id);
  int rls, cls, s1sa = 0, s1en = -1, s2sa = 0, s2en = -1;
  update_between_compute_map(blk, dcol, dnode, scol, snode, row_list, col_list, rls, cls, s1sa, s1en,
                             s2sa, s2en);
  if (rls == 0 || cls == 0) return;
  if (n1 >= min_width_blas) {
    int ndiag;
    if (diag) {
      ndiag = s1en - s1sa + 1;
      p_dsyrk("U", "T", &ndiag, &n1, &mone, csrc, &n1, &zero, buffer, &cls);
      int rest = s2en - s2sa + 1 - ndiag;
      if (rest > 0)
        p_dgemm("T", "N", &ndiag, &rest, &n1, &mone, csrc, &n1, rsrc + (i64)n1 * ndiag, &n1, &zero,
                buffer + (i64)cls * ndiag, &cls);
    } else {
      ndiag = 0;
      int mm = s1en - s1sa + 1, nn = s2en - s2sa + 1;
      p_dgemm("T", "N", &mm, &nn, &n1, &mone, csrc, &n1, rsrc, &n1, &zero, buffer, &cls);
    }
    expand_buffer(dest, n, row_list, rls, col_list, cls, ndiag, buffer);
  } else {
    int ndiag = diag ? (s1en - s1sa + 1) : 0;
    update_direct(n, dest, n1, csrc, rsrc, row_list, rls, col_list, cls, ndiag);
  }
}
// src/spllt_kernels_mod.F90:2519-2546
static void build_rowmap(const Node& node, int* rowmap) {
  int a_nr = (int)node.index.size(), a_nb = node.nb, rr = 1;
  for (int row = 1; row <= a_nr; row += a_nb) {
    for (int i = row; i <= std::min(row + a_nb - 1, a_nr); ++i) rowmap[node.index[i - 1]] = rr;
    rr++;
  }
}
// src/spllt_kernels_mod.F90:2446-2516
static void activate_node(Oracle& o, int snode) {
  Node& node = o.nd(snode);
  i64 blk = node.blk_sa;
  int l_nb = node.nb;
  int sz = ((int)node.index.size() - 1) / l_nb + 1;
  for (int i = node.sa; i <= node.en; i += l_nb) {
    i64 dblk = blk, size_bcol = 0;
    int nbcol = o.bc[dblk].bcol;
    for (blk = dblk; blk <= dblk + sz - 1; ++blk) size_bcol += (i64)o.bc[blk].blkm * o.bc[blk].blkn;
    o.lcol[nbcol].assign(size_bcol, 0.0);
    sz--;
  }
}
// src/spllt_kernels_mod.F90:2301-2364
static void init_node(Oracle& o, int snode, const double* val) {
  Node& node = o.nd(snode);
  i64 dblk = node.blk_sa;
  for (int cb = node.sa; cb <= node.en; cb += node.nb) {
    int bcol = o.bc[dblk].bcol;
    std::fill(o.lcol[bcol].begin(), o.lcol[bcol].end(), 0.0);
    const std::vector<i64>&d = o.lmap_dst[bcol], &s = o.lmap_src[bcol];
    for (size_t i = 0; i < d.size(); ++i) o.lcol[bcol][d[i] - 1] = val[s[i] - 1];
    dblk = o.bc[dblk].last_blk + 1;
  }
}

struct Work {
  std::vector<double> workspace;
  std::vector<int> row_list, col_list, map;
};

static inline double* tile(Oracle& o, i64 blk) { return o.lcol[o.bc[blk].bcol].data() + o.bc[blk].sa - 1; }

// src/spllt_factorization_mod.F90:474-563 (task form) == src/spllt_kernels_mod.F90:97-222 (inline form)
static void factorize_node(Oracle& o, Node& node, bool tasks) {
  int numcol = node.en - node.sa + 1, numrow = (int)node.index.size();
  int s_nb = node.nb, nc = (numcol - 1) / s_nb + 1, nr = (numrow - 1) / s_nb + 1;
  i64 dblk = node.blk_sa;
  for (int kk = 1; kk <= nc; ++kk) {
    Block& bkk = o.bc[dblk];
    double* pkk = tile(o, dblk);
    {
      int m = bkk.blkm, n = bkk.blkn;
      if (tasks) {
        // src/spllt_factorization_task_mod.F90:351-480 : depend(inout: bc_kk%c(1))
#pragma omp task firstprivate(m, n, pkk) depend(inout : pkk[0])
        factor_diag_block(m, n, pkk);
      } else
        factor_diag_block(m, n, pkk);
    }
    for (int ii = kk + 1; ii <= nr; ++ii) {
      i64 blk = dblk + ii - kk;
      double* pik = tile(o, blk);
      int m = o.bc[blk].blkm, n = o.bc[blk].blkn;
      if (tasks) {
        // :482-646 : depend(in: bc_kk%c(1)) depend(inout: bc_ik%c(1))
#pragma omp task firstprivate(m, n, pik, pkk) depend(in : pkk[0]) depend(inout : pik[0])
        solve_block(m, n, pik, pkk);
      } else
        solve_block(m, n, pik, pkk);
    }
    for (int jj = kk + 1; jj <= nc; ++jj) {
      i64 blk2 = dblk + jj - kk;
      for (int ii = jj; ii <= nr; ++ii) {
        i64 blk1 = dblk + ii - kk;
        i64 blk = get_dest_block(o.bc[blk2], o.bc[blk1]);
        double *pij = tile(o, blk), *pjk = tile(o, blk2), *pik = tile(o, blk1);
        int m = o.bc[blk].blkm, n = o.bc[blk].blkn, n1 = o.bc[blk1].blkn;
        bool diag = (o.bc[blk].dblk == o.bc[blk].id);
        if (tasks) {
          // :648-890 : depend(in: bc_ik%c(1), bc_jk%c(1)) depend(inout: bc_ij%c(1))
#pragma omp task firstprivate(m, n, n1, diag, pij, pjk, pik) depend(in : pjk[0], pik[0]) depend(inout : pij[0])
          update_block(m, n, pij, diag, n1, pjk, pik);
        } else
          update_block(m, n, pij, diag, n1, pjk, pik);
      }
    }
    dblk = o.bc[dblk].last_blk + 1;
  }
}

// Ancestor walk shared by src/spllt_factorization_mod.F90:630-748 (tasks, whole tree) and
// src/spllt_kernels_mod.F90:388-558 (inline, bounded by the subtree root).
// Returns cptr (1-based position in node%index of the first row not consumed).
static int apply_node_between(Oracle& o, Node& node, int root_limit, bool tasks, std::vector<Work>& ws) {
  int numcol = node.en - node.sa + 1, numrow = (int)node.index.size();
  int s_nb = node.nb, nc = (numcol - 1) / s_nb + 1;
  int a_num = node.parent;
  int cptr = 1 + numcol;
  std::vector<int>& map = ws[0].map;  // submission-side map (fkeep%map / th 0)
  while (a_num > 0) {
    if (root_limit > 0 && a_num > root_limit) break;
    if (a_num > o.nnodes) break;  // virtual root has no columns
    Node& anode = o.nd(a_num);
    for (; cptr <= numrow; ++cptr)
      if (node.index[cptr - 1] >= anode.sa) break;
    if (cptr > numrow) break;
    bool map_done = false;
    for (;;) {  // bcols
      if (cptr > numrow) break;
      if (node.index[cptr - 1] > anode.en) break;
      int cb = (node.index[cptr - 1] - anode.sa) / anode.nb + 1;
      i64 a_dblk = anode.blk_sa;
      for (int jb = 2; jb <= cb; ++jb) a_dblk = o.bc[a_dblk].last_blk + 1;
      int jlast = std::min(anode.sa + cb * anode.nb - 1, anode.en);
      int cptr2;
      for (cptr2 = cptr; cptr2 <= numrow; ++cptr2)
        if (node.index[cptr2 - 1] > jlast) break;
      cptr2--;
      if (!map_done) {
        build_rowmap(anode, map.data());
        map_done = true;
      }
      int ii = map[node.index[cptr - 1]];
      int ilast = cptr;
      auto emit = [&](int rsa, int ren, int iiblk) {
        i64 a_blk = a_dblk + iiblk - cb;
        Block& abc = o.bc[a_blk];
        i64 dblk = node.blk_sa;
        for (int kk = 1; kk <= nc; ++kk) {
          Block& bkk = o.bc[dblk];
          int n1 = bkk.blkn;
          // src/spllt_factorization_task_mod.F90:1215-1219
          i64 csrc = 1 + (i64)(cptr - (kk - 1) * s_nb - 1) * n1;
          i64 rsrc = 1 + (i64)(rsa - (kk - 1) * s_nb - 1) * n1;
          double* lcol1 = o.lcol[bkk.bcol].data();
          double* dest = tile(o, a_blk);
          const double *pc = lcol1 + csrc - 1, *pr = lcol1 + rsrc - 1;
          int scol = bkk.bcol - o.bc[node.blk_sa].bcol + 1;
          int dcol = abc.bcol - o.bc[anode.blk_sa].bcol + 1;
          int m = abc.blkm, n = abc.blkn, mwb = o.min_width_blas;
          const Block* pabc = &abc;
          const Node *pan = &anode, *psn = &node;
          if (tasks) {
            // dependencies: first/last source tiles of each operand + dest tile
            // (src/spllt_factorization_task_mod.F90:1239-1241)
            i64 jk_sa = (cptr - 1) / s_nb - (scol - 1) + dblk, jk_en = (cptr2 - 1) / s_nb - (scol - 1) + dblk;
            i64 ik_sa = (rsa - 1) / s_nb - (scol - 1) + dblk, ik_en = (ren - 1) / s_nb - (scol - 1) + dblk;
            double *d1 = tile(o, jk_sa), *d2 = tile(o, jk_en), *d3 = tile(o, ik_sa), *d4 = tile(o, ik_en);
            std::vector<Work>* pws = &ws;
#pragma omp task firstprivate(m, n, pabc, dcol, pan, n1, scol, psn, dest, pc, pr, mwb, pws) \
    depend(in : d1[0], d2[0], d3[0], d4[0]) depend(inout : dest[0])
            {
              int th = 0;
#ifdef _OPENMP
              th = omp_get_thread_num();
#endif
              Work& w = (*pws)[th];
              update_between(m, n, *pabc, dcol, *pan, n1, scol, *psn, dest, pc, pr, w.row_list.data(),
                             w.col_list.data(), w.workspace.data(), mwb);
            }
          } else {
            Work& w = ws[0];
            update_between(m, n, *pabc, dcol, *pan, n1, scol, *psn, dest, pc, pr, w.row_list.data(),
                           w.col_list.data(), w.workspace.data(), mwb);
          }
          dblk = o.bc[dblk].last_blk + 1;
        }
      };
      int i;
      for (i = cptr; i <= numrow; ++i) {
        int k = map[node.index[i - 1]];
        if (k != ii) {
          emit(ilast, i - 1, ii);
          ii = k;
          ilast = i;
        }
      }
      emit(ilast, i - 1, ii);
      cptr = cptr2 + 1;
    }
    a_num = anode.parent;
  }
  return cptr;
}

// src/spllt_kernels_mod.F90:225-325
static void subtree_expand_buffer(bool is_diag, int cptr, int cptr2, const int* col_list, int rptr,
                                  int rptr2, int* row_list, const Node& node, int am, int an,
                                  const Node& root, int m, const double* workspace, double* buffer) {
  (void)am;
  i64 b_sz = am - an;
  int arow = 1;
  int cls = cptr2 - cptr + 1, rls = rptr2 - rptr + 1;
  int ndiag = is_diag ? m : 0;
  for (int i = 1; i <= rls; ++i) {
    while (root.index[arow - 1] != node.index[rptr + i - 1 - 1]) arow++;
    row_list[i - 1] = arow;
  }
  for (int i = 1; i <= rls; ++i) {
    arow = row_list[i - 1];
    i64 buf_rptr = (i64)(arow - an - 1) * b_sz;
    i64 ii = (i64)(i - 1) * m;
    int imax = cls;
    if (i <= ndiag) imax = i;
    for (int j = 1; j <= imax; ++j) {
      i64 bp = buf_rptr + (col_list[j - 1] - an);
      buffer[bp - 1] += workspace[ii + j - 1];
    }
  }
}

// src/spllt_kernels_mod.F90:328-778 (spllt_subtree_apply_node)
static void subtree_apply_node(Oracle& o, Node& node, int root, double* buffer, std::vector<Work>& ws) {
  int numcol = node.en - node.sa + 1, numrow = (int)node.index.size();
  int s_nb = node.nb, nc = (numcol - 1) / s_nb + 1;
  int cptr = apply_node_between(o, node, root, false, ws);
  Work& w = ws[0];
  Node& anode = o.nd(root);
  int am = (int)anode.index.size(), an = anode.en - anode.sa + 1, b_sz = am - an;
  int buff_col = 1;
  bool map_done = false;
  std::vector<int>& map = w.map;
  int* col_list = w.col_list.data();
  int* row_list = w.row_list.data();
  double* workspace = w.workspace.data();
  for (;;) {  // buff_bcols  :575-776
    if (cptr > numrow) break;
    while (anode.index[buff_col - 1] != node.index[cptr - 1]) buff_col++;
    int cb = (buff_col - an - 1) / anode.nb + 1;
    int jlast = an + std::min(cb * anode.nb, b_sz);
    int cptr2;
    for (cptr2 = cptr; cptr2 <= numrow; ++cptr2)
      if (node.index[cptr2 - 1] > anode.index[jlast - 1]) break;
    cptr2--;
    int acol = buff_col;
    for (int j = cptr; j <= cptr2; ++j) {
      while (anode.index[acol - 1] != node.index[j - 1]) acol++;
      col_list[j - cptr] = acol;
    }
    int m = cptr2 - cptr + 1;
    if (!map_done) {
      int rr = 1;
      for (int row = an + 1; row <= am; row += anode.nb) {
        for (int i = row; i <= std::min(row + anode.nb - 1, am); ++i) map[anode.index[i - 1]] = rr;
        rr++;
      }
      map_done = true;
    }
    int ii = map[node.index[cptr - 1]];
    int ilast = cptr;
    auto emit = [&](int rsa, int ren, int k) {
      bool is_diag = (k == cb);
      int n = ren - rsa + 1;
      i64 dblk = node.blk_sa;
      for (int kk = 1; kk <= nc; ++kk) {
        int n1 = o.bc[dblk].blkn;
        i64 csrc = 1 + (i64)(cptr - (kk - 1) * s_nb - 1) * n1;
        i64 rsrc = 1 + (i64)(rsa - (kk - 1) * s_nb - 1) * n1;
        const double* lc = o.lcol[o.bc[dblk].bcol].data();
        double alpha = (kk == 1) ? 0.0 : 1.0;
        if (is_diag) {
          p_dsyrk("U", "T", &m, &n1, &one, lc + csrc - 1, &n1, &alpha, workspace, &m);
          if (n - m > 0) {
            int nm = n - m;
            p_dgemm("T", "N", &m, &nm, &n1, &one, lc + csrc - 1, &n1, lc + rsrc - 1 + (i64)n1 * m, &n1,
                    &alpha, workspace + (i64)m * m, &m);
          }
        } else {
          p_dgemm("T", "N", &m, &n, &n1, &one, lc + csrc - 1, &n1, lc + rsrc - 1, &n1, &alpha, workspace,
                  &m);
        }
        dblk = o.bc[dblk].last_blk + 1;
      }
      subtree_expand_buffer(is_diag, cptr, cptr2, col_list, rsa, ren, row_list, node, am, an, anode, m,
                            workspace, buffer);
    };
    int i, k = ii;
    for (i = cptr; i <= numrow; ++i) {
      k = map[node.index[i - 1]];
      if (k != ii) {
        // the reference tests the NEW row block k against cb (:630, `is_diag = k.eq.cb`)
        // while flushing block ii < k, so is_diag is always false here: the full
        // product (including the unused upper triangle) goes into the buffer.
        emit(ilast, i - 1, k);
        ii = k;
        ilast = i;
      }
    }
    emit(ilast, i - 1, k);
    cptr = cptr2 + 1;
  }
}

// src/spllt_kernels_mod.F90:1122-1160
static void scatter_block(int s_m, int s_n, const int* rsrc_index, const int* csrc_index, const double* src,
                          int lds, const int* rdest_index, const int* cdest_index, double* dest, int ldd) {
  int dr = 1;
  for (int sr = 1; sr <= s_m; ++sr) {
    int srow = rsrc_index[sr - 1];
    while (rdest_index[dr - 1] != srow) dr++;
    int dc = 1;
    for (int sc = 1; sc <= s_n; ++sc) {
      int scol = csrc_index[sc - 1];
      while (cdest_index[dc - 1] != scol) dc++;
      dest[(i64)(dr - 1) * ldd + dc - 1] -= src[(i64)(sr - 1) * lds + sc - 1];
    }
  }
}

// src/spllt_factorization_mod.F90:39-191 (spllt_subtree_apply_buffer) with the task body of
// src/spllt_factorization_task_mod.F90:14-113
static void subtree_apply_buffer(Oracle& o, int root, bool tasks, std::vector<Work>& ws) {
  Node& rn = o.nd(root);
  int size_snode = (int)rn.index.size();
  int m = size_snode, n = rn.en - rn.sa + 1;
  if (m - n == 0) return;
  int anode = rn.parent;
  int cptr = 1 + n;
  std::vector<int>& map = ws[0].map;
  double* buffer = rn.buffer.data();
  while (anode > 0 && anode <= o.nnodes) {
    Node& an = o.nd(anode);
    for (; cptr <= size_snode; ++cptr)
      if (rn.index[cptr - 1] >= an.sa) break;
    if (cptr > size_snode) break;
    bool map_done = false;
    int a_nb = an.nb;
    for (;;) {
      if (cptr > size_snode) break;
      if (rn.index[cptr - 1] > an.en) break;
      int cb = (rn.index[cptr - 1] - an.sa) / a_nb + 1;
      i64 dblk = an.blk_sa;
      for (int jb = 2; jb <= cb; ++jb) dblk = o.bc[dblk].last_blk + 1;
      int jlast = std::min(an.sa + cb * a_nb - 1, an.en);
      int cptr2;
      for (cptr2 = cptr; cptr2 <= size_snode; ++cptr2)
        if (rn.index[cptr2 - 1] > jlast) break;
      cptr2--;
      if (!map_done) {
        build_rowmap(an, map.data());  // :136-149 builds the same row -> row-block map
        map_done = true;
      }
      int jb = map[rn.index[cptr - 1]];
      int ilast = cptr;
      auto emit = [&](int rsa, int ren, int jbb) {
        i64 dest = dblk + jbb - cb;
        Block& db = o.bc[dest];
        int b_sz = m - n;
        int sm = ren - rsa + 1, sn = cptr2 - cptr + 1;
        i64 bsa = (i64)(rsa - n - 1) * b_sz + cptr - n;
        const int* ri = rn.index.data() + rsa - 1;
        const int* ci = rn.index.data() + cptr - 1;
        const double* src = buffer + bsa - 1;
        const int* rdi = an.index.data() + (jbb - 1) * a_nb;
        const int* cdi = an.index.data() + (cb - 1) * a_nb;
        double* dc = tile(o, dest);
        int ldd = db.blkn;
        if (tasks) {
#pragma omp task firstprivate(sm, sn, ri, ci, src, b_sz, rdi, cdi, dc, ldd) depend(in : buffer[0]) depend(inout : dc[0])
          scatter_block(sm, sn, ri, ci, src, b_sz, rdi, cdi, dc, ldd);
        } else
          scatter_block(sm, sn, ri, ci, src, b_sz, rdi, cdi, dc, ldd);
      };
      int i;
      for (i = cptr; i <= size_snode; ++i) {
        int k = map[rn.index[i - 1]];
        if (k != jb) {
          emit(ilast, i - 1, jb);
          jb = k;
          ilast = i;
        }
      }
      emit(ilast, i - 1, jb);
      cptr = cptr2 + 1;
    }
    anode = an.parent;
  }
}

// src/spllt_kernels_mod.F90:780-821
static void subtree_factorize(Oracle& o, int root, const double* val, std::vector<Work>& ws) {
  Node& rn = o.nd(root);
  std::fill(rn.buffer.begin(), rn.buffer.end(), 0.0);
  for (int node = rn.least_desc; node <= root; ++node) init_node(o, node, val);
  for (int node = rn.least_desc; node <= root; ++node) {
    Node& sn = o.nd(node);
    factorize_node(o, sn, false);
    subtree_apply_node(o, sn, root, rn.buffer.data(), ws);
  }
}

// src/spllt_stf_mod.F90:18-192 (spllt_stf_factorize) + spllt_wait.  nthreads<=1: sequential
// submission order; otherwise OpenMP tasks with the dependency contract of
// src/spllt_factorization_task_mod.F90.
// last_node < nnodes: BOUNDED SAMPLE for timing the CPU arm of bench.py -- only nodes
// 1..last_node are initialised, factorized and applied (a prefix of the postorder is closed under
// descendants, so every node in it sees exactly the updates it sees in the full factorization;
// their updates into later ancestors are computed as usual, the ancestors stay unfactorized).
// last_node must not cut a pruned subtree (small[last_node] >= 0).
static void factor_prefix(void* h, const double* val, int nthreads, int last_node);
extern "C" void orc_factor(void* h, const double* val, int nthreads) {
  factor_prefix(h, val, nthreads, ((Oracle*)h)->nnodes);
}
extern "C" void orc_factor_prefix(void* h, const double* val, int nthreads, int last_node) {
  Oracle& o = *(Oracle*)h;
  factor_prefix(h, val, nthreads, std::max(0, std::min(last_node, o.nnodes)));
}
static void factor_prefix(void* h, const double* val, int nthreads, int last_node) {
  Oracle& o = *(Oracle*)h;
  if (o.n == 0) return;
  bool tasks = nthreads > 1;
  int nw = tasks ? nthreads : 1;
  // spllt_factorization_init  src/spllt_factorization_mod.F90:347-423
  const bool sample = last_node < o.nnodes;
  // sample mode: storage of the nodes beyond the sample (update targets only) is allocated once
  // and left alone afterwards, so that a timed step costs what the sample costs
  if (!sample || (int)o.lcol.size() != o.nbcol + 1) o.lcol.assign(o.nbcol + 1, {});
  std::vector<Work> ws(nw);
  for (Work& w : ws) {
    w.workspace.assign((size_t)o.maxmn * o.maxmn, 0.0);
    w.row_list.assign(o.maxmn, 0);
    w.col_list.assign(o.maxmn, 0);
    w.map.assign(o.n + 1, 0);
  }
  for (int s = 1; s <= o.nnodes; ++s) {
    if (sample && s > last_node && !o.lcol[o.bc[o.nd(s).blk_sa].bcol].empty()) continue;
    activate_node(o, s);
  }
  for (int s = 1; s <= last_node; ++s)
    if (o.small[s] == 1) {
      Node& rn = o.nd(s);
      i64 b = (i64)rn.index.size() - (rn.en - rn.sa + 1);
      rn.buffer.assign((size_t)(b * b), 0.0);
    }
  auto body = [&]() {
    for (int s = 1; s <= o.nnodes; ++s) {
      if (o.small[s] != 0 || s > last_node) continue;
      if (tasks) {
        // src/spllt_factorization_task_mod.F90:1333-1414: inout on every tile of the node
        Oracle* po = &o;
#pragma omp task firstprivate(po, s, val)
        init_node(*po, s, val);
      } else
        init_node(o, s, val);
    }
#pragma omp taskwait
    for (int s = 1; s <= last_node; ++s) {
      if (o.small[s] < 0) continue;
      if (o.small[s] == 1) {
        // spllt_subtree_factorize_apply  src/spllt_factorization_mod.F90:196-261
        Node& rn = o.nd(s);
        double* buf = rn.buffer.data();
        if (tasks) {
          Oracle* po = &o;
          std::vector<Work>* pws = &ws;
          static double dummy;
          double* dep = rn.buffer.empty() ? &dummy : buf;
#pragma omp task firstprivate(po, s, val, pws) depend(out : dep[0])
          {
            int th = 0;
#ifdef _OPENMP
            th = omp_get_thread_num();
#endif
            std::vector<Work> one_ws(1);
            one_ws[0].workspace.swap((*pws)[th].workspace);
            one_ws[0].row_list.swap((*pws)[th].row_list);
            one_ws[0].col_list.swap((*pws)[th].col_list);
            // private row->block map: the submission thread owns ws[0].map
            one_ws[0].map.assign(po->n + 1, 0);
            subtree_factorize(*po, s, val, one_ws);
            one_ws[0].workspace.swap((*pws)[th].workspace);
            one_ws[0].row_list.swap((*pws)[th].row_list);
            one_ws[0].col_list.swap((*pws)[th].col_list);
          }
        } else
          subtree_factorize(o, s, val, ws);
        subtree_apply_buffer(o, s, tasks, ws);
      } else {
        // spllt_factorize_apply_node  src/spllt_factorization_mod.F90:567-751
        factorize_node(o, o.nd(s), tasks);
        apply_node_between(o, o.nd(s), 0, tasks, ws);
      }
    }
#pragma omp taskwait
  };
  if (tasks) {
#pragma omp parallel num_threads(nthreads)
#pragma omp single
    body();
  } else
    body();
}

// ---------------------------------------------------------------- solve set-up
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// src/spllt_data_mod.F90:728-780 (spllt_create_subtree)
static void create_subtree(Oracle& o) {
  o.trees.clear();
  o.assoc_tree.assign(o.nnodes + 1, 0);
  int ntree = 0;
  for (int i = 1; i <= o.nnodes; ++i) ntree += (o.small[i] == 1);
  if (ntree == 0) return;
  Tree t{1, 0, 0, 0};
  for (int i = 1; i <= o.nnodes; ++i) {
    if (o.small[i] < 0 || o.small[i] == 1) {
      if (t.nnode == 0) {
        t.node_en = (o.small[i] != 1) ? -o.small[i] : i;
        t.node_sa = i;
        t.nnode = t.node_en - t.node_sa + 1;
      }
      if (o.small[i] == 1) {
        o.assoc_tree[i] = t.num;
        o.trees.push_back(t);
        t = Tree{t.num + 1, 0, 0, 0};
      }
    }
  }
}

// src/spllt_solve_dep_mod.F90:1861-2030 (get_solve_blocks) + :2033-2143 (sblock_assoc_mem)
static void get_solve_blocks(Oracle& o, int nb, int nrhs) {
  o.worksize = 0;
  int nblock = 0;
  for (int i = 1; i <= o.nnodes; ++i) {
    Node& nd = o.nd(i);
    int ncol = nd.en - nd.sa + 1;
    int nbrow = ceil_div(ncol, nb);
    int nbcol = nbrow;
    nblock += (nbrow + 1) * nbrow / 2;
    nbrow = ceil_div((int)nd.index.size() - ncol, nb);
    nblock += nbrow * nbcol;
  }
  o.sbc.assign(nblock + 1, SBlock());
  int iblock = 1, bcol = 0;
  i64 rhs_sa = 1, w_sa = 1;
  for (int i = 1; i <= o.nnodes; ++i) {
    Node& nd = o.nd(i);
    int ncol = nd.en - nd.sa + 1, nrow = (int)nd.index.size();
    int nbrowL1 = ceil_div(ncol, nb), nbrowL2 = ceil_div(nrow - ncol, nb);
    int nbcol = nbrowL1, nbrow = nbrowL1 + nbrowL2;
    int n = nb, sa = 1, blk_sa = iblock;
    for (int j = 1; j <= nbcol; ++j) {
      sa = 1;
      if (j * nb > ncol) n = ncol - (nbcol - 1) * nb;
      int dblk = iblock, last_blk = dblk + nbrow - j;
      int m = nb, lsa = 1;
      for (int k = j; k <= nbrow; ++k) {
        bool inL1 = (k <= nbrowL1);
        if (k == nbrowL1 + 1) m = nb;  // "Treate L_2": m reset (:1977)
        if (inL1) {
          if (k * nb > ncol) m = ncol - (nbrowL1 - 1) * nb;
        } else {
          if ((k - nbrowL1) * nb > nrow - ncol) m = nrow - ncol - (nbrowL2 - 1) * nb;
        }
        SBlock& b = o.sbc[iblock];
        b.bcol = bcol + j; b.blkm = m; b.blkn = n; b.dblk = dblk; b.id = iblock;
        b.last_blk = last_blk; b.node = i; b.sa = sa;
        if (j == 1) {
          b.idx_off = lsa - 1;
          lsa += m;
          b.ldu = m;
          if (inL1) {
            b.upd = rhs_sa - 1;  // into y
            rhs_sa += (i64)m * nrhs;
          } else {
            o.worksize += (i64)m * nrhs;
            b.upd = -(w_sa - 1) - 2;  // into w
            w_sa += (i64)m * nrhs;
          }
        } else {
          int lblk = blk_sa + k - 1;
          b.idx_off = o.sbc[lblk].idx_off;
          b.ldu = o.sbc[lblk].ldu;
          b.upd = o.sbc[lblk].upd;
        }
        sa += m * n;
        iblock++;
      }
    }
    bcol += nbcol;
    nd.sblk_sa = blk_sa;
    nd.sblk_en = iblock - 1;
    nd.snb = nb;
  }
}

static bool index_intersect(const int* a, int na, const int* b, int nbb) {
  int j = 0, k = 0;
  while (j < na && k < nbb) {
    if (a[j] < b[k]) j++;
    else if (a[j] > b[k]) k++;
    else return true;
  }
  return false;
}

// Data-flow part of src/spllt_solve_dep_mod.F90:27-248,934-1164,1257-1431: which tiles'
// update vectors are reduced into (fwd) / gathered by (bwd) a tile.  The task-ordering
// entries of the reference lists (OpenMP artefacts) contribute nothing to the values.
static void compute_solve_wdep(Oracle& o) {
  for (int s = 1; s <= o.nnodes; ++s) {
    Node& nd = o.nd(s);
    int nfirst = o.sbc[nd.sblk_sa].last_blk - nd.sblk_sa + 1;  // row blocks of the node
    // fwd: tiles of the first block column collect from the children's L2 row blocks
    for (int t = 0; t < nfirst; ++t) {
      SBlock& b = o.sbc[nd.sblk_sa + t];
      const int* bi = nd.index.data() + b.idx_off;
      for (int c : nd.child) {
        if (c > o.nnodes) continue;
        Node& cn = o.nd(c);
        int cncol = cn.en - cn.sa + 1;
        int cL1 = ceil_div(cncol, cn.snb);
        int cfirst = o.sbc[cn.sblk_sa].last_blk - cn.sblk_sa + 1;
        for (int u = cL1; u < cfirst; ++u) {
          SBlock& cbk = o.sbc[cn.sblk_sa + u];
          if (index_intersect(bi, b.blkm, cn.index.data() + cbk.idx_off, cbk.blkm))
            b.fwd_wdep.push_back(cbk.id);
        }
      }
    }
    // bwd: L2 tiles of the last block column gather from the parent's row blocks
    int par = nd.parent;
    if (par >= 1 && par <= o.nnodes) {
      Node& pn = o.nd(par);
      int pfirst = o.sbc[pn.sblk_sa].last_blk - pn.sblk_sa + 1;
      int dlast = o.sbc[nd.sblk_en].dblk;
      int ncol = nd.en - nd.sa + 1;
      int nL1 = ceil_div(ncol, nd.snb);
      for (int blk = dlast + 1; blk <= nd.sblk_en; ++blk) {
        SBlock& b = o.sbc[blk];
        (void)nL1;
        const int* bi = nd.index.data() + b.idx_off;
        for (int t = 0; t < pfirst; ++t) {
          SBlock& pb = o.sbc[pn.sblk_sa + t];
          if (index_intersect(bi, b.blkm, pn.index.data() + pb.idx_off, pb.blkm))
            b.bwd_wdep.push_back(pb.id);
        }
      }
    }
  }
}

extern "C" i64 orc_prepare_solve(void* h, int nb, int nrhs) {
  Oracle& o = *(Oracle*)h;
  create_subtree(o);
  get_solve_blocks(o, nb, nrhs);
  compute_solve_wdep(o);
  o.s_nrhs = nrhs;
  o.y.assign((size_t)o.n * nrhs, 0.0);
  o.w.assign((size_t)o.worksize, 0.0);
  return o.worksize;
}
extern "C" int orc_num_sblocks(void* h) { return (int)((Oracle*)h)->sbc.size() - 1; }
// 9 columns: id blkm blkn sa dblk last_blk bcol node ldu
extern "C" void orc_get_sblocks(void* h, int* out) {
  Oracle& o = *(Oracle*)h;
  for (size_t b = 1; b < o.sbc.size(); ++b) {
    const SBlock& k = o.sbc[b];
    int* r = out + 9 * (b - 1);
    r[0] = k.id; r[1] = k.blkm; r[2] = k.blkn; r[3] = k.sa; r[4] = k.dblk; r[5] = k.last_blk;
    r[6] = k.bcol; r[7] = k.node; r[8] = k.ldu;
  }
}

// ---------------------------------------------------------------- solve kernels
static inline double* upd_ptr(Oracle& o, const SBlock& b) {
  return b.upd >= 0 ? o.y.data() + b.upd : o.w.data() + (-(b.upd + 2));
}
// src/spllt_solve_kernels_mod.F90:11-47
static void slv_solve(int n, int nelim, const double* dest, const char* trans, int nrhs, double* rhs,
                      int ldr) {
  if (nelim == 0) return;
  int inc = 1;
  if (nrhs == 1)
    p_dtrsv("U", trans, "N", &nelim, dest, &n, rhs, &inc);
  else
    p_dtrsm("L", "U", trans, "N", &nelim, &nrhs, &one, dest, &n, rhs, &ldr);
}
// src/spllt_solve_kernels_mod.F90:51-138
static void slv_fwd_update(int m, int nelim, const double* dest, int ldd, int nrhs, const double* rhs,
                           int ldr, double* xlocal, int ldx, bool reset) {
  if (nelim == 0) return;
  double alpha = -1.0, beta = reset ? 0.0 : 1.0;
  if (reset)
    for (int i = 0; i < m; ++i) xlocal[i] = 0.0;
  int inc = 1;
  if (nrhs == 1) {
    if (m - nelim > 10 && nelim > 4) {
      p_dgemv("T", &nelim, &m, &alpha, dest, &ldd, rhs, &inc, &beta, xlocal, &inc);
    } else {
      i64 j = 0;
      for (int i = 0; i < m; ++i) {
        double w = 0.0;
        for (int k = 0; k < nelim; ++k) w -= dest[j++] * rhs[k];
        j += ldd - nelim;
        xlocal[i] = beta * xlocal[i] + w;
      }
    }
  } else {
    p_dgemm("T", "N", &m, &nrhs, &nelim, &alpha, dest, &ldd, rhs, &ldr, &beta, xlocal, &ldx);
  }
}
// src/spllt_solve_kernels_mod.F90:141-210
static void slv_bwd_update(int m, int nelim, const double* dest, int ldd, int nrhs, const double* xlocal,
                           int ldx, double* rhs, int ldr) {
  if (nelim == 0) return;
  int inc = 1;
  if (nrhs == 1) {
    if (m - nelim > 10 && nelim > 4) {
      p_dgemv("N", &nelim, &m, &mone, dest, &ldd, xlocal, &inc, &one, rhs, &inc);
    } else {
      i64 j = 0;
      for (int i = 0; i < m; ++i) {
        double w = xlocal[i];
        for (int k = 0; k < nelim; ++k) rhs[k] -= dest[j++] * w;
        j += ldd - nelim;
      }
    }
  } else {
    p_dgemm("N", "N", &nelim, &nrhs, &m, &mone, dest, &ldd, xlocal, &ldx, &one, rhs, &ldr);
  }
}
// src/spllt_solve_dep_mod.F90:1684-1722 / :1726-1761
static void update_upd(Oracle& o, int blk, int child_blk, int nrhs, bool zero_child) {
  SBlock &b = o.sbc[blk], &c = o.sbc[child_blk];
  const int* bi = o.nd(b.node).index.data() + b.idx_off;
  const int* ci = o.nd(c.node).index.data() + c.idx_off;
  double *pu = upd_ptr(o, b), *pc = upd_ptr(o, c);
  int j = 0, k = 0;
  while (j < b.blkm && k < c.blkm) {
    if (bi[j] < ci[k]) j++;
    else if (bi[j] > ci[k]) k++;
    else {
      for (int r = 0; r < nrhs; ++r) {
        pu[j + (i64)r * b.ldu] += pc[k + (i64)r * c.ldu];
        if (zero_child) pc[k + (i64)r * c.ldu] = 0.0;
      }
      j++;
      k++;
    }
  }
}

// src/spllt_solve_kernels_mod.F90:293-387 with the task bodies of
// src/include/spllt_solve_fwd_block_worker.F90.inc:26-62 and ..fwd_update_worker..:28-43
// (pointer set-up as in src/task_manager_seq.F90:290-303, :417-426)
static void solve_fwd_node(Oracle& o, int node, int nrhs, double* rhs) {
  Node& nd = o.nd(node);
  int n = o.n;
  int numcol = nd.en - nd.sa + 1, nc = ceil_div(numcol, nd.snb);
  int dblk = nd.sblk_sa;
  for (int jj = 1; jj <= nc; ++jj) {
    SBlock& d = o.sbc[dblk];
    double* py = upd_ptr(o, d);
    int ldy = d.ldu;
    const int* pidx = nd.index.data() + d.idx_off;
    bool first = (nd.sblk_sa == dblk);
    // fwd block task
    for (int i = 0; i < d.blkn; ++i)
      for (int r = 0; r < nrhs; ++r) {
        double v = rhs[(o.porder[pidx[i]] - 1) + (i64)r * n];
        py[i + (i64)r * ldy] = first ? v : v + py[i + (i64)r * ldy];
      }
    if (first)
      for (int dep : d.fwd_wdep) update_upd(o, dblk, dep, nrhs, true);
    const double* lc = o.lcol[d.bcol].data();
    slv_solve(d.blkm, d.blkn, lc + d.sa - 1, "T", nrhs, py, ldy);
    // fwd update tasks
    for (int blk = dblk + 1; blk <= d.last_blk; ++blk) {
      SBlock& b = o.sbc[blk];
      bool reduction = (nd.sblk_sa == b.dblk);
      slv_fwd_update(b.blkm, b.blkn, lc + b.sa - 1, b.blkn, nrhs, py, ldy, upd_ptr(o, b), b.ldu,
                     reduction);
      if (reduction)
        for (int dep : b.fwd_wdep) update_upd(o, blk, dep, nrhs, true);
    }
    dblk = d.last_blk + 1;
  }
}

// src/spllt_solve_kernels_mod.F90:390-484 with src/include/spllt_solve_bwd_update_worker.F90.inc:29-45
// and ..bwd_block_worker..:29-37
static void solve_bwd_node(Oracle& o, int node, int nrhs, double* rhs) {
  Node& nd = o.nd(node);
  int n = o.n;
  int numcol = nd.en - nd.sa + 1, nc = ceil_div(numcol, nd.snb);
  int dblk = o.sbc[nd.sblk_en].dblk;
  for (int jj = nc; jj >= 1; --jj) {
    SBlock& d = o.sbc[dblk];
    double* py = upd_ptr(o, d);
    int ldy = d.ldu;
    const double* lc = o.lcol[d.bcol].data();
    for (int blk = d.last_blk; blk >= dblk + 1; --blk) {
      SBlock& b = o.sbc[blk];
      if (nd.sblk_en == b.last_blk)
        for (int dep : b.bwd_wdep) update_upd(o, blk, dep, nrhs, false);
      slv_bwd_update(b.blkm, b.blkn, lc + b.sa - 1, b.blkn, nrhs, upd_ptr(o, b), b.ldu, py, ldy);
    }
    slv_solve(d.blkn, d.blkm, lc + d.sa - 1, "N", nrhs, py, ldy);
    const int* pidx = nd.index.data() + d.idx_off;
    for (int i = 0; i < d.blkn; ++i)
      for (int r = 0; r < nrhs; ++r) rhs[(o.porder[pidx[i]] - 1) + (i64)r * n] = py[i + (i64)r * ldy];
    if (jj > 1) dblk = o.sbc[dblk - 1].dblk;
  }
}

// src/spllt_solve_mod.F90:167-224 (job dispatch), :244-339 (solve_fwd), :341-411 (solve_bwd);
// subtree bodies src/task_manager_seq.F90:677-747.  x is n x nrhs column-major, in/out.
extern "C" int orc_solve(void* h, int nrhs, double* x, int job) {
  Oracle& o = *(Oracle*)h;
  if (o.n == 0) return 0;
  if (job < 0 || job > 2) {
    o.info_flag = -10;  // SPLLT_WARNING_PARAM_VALUE, src/spllt_data_mod.F90:33
    return -10;
  }
  if (nrhs != o.s_nrhs) return -1;
  if (job == 0 || job == 1) {
    for (const Tree& t : o.trees)
      for (int node = t.node_sa; node <= t.node_en; ++node) solve_fwd_node(o, node, nrhs, x);
    for (int node = 1; node <= o.nnodes; ++node)
      if (o.small[node] == 0) solve_fwd_node(o, node, nrhs, x);
  }
  if (job == 0 || job == 2) {
    for (int node = o.nnodes; node >= 1; --node) {
      if (o.small[node] == 0)
        solve_bwd_node(o, node, nrhs, x);
      else if (o.small[node] == 1) {
        const Tree& t = o.trees[o.assoc_tree[node] - 1];
        for (int s = t.node_en; s >= t.node_sa; --s) solve_bwd_node(o, s, nrhs, x);
      }
    }
  }
  return 0;
}

// job=1 leaves the intermediate vector in y (pivot order of L1 row blocks); expose it.
extern "C" void orc_get_y(void* h, double* out) {
  Oracle& o = *(Oracle*)h;
  std::copy(o.y.begin(), o.y.end(), out);
}

// ---------------------------------------------------------------- acceptance metric
// src/utils_mod.F90:191-237 (compute_residual / compute_Ax), :258-294 (norms),
// :432-478 (check_backward_error_multi).  err[r] = ||b-Ax||_2 / (||b||_2 + max|a_ij| ||x||_2).
// Returns the number of right-hand sides with err <= 1e-14.
extern "C" int orc_chkerr(int n, const int* ptr, const int* row, const double* val, int nrhs,
                          const double* x, const double* b, double* err) {
  std::vector<double> res((size_t)n * nrhs, 0.0);
  for (int i = 1; i <= n; ++i)
    for (int j = ptr[i - 1]; j <= ptr[i] - 1; ++j) {
      int r = row[j - 1];
      for (int k = 0; k < nrhs; ++k) {
        res[(r - 1) + (size_t)k * n] += val[j - 1] * x[(i - 1) + (size_t)k * n];
        if (r == i) continue;
        res[(i - 1) + (size_t)k * n] += val[j - 1] * x[(r - 1) + (size_t)k * n];
      }
    }
  double norm_max = 0.0;
  for (int j = 0; j < ptr[n] - 1; ++j) norm_max = std::max(norm_max, std::fabs(val[j]));
  int cpt = 0;
  for (int k = 0; k < nrhs; ++k) {
    double nr = 0, nb_ = 0, nx = 0;
    for (int i = 0; i < n; ++i) {
      double r = b[i + (size_t)k * n] - res[i + (size_t)k * n];
      nr += r * r;
      nb_ += b[i + (size_t)k * n] * b[i + (size_t)k * n];
      nx += x[i + (size_t)k * n] * x[i + (size_t)k * n];
    }
    err[k] = std::sqrt(nr) / (std::sqrt(nb_) + norm_max * std::sqrt(nx));
    if (err[k] == err[k] && err[k] <= 1e-14) cpt++;
  }
  return cpt;
}
