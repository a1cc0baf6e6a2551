// Host analysis of the B200 build: everything spllt_analyse does after the ordering /
// symbolic factorization (src/spllt_analyse_mod.F90:210-558), re-expressed as flat tables,
// plus the value-independent schedules that replace the run-time task DAG:
//   * closed-form tile table (what :338-469 builds with nested counters),
//   * A -> L scatter map (:1033-1171),
//   * tree pruning (:806-987) -- also reused as the subtree -> GPU partitioner,
//   * block-column level sets + tile work lists for the factorization (replaces
//     src/spllt_factorization_mod.F90:474-751 unrolling tasks at run time),
//   * inter-node update maps (replaces update_between_compute_map being recomputed per
//     task, src/spllt_kernels_mod.F90:1606-1723),
//   * level sets for the triangular solves (replaces src/spllt_solve_dep_mod.F90).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <numeric>

#include "model.h"

namespace spllt {

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline i64 rup(i64 a, i64 b) { return (a + b - 1) / b * b; }

// ------------------------------------------------------------------------------------------
// Heap sort of (key, tag) pairs into ascending key order.  The pruning heuristic depends on
// how ties are permuted, so this is the same textbook sift-down variant the reference uses
// (src/spllt_utils_mod.F90:22-140): right child only if strictly larger, stop on <=.
static void sift(std::vector<i64>& key, std::vector<int>& tag, int root, int last) {
  i64 k = key[root];
  int t = tag[root];
  int hole = root;
  for (int ch = 2 * hole + 1; ch <= last; ch = 2 * hole + 1) {
    if (ch < last && key[ch + 1] > key[ch]) ++ch;
    if (key[ch] <= k) break;
    key[hole] = key[ch];
    tag[hole] = tag[ch];
    hole = ch;
  }
  key[hole] = k;
  tag[hole] = t;
}
static void heap_sort(std::vector<i64>& key, std::vector<int>& tag, int cnt) {
  if (cnt <= 1) return;
  for (int r = cnt / 2 - 1; r >= 0; --r) sift(key, tag, r, cnt - 1);
  for (int i = cnt - 1; i >= 1; --i) {
    std::swap(key[0], key[i]);
    std::swap(tag[0], tag[i]);
    sift(key, tag, 0, i - 1);
  }
}

// Tree pruning (src/spllt_analyse_mod.F90:806-987).  Nodes are 0-based, the virtual root is
// node `nnodes`.  small[] uses the reference's encoding (1 = subtree root, -(root+1) =
// inside that subtree, 0 = upper tree) so it can be compared with akeep%small directly.
void prune_tree(Analysis& A, int nth, std::vector<int>& small) {
  const int nn = A.nnodes;
  small.assign(nn, 0);
  if (nn == 0) return;
  if (nth < 1) nth = 1;
  // children lists (ascending), including those of the virtual root
  std::vector<int> cptr(nn + 2, 0), clist(nn);
  for (int s = 0; s < nn; ++s) cptr[(A.nodes[s].parent < 0 ? nn : A.nodes[s].parent) + 1]++;
  for (int s = 0; s <= nn; ++s) cptr[s + 1] += cptr[s];
  {
    std::vector<int> fill(cptr.begin(), cptr.end() - 1);
    for (int s = 0; s < nn; ++s) clist[fill[A.nodes[s].parent < 0 ? nn : A.nodes[s].parent]++] = s;
  }
  int totleaves = 0;
  for (int s = 0; s <= nn; ++s)
    if (cptr[s + 1] == cptr[s]) ++totleaves;
  const i64 total = A.weight[nn];
  const double cap = (double)nth * std::max(2.0, std::pow(std::log((double)nth) / std::log(2.0), 2));
  auto mark = [&](int c) {
    for (int k = A.nodes[c].least_desc; k <= c; ++k) small[k] = -(c + 1);
    small[c] = 1;
  };
  std::vector<i64> key(nn + 1), load(nth);
  std::vector<int> layer(nn + 1);
  int cnt = 0;
  // the reference assigns default-real literals to double variables (:840, :903)
  for (double thr = (double)0.01f;; thr /= 2.0) {
    std::fill(small.begin(), small.end(), 0);
    cnt = 1;
    layer[0] = nn;
    key[0] = -total;
    int stuck = 0;  // entries of the layer known to have no child above the threshold
    bool again = false;
    for (;;) {
      if (cnt <= 0 || (double)cnt > cap) break;
      heap_sort(key, layer, cnt);
      std::fill(load.begin(), load.end(), 0);
      for (int i = 0; i < cnt; ++i) *std::min_element(load.begin(), load.end()) += std::llabs(key[i]);
      float bal = (float)*std::min_element(load.begin(), load.end()) /
                  (float)*std::max_element(load.begin(), load.end());
      if ((double)bal > (double)0.9f && cnt >= nth) break;
      bool grew = false, done = false;
      while (!grew) {
        if (stuck == totleaves) { done = true; break; }
        if (stuck == cnt) {
          if ((double)cnt >= cap || thr / 2.0 < (double)1e-4f) done = true;
          else again = true;
          break;
        }
        int v = layer[stuck];
        for (int q = cptr[v]; q < cptr[v + 1]; ++q) {
          int c = clist[q];
          if ((double)A.weight[c] > thr * (double)total) {
            grew = true;
            layer[cnt] = c;
            key[cnt] = -A.weight[c];
            ++cnt;
          } else {
            mark(c);
          }
        }
        if (!grew) ++stuck;
      }
      if (done || again) break;
      layer[stuck] = layer[cnt - 1];
      key[stuck] = key[cnt - 1];
      --cnt;
    }
    if (!again) break;
  }
  for (int i = 0; i < cnt; ++i)
    for (int q = cptr[layer[i]]; q < cptr[layer[i] + 1]; ++q) mark(clist[q]);
}

// ------------------------------------------------------------------------------------------
int build_analysis(int n, const int* ptr, const int* row, int nb, int nemin, int ncpu, int prune,
                   int ordering, const int* user_order, Analysis& A) {
  A = Analysis();
  A.n = n;
  A.nb = nb < 1 ? 256 : nb;  // nb_default, src/spllt_data_mod.F90:39, src/spllt_analyse_mod.F90:316
  A.nemin = nemin;
  A.ncpu = ncpu < 1 ? 1 : ncpu;
  A.prune = prune;
  if (n <= 0) return 0;
  A.nnz = (i64)ptr[n] - 1;
  int rc = symbolic_analyse(n, ptr, row, nemin, ordering, user_order, A.sym);
  if (rc) return rc;
  const Symbolic& S = A.sym;
  const int nn = A.nnodes = S.nnodes;
  nb = A.nb;

  A.porder.resize(n);
  for (int v = 0; v < n; ++v) A.porder[S.order[v] - 1] = v;

  A.nodes.resize(nn);
  A.index.resize(S.rlist.size());
  for (size_t k = 0; k < S.rlist.size(); ++k) A.index[k] = S.rlist[k] - 1;
  A.col2node.resize(n);
  A.weight.assign(nn + 1, 0);
  i64 blk = 0, off = 0, rowbase = 0, nfac = 0;
  int bcol = 0;
  A.maxmn = 0;
  for (int s = 0; s < nn; ++s) {
    HNode& nd = A.nodes[s];
    nd.sa = S.sptr[s] - 1;
    nd.en = S.sptr[s + 1] - 2;
    nd.n = nd.en - nd.sa + 1;
    nd.m = (int)(S.rptr[s + 1] - S.rptr[s]);
    nd.idx_off = S.rptr[s] - 1;
    nd.parent = S.sparent[s] > nn ? -1 : S.sparent[s] - 1;
    nd.nchild = 0;
    nd.least_desc = s;
    nd.nc = cdiv(nd.n, nb);
    nd.nr = cdiv(nd.m, nb);
    nd.bcol0 = bcol;
    nd.blk0 = blk;
    nd.small = 0;
    nd.owner = 0;
    nd.ld = (int)rup(nd.n, LDPAD);
    nd.off = 0;
    nd.row_base = rowbase;
    rowbase += nd.m - nd.n;
    bcol += nd.nc;
    // sum_{c<nc} (nr - c) tiles
    blk += (i64)nd.nc * nd.nr - (i64)nd.nc * (nd.nc - 1) / 2;
    for (int c = nd.sa; c <= nd.en; ++c) A.col2node[c] = s;
    // flops of the node, src/spllt_analyse_mod.F90:1013-1018
    i64 mm = nd.m - nd.n, f = 0;
    for (i64 j = 1; j <= nd.n; ++j) f += (mm + j) * (mm + j);
    A.weight[s] += f;
    A.weight[nd.parent < 0 ? nn : nd.parent] += A.weight[s];
    nfac += (i64)nd.n * nd.m - (i64)nd.n * (nd.n - 1) / 2;
    A.maxmn = std::max(A.maxmn, std::min(nb, nd.m));
  }
  A.nbcol = bcol;
  A.final_blk = blk;
  A.num_flops = A.weight[nn];
  A.num_factor = nfac;
  for (int s = 0; s < nn; ++s) {
    int p = A.nodes[s].parent;
    if (p >= 0) {
      A.nodes[p].nchild++;
      A.nodes[p].least_desc = std::min(A.nodes[p].least_desc, A.nodes[s].least_desc);
    }
  }
  std::vector<int> small;
  if (prune) {
    prune_tree(A, A.ncpu, small);
    for (int s = 0; s < nn; ++s) A.nodes[s].small = small[s];
  }
  // Arena layout: pruned subtrees first, the upper tree (small == 0) last, so that the part
  // of L that several GPUs contribute to is one contiguous slice (multi-GPU reduction).
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1) A.top_begin = off;
    for (int s = 0; s < nn; ++s) {
      HNode& nd = A.nodes[s];
      if ((nd.small == 0) != (pass == 1)) continue;
      nd.off = off;
      off += rup((i64)nd.m * nd.ld, 16);
    }
  }
  A.arena = off;
  A.own_begin = 0;
  A.own_end = off;
  A.bcol_owner.assign(A.nbcol, 0);
  A.bcol_step.assign(A.nbcol, -1);
  A.top_steps.clear();

  // block column -> node table
  A.bcol_node.resize(A.nbcol);
  A.bcol_c.resize(A.nbcol);
  for (int s = 0; s < nn; ++s)
    for (int c = 0; c < A.nodes[s].nc; ++c) {
      A.bcol_node[A.nodes[s].bcol0 + c] = s;
      A.bcol_c[A.nodes[s].bcol0 + c] = c;
    }

  // ---- A -> L map.  Entry (i, j) of the user's lower triangle lands in pivot column
  // min(p_i, p_j), pivot row max(p_i, p_j); per pivot column the entries keep the order in
  // which a scan of the user's columns meets them (what spllt_make_map's bucket pass yields).
  {
    std::vector<i64> cnt(n + 1, 0);
    for (int j = 0; j < n; ++j) {
      int pj = S.order[j] - 1;
      for (i64 e = ptr[j] - 1; e < ptr[j + 1] - 1; ++e) cnt[std::min(pj, S.order[row[e] - 1] - 1) + 1]++;
    }
    for (int c = 0; c < n; ++c) cnt[c + 1] += cnt[c];
    std::vector<i64> colptr(cnt);
    std::vector<int> prow(A.nnz);
    std::vector<i64> psrc(A.nnz);
    for (int j = 0; j < n; ++j) {
      int pj = S.order[j] - 1;
      for (i64 e = ptr[j] - 1; e < ptr[j + 1] - 1; ++e) {
        int pi = S.order[row[e] - 1] - 1;
        i64 slot = cnt[std::min(pi, pj)]++;
        prow[slot] = std::max(pi, pj);
        psrc[slot] = e;
      }
    }
    A.lmap_ptr.assign(A.nbcol + 1, 0);
    A.lmap_row.resize(A.nnz);
    A.lmap_col.resize(A.nnz);
    A.lmap_src.resize(A.nnz);
    std::vector<int> where(n, -1);  // pivot row -> row position in the current node
    i64 w = 0;
    for (int s = 0; s < nn; ++s) {
      const HNode& nd = A.nodes[s];
      const int* idx = A.index.data() + nd.idx_off;
      for (int r = 0; r < nd.m; ++r) where[idx[r]] = r;
      for (int c = 0; c < nd.nc; ++c) {
        int c0 = nd.sa + c * nb, c1 = std::min(c0 + nb - 1, nd.en);
        for (int col = c0; col <= c1; ++col)
          for (i64 e = colptr[col]; e < colptr[col + 1]; ++e) {
            int r = where[prow[e]];
            if (r < 0) return -3;  // entry outside the symbolic pattern
            A.lmap_row[w] = r;
            A.lmap_col[w] = col - nd.sa;
            A.lmap_src[w] = psrc[e];
            ++w;
          }
        A.lmap_ptr[nd.bcol0 + c + 1] = w;
      }
      for (int r = 0; r < nd.m; ++r) where[idx[r]] = -1;
    }
  }

  // ---- schedule depth of block columns: (s, c) runs after (s, c-1) and after every child
  int ndepth = 0;
  for (int s = 0; s < nn; ++s) A.nodes[s].depth0 = 0;
  for (int s = 0; s < nn; ++s) {
    HNode& nd = A.nodes[s];
    int last = nd.depth0 + nd.nc - 1;
    ndepth = std::max(ndepth, last + 1);
    if (nd.parent >= 0) A.nodes[nd.parent].depth0 = std::max(A.nodes[nd.parent].depth0, last + 1);
  }
  A.ndepth = ndepth;
  return 0;
}

// Closed form of the tile table that src/spllt_analyse_mod.F90:381-469 fills with running
// counters.  Tile r (block row, c <= r < nr) of block column c of a node.
void ref_blocks(const Analysis& A, std::vector<RefBlock>& out) {
  out.resize(A.final_blk);
  const int nb = A.nb;
  for (int s = 0; s < A.nnodes; ++s) {
    const HNode& nd = A.nodes[s];
    i64 id = nd.blk0;
    for (int c = 0; c < nd.nc; ++c) {
      i64 dblk = id + 1;
      int blkn = std::min(nb, nd.n - c * nb);
      for (int r = c; r < nd.nr; ++r, ++id) {
        RefBlock& b = out[id];
        b.id = id + 1;
        b.blkm = std::min(nb, nd.m - r * nb);
        b.blkn = blkn;
        b.sa = 1 + (i64)(r - c) * nb * blkn;
        b.dblk = dblk;
        b.last_blk = dblk + (nd.nr - c) - 1;
        b.node = s + 1;
        b.bcol = nd.bcol0 + c + 1;
        b.dep_initial = (r == c) ? c : c + 1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Factorization schedule.
//
// Depth d holds every block column (s, c) with depth0(s) + c == d.  Inside a depth all block
// columns advance together through their inner panels (width IB):
//   potrf(panel diag) -> trsm(rows below) -> tile update of the rest of the block column
// then one "outer" launch group applies the finished block columns to
//   * the later block columns of the same node (a3, K = width of the block column), and
//   * if the node is complete, to its ancestors (a4, K = n, scatter through q_* maps).
struct Region {       // lower part of columns [jbeg, jend) x rows [max(j, ibeg_min), iend) of one source node
  const HNode* nd;
  int jbeg, jend, ibeg_min, iend, k0, kk, src;
  int excl = 0;       // 1: the launch holds this source node only -> scatter without atomics
};

static i64 count_tiles(const Region& r, int T, int TN) {
  i64 c = 0;
  for (int j0 = r.jbeg; j0 < r.jend; j0 += TN) {
    int istart = std::max(j0, r.ibeg_min);
    if (r.iend > istart) c += (r.iend - istart + T - 1) / T;
  }
  return c;
}

// Algorithmic flops of one tile task: 2 kk per destination entry (i, j) with i >= j inside the
// mt x nt tile -- what the reference's dgemm / dsyrk calls on the same tile would count
// (src/spllt_kernels_mod.F90:1261-1292, :2197-2213) -- independent of the padded T x TN extent
// the kernel issues.
double tile_algo_flops(const TileTask& t) {
  double entries = 0;
  for (int j = 0; j < t.nt; ++j) {
    const int first = std::max(t.i0, t.j0 + j);          // first row with i >= j in this column
    entries += std::max(0, t.i0 + t.mt - first);
  }
  return 2.0 * t.kk * entries;
}

// tiles of T rows x TN columns covering the lower part of the region
static void emit_tiles(Analysis& A, std::vector<TileTask>& dst, const Region& r, int T, int TN) {
  const HNode& nd = *r.nd;
  for (int j0 = r.jbeg; j0 < r.jend; j0 += TN) {
    int nt = std::min(TN, r.jend - j0);
    int istart = std::max(j0, r.ibeg_min);
    for (int i0 = istart; i0 < r.iend; i0 += T) {
      TileTask t;
      t.off = nd.off;
      t.ld = nd.ld;
      t.i0 = i0;
      t.j0 = j0;
      t.k0 = r.k0;
      t.mt = std::min(T, r.iend - i0);
      t.nt = nt;
      t.kk = r.kk;
      t.src = r.src;
      t.qoff = nd.row_base - nd.n;
      t.node = (int)(&nd - A.nodes.data());
      t.pad = r.excl;
      dst.push_back(t);
      A.tile_flops += 2.0 * T * TN * r.kk;
      A.tile_flops_algo += tile_algo_flops(t);
    }
  }
}

void build_factor_schedule(Analysis& A, int tile_l_min) {
  const int nn = A.nnodes, nb = A.nb;
  const char* tw = getenv("SPLLT_B200_TILE_WAVE");
  const i64 tile_wave = tw ? atoll(tw) : 148;
  {
    const char* tn = getenv("SPLLT_B200_TILE_N");
    const bool tma = (A.nb % 2 == 0) && !getenv("SPLLT_B200_NO_TMA");
    A.tile_n = (tma && !(tn && atoi(tn) == 128)) ? 64 : 128;
  }
  A.panel_tasks.clear();
  A.npanel_groups = 0;
  A.tile_tasks.clear();
  A.launches.clear();
  A.tile_flops = 0;
  A.tile_flops_algo = 0;

  // ---- inter-node update maps
  i64 nrows = 0;
  for (int s = 0; s < nn; ++s) nrows += A.nodes[s].m - A.nodes[s].n;
  A.q_base.assign(nrows, 0);
  A.q_ld.assign(nrows, 0);
  A.q_rp.assign(nrows, 0);
  A.rowpos.clear();
  // Multi-GPU: a node of a subtree this rank owns sends its updates into the upper tree to the
  // subtree's generated element (local HBM) instead of the upper-tree block columns themselves,
  // most of which live on other GPUs.  (Scattering every node's update straight into the owners
  // was measured first: 5.7e8 remote reductions per rank on Poisson 100^3 at 8 GPUs against 3.2e7
  // entries of the generated element -- NVLink atomics became the bottleneck of the subtree phase.)
  A.gen.clear();
  A.gen_doubles = 0;
  A.gq_base.clear();
  A.gq_ld.clear();
  A.gq_bcol.clear();
  A.gq_rp.clear();
  std::vector<int> gen_of(nn, -1);          // node -> generated element of its subtree (owned nodes only)
  if (A.world > 1) {
    for (int s = nn - 1; s >= 0; --s) {     // parents before children (postorder numbering)
      const HNode& nd = A.nodes[s];
      if (nd.owner != A.rank) continue;
      const int par = nd.parent;
      if (par >= 0 && A.nodes[par].owner == A.rank) {
        gen_of[s] = gen_of[par];
      } else if (nd.m > nd.n) {
        GenElem g;
        g.root = s;
        g.b = nd.m - nd.n;
        g.off = A.gen_doubles;
        g.map0 = 0;
        A.gen_doubles += rup((i64)g.b * g.b, 16);
        gen_of[s] = (int)A.gen.size();
        A.gen.push_back(g);
      }
    }
  }
  for (int s = 0; s < nn; ++s) {
    const HNode& nd = A.nodes[s];
    const int* idx = A.index.data() + nd.idx_off;
    int r = nd.n;
    if (gen_of[s] >= 0) {
      // rows [r_top, m) belong to the upper tree (pivot order: the upper tree follows every subtree)
      const GenElem& g = A.gen[gen_of[s]];
      const HNode& rt = A.nodes[g.root];
      const int* ridx = A.index.data() + rt.idx_off + rt.n;
      int r_top = nd.m;
      for (int q = nd.n; q < nd.m; ++q)
        if (A.nodes[A.col2node[idx[q]]].owner < 0) { r_top = q; break; }
      if (r_top < nd.m) {
        // positions of rows [r_top, m) in the element = in the root's list of rows below (merge)
        const i64 rp = (i64)A.rowpos.size() - r_top;
        int pa = 0;
        for (int q = r_top; q < nd.m; ++q) {
          while (ridx[pa] != idx[q]) ++pa;
          A.rowpos.push_back(pa);
          const i64 gq = nd.row_base + (q - nd.n);
          A.q_base[gq] = -(1 + g.off + pa);
          A.q_ld[gq] = g.b;
          A.q_rp[gq] = rp;
        }
      }
      // the ancestors inside the subtree are handled below, up to r_top
      const int m_sub = r_top;
      while (r < m_sub) {
        int a = A.col2node[idx[r]];
        const HNode& an = A.nodes[a];
        int r1 = r;
        while (r1 < m_sub && idx[r1] <= an.en) ++r1;
        i64 rp = (i64)A.rowpos.size() - r;
        const int* aidx = A.index.data() + an.idx_off;
        int pa = idx[r] - an.sa;
        // rows [r, m_sub) are rows of a inside the subtree; rows [m_sub, m) are rows of a too (upper
        // tree), but the product entries (i >= m_sub, j in [r, r1)) belong to a's rows -> still needed
        for (int q = r; q < nd.m; ++q) {
          while (aidx[pa] != idx[q]) ++pa;
          A.rowpos.push_back(pa);
        }
        for (int q = r; q < r1; ++q) {
          i64 gq = nd.row_base + (q - nd.n);
          A.q_base[gq] = an.off + (idx[q] - an.sa);
          A.q_ld[gq] = an.ld;
          A.q_rp[gq] = rp;
        }
        r = r1;
      }
      continue;
    }
    while (r < nd.m) {
      int a = A.col2node[idx[r]];
      const HNode& an = A.nodes[a];
      int r1 = r;
      while (r1 < nd.m && idx[r1] <= an.en) ++r1;
      // rows [r, r1) are columns of ancestor a; rows [r, m) are all rows of a
      i64 rp = (i64)A.rowpos.size() - r;
      const int* aidx = A.index.data() + an.idx_off;
      int pa = idx[r] - an.sa;  // rows of a start with its own columns in order
      for (int q = r; q < nd.m; ++q) {
        while (aidx[pa] != idx[q]) ++pa;
        A.rowpos.push_back(pa);
      }
      for (int q = r; q < r1; ++q) {
        i64 g = nd.row_base + (q - nd.n);
        A.q_base[g] = an.off + (idx[q] - an.sa);
        A.q_ld[g] = an.ld;
        A.q_rp[g] = rp;
      }
      r = r1;
    }
  }
  // where the generated elements land in the upper tree: the maps of the subtree ROOTS' rows below
  // (element row / column i = root row n + i), used once per factorization by the apply kernel
  for (GenElem& g : A.gen) {
    const HNode& nd = A.nodes[g.root];
    const int* idx = A.index.data() + nd.idx_off + nd.n;
    g.map0 = (i64)A.gq_base.size();
    A.gq_base.resize(g.map0 + g.b);
    A.gq_ld.resize(g.map0 + g.b);
    A.gq_bcol.resize(g.map0 + g.b);
    A.gq_rp.resize(g.map0 + g.b);
    int i = 0;
    while (i < g.b) {
      const int a = A.col2node[idx[i]];
      const HNode& an = A.nodes[a];
      int i1 = i;
      while (i1 < g.b && idx[i1] <= an.en) ++i1;
      const i64 rp = (i64)A.rowpos.size() - i;
      const int* aidx = A.index.data() + an.idx_off;
      int pa = idx[i] - an.sa;
      for (int q = i; q < g.b; ++q) {
        while (aidx[pa] != idx[q]) ++pa;
        A.rowpos.push_back(pa);
      }
      for (int q = i; q < i1; ++q) {
        A.gq_base[g.map0 + q] = an.off + (idx[q] - an.sa);
        A.gq_ld[g.map0 + q] = an.ld;
        A.gq_bcol[g.map0 + q] = an.bcol0 + (idx[q] - an.sa) / nb;
        A.gq_rp[g.map0 + q] = rp;
      }
      i = i1;
    }
  }

  // ---- level sets, one pass per phase (phase 0: nodes owned by this rank, phase 1: shared top)
  std::vector<TileTask> ts, tl;
  std::vector<Region> regions, regions_bg;
  int cur_phase = 0;
  auto add_tiles = [&](Analysis&, std::vector<TileTask>&, std::vector<TileTask>&, const HNode& nd, int jbeg, int jend,
                       int ibeg_min, int iend, int k0, int kk, int src, int) {
    if (jend <= jbeg || iend <= jbeg) return;
    regions.push_back({&nd, jbeg, jend, ibeg_min, iend, k0, kk, src});
  };
  // Deferring the non-urgent inter-node updates to a low-priority stream is implemented and
  // parity-tested but did not pay off on B200 (FP64 tile CTAs are register-file bound, so the
  // panel kernels cannot share an SM with two of them): opt-in.
  const bool defer = getenv("SPLLT_B200_DEFER") != nullptr;
  auto flush_tiles = [&](int depth, int tag, int stream = 0, int deadline = 0) {
    std::vector<Region>& regions_ = stream ? regions_bg : regions;
    // Tile size per launch: 128 x 128 tiles only pay off when the launch fills the machine;
    // a launch with less than a wave of them is latency bound and runs faster on 64 x 64
    // tiles spread over more SMs.  Tasks are sorted by decreasing work so the tail of every
    // launch consists of small tiles.
    auto& regions = regions_;
    i64 nlarge = 0;
    std::vector<char> big(regions.size(), 0);
    for (size_t i = 0; i < regions.size(); ++i) {
      const Region& r = regions[i];
      big[i] = (r.jend - r.jbeg) >= tile_l_min && (r.iend - r.jbeg) >= tile_l_min && r.kk >= 16;
      if (big[i]) nlarge += count_tiles(r, 128, 128);
    }
    const bool use_large = nlarge >= tile_wave;
    for (size_t i = 0; i < regions.size(); ++i) {
      if (use_large && big[i]) emit_tiles(A, tl, regions[i], 128, A.tile_n);
      else emit_tiles(A, ts, regions[i], 64, 64);
    }
    regions.clear();
    auto work = [](const TileTask& t) { return (i64)t.kk * t.mt * t.nt; };
    std::stable_sort(ts.begin(), ts.end(), [&](const TileTask& a, const TileTask& b) { return work(a) > work(b); });
    std::stable_sort(tl.begin(), tl.end(), [&](const TileTask& a, const TileTask& b) { return work(a) > work(b); });
    if (!ts.empty()) {
      A.launches.push_back({L_TILE_S, depth, (i64)A.tile_tasks.size(), (i64)ts.size(), cur_phase, tag, stream, deadline});
      A.tile_tasks.insert(A.tile_tasks.end(), ts.begin(), ts.end());
      ts.clear();
    }
    if (!tl.empty()) {
      A.launches.push_back({L_TILE_L, depth, (i64)A.tile_tasks.size(), (i64)tl.size(), cur_phase, tag, stream, deadline});
      A.tile_tasks.insert(A.tile_tasks.end(), tl.begin(), tl.end());
      tl.clear();
    }
  };

  // Time slots at PANEL granularity (as-soon-as-possible): a node's first panel runs in the slot
  // after its last child's last panel, and its panels occupy consecutive slots.  The number of
  // slots is the true critical path of the tree in panel steps (162 for 3-D Poisson 64^3,
  // nb = 512) instead of the sum over block-column levels of the widest block column (253).
  // One slot = one panel launch + one tile launch group holding, for every node active in the
  // slot: the inner update of its block column, or (last panel of a block column) the outer
  // update of the node's later block columns, and (last panel of the node) its inter-node
  // updates -- the latency-bound inner updates ride along with the throughput-bound ones.
  struct Step { int node, c, p; };
  std::vector<int> nsteps(nn, 0);
  for (int s = 0; s < nn; ++s)
    for (int c = 0; c < A.nodes[s].nc; ++c) nsteps[s] += cdiv(std::min(nb, A.nodes[s].n - c * nb), IB);
  auto emit_panel = [&](const HNode& nd, int k0, int pw) {
    PanelTask pt;
    pt.d_off = nd.off + (i64)k0 * nd.ld + k0;
    pt.ld = nd.ld;
    pt.pw = pw;
    pt.col0 = nd.sa + k0;
    pt.pad = 0;
    int r = k0 + pw;
    size_t g0 = A.panel_tasks.size();
    pt.group = A.npanel_groups++;
    do {
      pt.r_off = nd.off + (i64)r * nd.ld + k0;
      pt.nrows = std::max(0, std::min(TRSM_ROWS, nd.m - r));
      pt.store = 0;
      A.panel_tasks.push_back(pt);
      r += TRSM_ROWS;
    } while (r < nd.m);
    A.panel_tasks.back().store = 1;
    for (size_t q = g0; q < A.panel_tasks.size(); ++q) A.panel_tasks[q].ngroup = (int)(A.panel_tasks.size() - g0);
  };

  // a3 inside a block column, TWO-LEVEL blocking.  Updating every remaining column of the block
  // column after each 64-column panel (K = 64) makes the block column pass through HBM
  // w / 128 times (11x read-modify-write amplification at w = 768) in tiles that do 1 MFLOP per
  // 128 KB moved.  Instead the panel only updates the rest of its MID-BLOCK (mid_w = 256 columns);
  // when a mid-block is complete, ONE K = mid_w update brings the block column's later columns up
  // to date.  Same arithmetic on every entry (right-looking, every update applied before the
  // entry's own panel), 2.2x less traffic, and most of those flops move to K = 256 tiles.
  // (Measured on one GPU: 0.5 % -- the K = 64 tiles hide behind the other supernodes' tiles of the
  // same launch; it matters for a chain that runs alone.)
  const int mid_w = std::max(IB, (getenv("SPLLT_B200_MID_BLOCK") ? atoi(getenv("SPLLT_B200_MID_BLOCK")) : 256) / IB * IB);
  // Look-ahead: only the NEXT panel's 64 columns must be up to date before its k_panel.  The update
  // of the columns behind them goes to `regions_ahead` (launch tag 3): it is forked onto a side stream
  // and runs while the next panels are factorized -- near the top of the tree a k_panel launch (one CTA
  // per 128 rows of one or two supernodes) leaves most SMs idle.  It is joined before the next tile
  // launch, so the order of the read-modify-write updates of any entry is unchanged.
  const bool chain_ahead = !getenv("SPLLT_B200_NO_CHAIN_AHEAD");
  std::vector<Region> regions_ahead;
  auto inner_updates = [&](const HNode& nd, int r0, int w, int k0, int pw) {
    const int bend = r0 + w;
    const int mb0 = r0 + (k0 - r0) / mid_w * mid_w, me = std::min(mb0 + mid_w, bend);
    const bool inside = k0 + pw < me;
    if (!inside && me >= bend) return;
    const int c0 = inside ? k0 + pw : me, cend = inside ? me : bend;
    const int ks = inside ? k0 : mb0, kw = inside ? pw : me - mb0;
    const int cn = chain_ahead ? std::min(c0 + IB, cend) : cend;
    add_tiles(A, ts, tl, nd, c0, cn, 0, nd.m, ks, kw, -1, tile_l_min);
    if (cn < cend) regions_ahead.push_back({&nd, cn, cend, 0, nd.m, ks, kw, -1});
  };
  auto flush_ahead = [&](int depth) {
    if (regions_ahead.empty()) return;
    regions.insert(regions.end(), regions_ahead.begin(), regions_ahead.end());
    regions_ahead.clear();
    flush_tiles(depth, 3);
  };

  // ---- phase 0: every node on one GPU; the subtrees this rank owns on several
  {
    const int phase = 0;
    cur_phase = phase;
    auto mine = [&](int s) { return A.world <= 1 || A.nodes[s].owner == A.rank; };
    std::vector<int> t0(nn, 0);
    int nslots = 0;
    for (int s = 0; s < nn; ++s) {
      if (!mine(s)) continue;
      int tend = t0[s] + nsteps[s];
      nslots = std::max(nslots, tend);
      int par = A.nodes[s].parent;
      if (par >= 0 && mine(par)) t0[par] = std::max(t0[par], tend);
    }
    std::vector<std::vector<Step>> at(nslots);
    int bg_deadline = 1 << 30;
    std::vector<Region> excl_regions;
    // Exclusive (atomic-free) launches: implemented and parity-tested, but measured SLOWER than
    // RED.ADD.F64 (the read-modify-write needs two dependent global round trips per row block,
    // the RED is fire-and-forget): 64^3 tile time 16.3 -> 24.8 ms.  Opt-in only.
    const bool excl_ok = (A.nb % 2 == 0) && !getenv("SPLLT_B200_NO_TMA") && getenv("SPLLT_B200_EXCL") && !defer &&
                         A.world <= 1;
    const i64 excl_min = getenv("SPLLT_B200_EXCL_MIN") ? atoll(getenv("SPLLT_B200_EXCL_MIN")) : 148;
    for (int s = 0; s < nn; ++s) {
      if (!mine(s)) continue;
      int t = t0[s];
      for (int c = 0; c < A.nodes[s].nc; ++c) {
        int w = std::min(nb, A.nodes[s].n - c * nb);
        for (int p = 0; p * IB < w; ++p) at[t++].push_back({s, c, p});
      }
    }
    for (int d = 0; d < nslots; ++d) {
      if (at[d].empty()) continue;
      i64 p0 = A.panel_tasks.size();
      for (const Step& st : at[d]) {
        const HNode& nd = A.nodes[st.node];
        int r0 = st.c * nb;
        int w = std::min(nb, nd.n - r0);
        int pw = std::min(IB, w - st.p * IB);
        int k0 = r0 + st.p * IB;
        const bool last_in_bcol = k0 + pw >= r0 + w;
        emit_panel(nd, k0, pw);
        inner_updates(nd, r0, w, k0, pw);   // a3 inside the block column (two-level blocking)
        if (last_in_bcol) {
          // a3: the finished block column updates the node's later block columns (K = w)
          if (st.c + 1 < nd.nc) add_tiles(A, ts, tl, nd, r0 + w, nd.n, 0, nd.m, r0, w, -1, tile_l_min);
          // a4: the finished node updates its ancestors (K = n).  Only the columns that belong
          // to an ancestor whose first panel runs in the very next slot are on the critical path;
          // the rest is deferred to the background stream with the slot of the first ancestor
          // that needs it as its deadline.  (Multi-GPU: ancestors in the upper tree live in their
          // owner's arena -- the scatter goes there through the peer-mapped q_base addresses.)
          if (st.c + 1 == nd.nc && nd.m > nd.n) {
            const int* idx = A.index.data() + nd.idx_off;
            int r = nd.n, split = nd.n, dl = 1 << 30;
            bool first = true;
            while (r < nd.m) {
              int a = A.col2node[idx[r]];
              int r1 = r;
              while (r1 < nd.m && idx[r1] <= A.nodes[a].en) ++r1;
              int need = mine(a) ? t0[a] : (1 << 30);   // upper tree: after the barrier
              if (first && need <= d + 1) split = r1;   // urgent: the parent starts in the next slot
              else dl = std::min(dl, need);
              first = false;
              r = r1;
            }
            if (!defer) split = nd.m;
            // A node whose inter-node update fills the machine on its own gets its own launch:
            // all destinations of one source node are distinct, so it needs no atomics.
            Region whole{&nd, nd.n, split, 0, nd.m, 0, nd.n, st.node, 1};
            if (excl_ok && split > nd.n && count_tiles(whole, 128, A.tile_n) >= excl_min) {
              excl_regions.push_back(whole);
            } else if (split > nd.n) add_tiles(A, ts, tl, nd, nd.n, split, 0, nd.m, 0, nd.n, st.node, tile_l_min);
            if (split < nd.m) {
              regions_bg.push_back({&nd, split, nd.m, 0, nd.m, 0, nd.n, st.node});
              bg_deadline = std::min(bg_deadline, dl);
            }
          }
        }
      }
      if ((i64)A.panel_tasks.size() > p0)
        A.launches.push_back({L_PANEL, d, p0, (i64)A.panel_tasks.size() - p0, phase, 0, 0, 0});
      flush_tiles(d, 4);
      flush_ahead(d);
      for (const Region& r : excl_regions) {   // one launch per big finishing node, after the shared ones
        emit_tiles(A, tl, r, 128, A.tile_n);
        std::stable_sort(tl.begin(), tl.end(), [](const TileTask& a, const TileTask& b) {
          return (i64)a.kk * a.mt * a.nt > (i64)b.kk * b.mt * b.nt;
        });
        A.launches.push_back({L_TILE_L, d, (i64)A.tile_tasks.size(), (i64)tl.size(), phase, 6, 0, 0});
        A.tile_tasks.insert(A.tile_tasks.end(), tl.begin(), tl.end());
        tl.clear();
      }
      excl_regions.clear();
      if (!regions_bg.empty()) {
        flush_tiles(d, 5, 1, bg_deadline);
        bg_deadline = 1 << 30;
      }
    }
  }

  // ---- phase 1 (multi-GPU): the upper tree, one STEP per block column in the global order of
  // A.top_steps.  Owner computes: the owner of a step's block column factorizes it (panel chain)
  // and pushes it into every peer's arena; every rank applies it to the destination block columns
  // IT owns -- later block columns of the node, and, when the node is complete, block columns of
  // its ancestors.  Updates are split by urgency (static look-ahead): destinations whose own step
  // comes within the next `window` steps are updated right away, the rest one step later, i.e.
  // while the owner of the next step is busy with its panel chain.  Per rank everything is ONE
  // in-order sequence, so a destination's chain always follows every update into it.
  if (A.world > 1 && A.dist_top) {
    cur_phase = 1;
    const int window = std::max(1, getenv("SPLLT_B200_TOP_WINDOW") ? atoi(getenv("SPLLT_B200_TOP_WINDOW")) : 2);
    std::vector<Region> rest;   // deferred updates of the previous step
    auto flush_rest = [&](int t) {
      if (rest.empty()) return;
      regions.insert(regions.end(), rest.begin(), rest.end());
      rest.clear();
      flush_tiles(t, 5);
    };
    for (int t = 0; t < (int)A.top_steps.size(); ++t) {
      const TopStep& ts_ = A.top_steps[t];
      const HNode& nd = A.nodes[ts_.node];
      const int g = nd.bcol0 + ts_.c;
      const int r0 = ts_.c * nb, w = std::min(nb, nd.n - r0);
      const bool own = ts_.owner == A.rank;
      if (own) {
        for (int k0 = r0; k0 < r0 + w; k0 += IB) {
          const int pw = std::min(IB, r0 + w - k0);
          i64 p0 = A.panel_tasks.size();
          emit_panel(nd, k0, pw);
          A.launches.push_back({L_PANEL, t, p0, (i64)A.panel_tasks.size() - p0, 1, 0, 0, 0});
          if (k0 + pw < r0 + w) {
            inner_updates(nd, r0, w, k0, pw);
            flush_tiles(t, 4);
            flush_ahead(t);
          }
        }
        // delivery: one push per peer, in the order in which the peers need the block column --
        // the owners of the following steps (cyclic ownership: they are all different).  The owner's
        // NVLink egress (rows x w doubles per peer) is the bottleneck of a delivery; in this order the
        // next chain waits for one copy, not for world - 1 of them.
        {
          unsigned done = 1u << A.rank;
          int k = 0;
          for (size_t t2 = t + 1; t2 < A.top_steps.size() && k < A.world - 1; ++t2) {
            const int o2 = A.top_steps[t2].owner;
            if (done >> o2 & 1u) continue;
            done |= 1u << o2;
            A.launches.push_back({L_PUSH, t, (i64)g, k++, 1, 8, 0, 1 << o2});
          }
          for (int o2 = 0; o2 < A.world; ++o2)      // (the last steps: everybody still needs the factor for the solve)
            if (!(done >> o2 & 1u)) A.launches.push_back({L_PUSH, t, (i64)g, k++, 1, 8, 0, 1 << o2});
        }
      } else {
        flush_rest(t);
        A.launches.push_back({L_WAIT, t, (i64)g, 1, 1, 9, 0, 0});
      }
      // updates from this block column (later block columns of the node) / this node (ancestors)
      std::vector<Region> later;
      for (int c2 = ts_.c + 1; c2 < nd.nc; ++c2) {
        const int g2 = nd.bcol0 + c2;
        if (A.bcol_owner[g2] != A.rank) continue;
        Region rg{&nd, c2 * nb, std::min((c2 + 1) * nb, nd.n), 0, nd.m, r0, w, -1};
        (A.bcol_step[g2] <= t + window ? regions : later).push_back(rg);
      }
      if (ts_.c + 1 == nd.nc && nd.m > nd.n) {
        const int* idx = A.index.data() + nd.idx_off;
        int r = nd.n;
        while (r < nd.m) {
          const int a = A.col2node[idx[r]];
          const int cb = (idx[r] - A.nodes[a].sa) / nb;
          const int cend = std::min(A.nodes[a].sa + (cb + 1) * nb - 1, A.nodes[a].en);
          int r1 = r;
          while (r1 < nd.m && idx[r1] <= cend) ++r1;
          const int g2 = A.nodes[a].bcol0 + cb;
          if (A.bcol_owner[g2] == A.rank) {
            Region rg{&nd, r, r1, 0, nd.m, 0, nd.n, ts_.node};
            (A.bcol_step[g2] <= t + window ? regions : later).push_back(rg);
          }
          r = r1;
        }
      }
      flush_tiles(t, 7);
      if (own) flush_rest(t);
      rest.insert(rest.end(), later.begin(), later.end());
    }
    flush_rest((int)A.top_steps.size());
  }
}

// Subtree -> GPU mapping (proportional mapping).  Starting from the roots, the heaviest
// candidate subtree is repeatedly split -- its root joins the shared upper tree, its children
// become candidates -- until the subtrees can be dealt to the ranks (largest first onto the
// least loaded rank) within 10 % of perfect balance.  The upper tree is distributed by block
// column (see below); a subtree's contributions into it are accumulated in the subtree's generated
// element (src/spllt_factorization_mod.F90:224-237) and scattered into the owning ranks' arenas over
// NVLink when the subtree is complete (build_factor_schedule, Engine::apply_generated).  (The reference's pruning,
// src/spllt_analyse_mod.F90:806-987, plays the mapping role for CPU workers but aims at many small
// subtrees: with nth = 8 it leaves 81 % of the flops of a 64^3 Poisson problem in the upper tree.)
void partition_tree(Analysis& A, int rank, int world) {
  A.rank = rank;
  A.world = world;
  const int nn = A.nnodes;
  for (int s = 0; s < nn; ++s) A.nodes[s].owner = (world <= 1) ? 0 : -1;
  if (world > 1 && nn > 0) {
    std::vector<std::vector<int>> child(nn);
    std::vector<int> cand;
    for (int s = 0; s < nn; ++s) {
      if (A.nodes[s].parent >= 0) child[A.nodes[s].parent].push_back(s);
      else cand.push_back(s);
    }
    auto by_weight = [&](int a, int b) { return A.weight[a] > A.weight[b] || (A.weight[a] == A.weight[b] && a < b); };
    std::vector<i64> load(world);
    std::vector<int> where;
    // SPLLT_B200_SUBTREES_PER_RANK = k: keep splitting until every rank gets at least k subtrees
    // (smaller subtrees, more -- and more parallel -- work in the upper tree); SPLLT_B200_BALANCE = the
    // accepted max / mean load of the subtree phase
    const int min_sub = std::max(1, getenv("SPLLT_B200_SUBTREES_PER_RANK") ? atoi(getenv("SPLLT_B200_SUBTREES_PER_RANK")) : 1);
    const double tol = getenv("SPLLT_B200_BALANCE") ? atof(getenv("SPLLT_B200_BALANCE")) : 1.10;
    for (int iter = 0; iter < nn; ++iter) {
      std::sort(cand.begin(), cand.end(), by_weight);
      std::fill(load.begin(), load.end(), 0);
      where.assign(cand.size(), 0);
      i64 total = 0;
      for (size_t k = 0; k < cand.size(); ++k) {
        int p = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        load[p] += A.weight[cand[k]];
        where[k] = p;
        total += A.weight[cand[k]];
      }
      i64 mx = *std::max_element(load.begin(), load.end());
      bool balanced = (int)cand.size() >= min_sub * world && (double)mx * world <= tol * (double)total;
      if (balanced || cand.empty() || child[cand[0]].empty()) break;
      int top = cand[0];   // heaviest: split it
      cand.erase(cand.begin());
      for (int c : child[top]) cand.push_back(c);
    }
    for (size_t k = 0; k < cand.size(); ++k)
      for (int q = A.nodes[cand[k]].least_desc; q <= cand[k]; ++q) A.nodes[q].owner = where[k];
  }
  // ---- upper tree in STEPS: one per block column, ordered by the as-soon-as-possible slot of its
  // first panel (the critical-path order of the tree), ties by node.  Block columns are dealt to
  // the ranks cyclically in that order: consecutive steps -- in particular consecutive block
  // columns of one node -- have different owners, so the panel chain of step t+1 overlaps the
  // updates the other ranks still apply from step t.
  A.dist_top = (world > 1) ? 1 : 0;
  A.top_steps.clear();
  A.bcol_step.assign(A.nbcol, -1);
  A.bcol_owner.assign(A.nbcol, 0);
  for (int s = 0; s < nn; ++s)
    for (int c = 0; c < A.nodes[s].nc; ++c) A.bcol_owner[A.nodes[s].bcol0 + c] = std::max(A.nodes[s].owner, 0);
  if (world > 1) {
    std::vector<int> t0(nn, 0);
    for (int s = 0; s < nn; ++s) {
      if (A.nodes[s].owner >= 0) continue;
      const HNode& nd = A.nodes[s];
      int t = t0[s];
      for (int c = 0; c < nd.nc; ++c) {
        A.top_steps.push_back({s, c, 0, t});
        t += cdiv(std::min(A.nb, nd.n - c * A.nb), IB);
      }
      if (nd.parent >= 0) t0[nd.parent] = std::max(t0[nd.parent], t);
    }
    std::stable_sort(A.top_steps.begin(), A.top_steps.end(),
                     [](const TopStep& a, const TopStep& b) { return a.slot < b.slot; });
    for (size_t t = 0; t < A.top_steps.size(); ++t) {
      TopStep& st = A.top_steps[t];
      st.owner = (int)(t % world);
      const int g = A.nodes[st.node].bcol0 + st.c;
      A.bcol_owner[g] = st.owner;
      A.bcol_step[g] = (int)t;
    }
  }
  // ---- arena layout, IDENTICAL on every rank (an arena offset names the same entry of L on every
  // GPU, so a peer address is peer_base + offset): the subtrees of rank 0, of rank 1, ..., then
  // the upper tree.  Every rank allocates the whole arena (7.8 GB for Poisson 100^3; HBM is 180 GB)
  // but only ever touches its own subtrees and the upper tree.
  i64 off = 0;
  A.own_begin = A.own_end = 0;
  const int npass = world > 1 ? world + 1 : 2;
  for (int pass = 0; pass < npass; ++pass) {
    const bool top_pass = pass == npass - 1;
    if (top_pass) A.top_begin = off;
    if (world > 1 && pass == rank) A.own_begin = off;
    for (int s = 0; s < nn; ++s) {
      HNode& nd = A.nodes[s];
      const bool take = world > 1 ? (top_pass ? nd.owner < 0 : nd.owner == pass) : ((nd.small == 0) == top_pass);
      if (!take) continue;
      nd.off = off;
      off += rup((i64)nd.m * nd.ld, 16);
    }
    if (world > 1 && pass == rank) A.own_end = off;
  }
  if (world <= 1) {
    A.own_begin = 0;
    A.own_end = off;
  }
  A.arena = off;
}

// ------------------------------------------------------------------------------------------
// Solve schedule: the same block-column level sets.  Forward: diag solve of the block
// column, then x[index[r]] -= L[r, bcol] * x_bcol for the rows below.  Backward: the reverse.
static void build_pipe_schedule(Analysis& A) {
  // Persistent-kernel solve: tasks in a topological order of the assembly tree, (depth0, node)
  // ascending for the forward sweep, the reverse for the backward sweep.  A task only ever
  // waits on tasks that precede it in its list, so claiming tasks in list order from running
  // CTAs cannot deadlock (tests/test_symbolic.py replays the lists and checks exactly that).
  const int nn = A.nnodes, cut = A.solve_cut;
  A.pnodes.assign(nn, PNode{});
  A.ptasks_f.clear();
  A.ptasks_b.clear();
  A.pipe_dest.clear();
  int strip = 0;
  for (int s = 0; s < nn; ++s) {
    const HNode& nd = A.nodes[s];
    PNode& p = A.pnodes[s];
    p.off = nd.off;
    p.idx_off = nd.idx_off;
    p.ld = nd.ld;
    p.m = nd.m;
    p.n = nd.n;
    p.sa = nd.sa;
    p.strip0 = strip;
    p.np = (nd.n + PS - 1) / PS;
    p.expect_f = p.expect_b = 0;
    p.pflag = -1;
    strip += p.np;
  }
  A.nstrips = strip;
  A.strip_node.assign(strip, 0);
  for (int s = 0; s < nn; ++s)
    for (int i = 0; i < A.pnodes[s].np; ++i) A.strip_node[A.pnodes[s].strip0 + i] = s;
  for (int s = 0; s < nn; ++s)
    if (A.nodes[s].parent >= 0) A.pnodes[s].pflag = A.pnodes[A.nodes[s].parent].strip0;
  // Which nodes go into which pair of lists.  Single GPU: one pair holding every node at or
  // above the cut.  Multi-GPU: list 0 = the subtrees this rank owns, list 1 = the
  // shared upper tree (solved redundantly on every rank between two all-reduces of the work
  // vector); each list has its own per-strip expected counts -- a contribution that crosses from
  // a subtree into the upper tree is complete when the all-reduce is, so it is not counted.
  const bool multi = A.world > 1;
  A.pexpect.assign(strip, 0);
  A.pexpect_top.assign(strip, 0);
  A.ptasks_ft.clear();
  A.ptasks_bt.clear();
  const i64 level_tasks = getenv("SPLLT_B200_PIPE_LEVEL_TASKS") ? atoi(getenv("SPLLT_B200_PIPE_LEVEL_TASKS")) : PIPE_LEVEL_TASKS;
  const i64 task_bytes = getenv("SPLLT_B200_PIPE_TASK_KB") ? 1024 * (i64)atoi(getenv("SPLLT_B200_PIPE_TASK_KB")) : PIPE_TASK_BYTES;
  auto is_small = [&](const HNode& nd) { return nd.n <= PS && nd.m - nd.n <= PIPE_SMALL_ROWS; };
  const int crit_rows = std::max(8, getenv("SPLLT_B200_PIPE_CRIT_ROWS") ? atoi(getenv("SPLLT_B200_PIPE_CRIT_ROWS")) / 8 * 8 : PS);
  const int crit_np = getenv("SPLLT_B200_PIPE_CRIT_NP") ? atoi(getenv("SPLLT_B200_PIPE_CRIT_NP")) : 3;
  for (int list = 0; list < (multi ? 2 : 1); ++list) {
    std::vector<PTask>& TF = list == 0 ? A.ptasks_f : A.ptasks_ft;
    std::vector<PTask>& TB = list == 0 ? A.ptasks_b : A.ptasks_bt;
    std::vector<int>& EX = list == 0 ? A.pexpect : A.pexpect_top;
    std::vector<char> in_list(nn, 0);
    std::vector<int> ord;
    for (int s = 0; s < nn; ++s) {
      const bool take = multi ? (list == 0 ? A.nodes[s].owner == A.rank : A.nodes[s].owner < 0) : A.nodes[s].depth0 >= cut;
      if (take) {
        ord.push_back(s);
        in_list[s] = 1;
      }
    }
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return A.nodes[a].depth0 < A.nodes[b].depth0; });
    auto dests = [&](int s, int r0, int r1, int* begin, int* count) {
      // distinct ancestor STRIPS owning rows [r0, r1) of node s (index is sorted => runs).
      // forward: the strips whose counters the task bumps; backward: the strips it waits for.
      *begin = (int)A.pipe_dest.size();
      const int* idx = A.index.data() + A.nodes[s].idx_off;
      int last = -1;
      for (int r = r0; r < r1; ++r) {
        const int d = A.col2node[idx[r]];
        const int st = A.pnodes[d].strip0 + (idx[r] - A.nodes[d].sa) / PS;
        if (st != last) {
          A.pipe_dest.push_back(st);
          if (in_list[d]) EX[st]++;
          A.pnodes[d].expect_f++;
          last = st;
        }
      }
      *count = (int)A.pipe_dest.size() - *begin;
    };
    // rows per BELOW task of a narrow node: large chunks where a tree level has many nodes, small
    // ones near the top where a few tall nodes must be spread over the whole device
    std::vector<i64> below_rows(A.ndepth + 1, 0);
    for (int s : ord)
      if (!is_small(A.nodes[s]) && A.pnodes[s].np <= PIPE_FAT_NP) below_rows[A.nodes[s].depth0] += A.nodes[s].m - A.nodes[s].n;
    auto chunk_of = [&](int s) {
      if (A.pnodes[s].np > PIPE_FAT_NP) return PS;
      const i64 by_level = (below_rows[A.nodes[s].depth0] / std::max<i64>(level_tasks, 1) + PS - 1) / PS * PS;
      const i64 by_bytes = (task_bytes / (8 * (i64)A.nodes[s].n) + PS - 1) / PS * PS;   // enough bytes per task
      return (int)std::max<i64>(PS, std::min<i64>(PIPE_FAT_ROWS, std::max(by_level, by_bytes)));
    };
    for (int s : ord) {
      const HNode& nd = A.nodes[s];
      if (is_small(nd)) {
        PTask t{s, P_SMALL, 0, 0, 0, 0, {0, 0}};
        dests(s, nd.n, nd.m, &t.dest_begin, &t.dest_count);
        TF.push_back(t);
        continue;
      }
      for (int i = 0; i < A.pnodes[s].np; ++i) TF.push_back(PTask{s, P_DIAG, i, 0, 0, 0, {0, 0}});
      // the first 64 rows are their own task: they feed the parent's first strip (forward) and are
      // the last to become ready (backward), i.e. they sit on the critical path of a chain of nodes
      // ... and so do all rows that map to the parent (backward: none of them can start before the
      // parent is complete), so those are cut into 64-row tasks that run side by side
      const int chunk = chunk_of(s);
      int crit_end = nd.n + PS;
      if (nd.parent >= 0) {
        const int* idx = A.index.data() + nd.idx_off;
        const int pend = A.nodes[nd.parent].en;
        while (crit_end < nd.m && idx[crit_end] <= pend) ++crit_end;
      }
      for (int r = nd.n; r < nd.m;) {
        // Experiment (SPLLT_B200_PIPE_CRIT_ROWS < 64): backward, the node's first 64 rows below map to
        // the parent's LOWEST strips, the last ones to be published; cut into smaller tasks that chunk
        // is streamed by several CTAs at once.  Measured: no effect (Poisson 64^3 2.09 vs 2.06 ms,
        // 100^3 9.54 vs 9.65 ms) -- the default keeps 64-row tasks.
        int rows = std::min(r < crit_end ? PS : chunk, nd.m - r);
        if (r < nd.n + PS && A.pnodes[s].np >= crit_np) rows = std::min(rows, std::min(crit_rows, nd.n + PS - r));
        PTask t{s, P_BELOW, r, rows, 0, 0, {0, 0}};
        dests(s, r, r + t.nrows, &t.dest_begin, &t.dest_count);
        TF.push_back(t);
        r += rows;
      }
    }
    // the backward tasks reuse the forward tasks' strip lists (same row ranges)
    std::vector<std::vector<PTask>> of_node(nn);
    for (const PTask& t : TF)
      if (t.kind != P_DIAG) of_node[t.node].push_back(t);
    for (auto it = ord.rbegin(); it != ord.rend(); ++it) {
      const int s = *it;
      const HNode& nd = A.nodes[s];
      if (is_small(nd)) {
        TB.push_back(of_node[s][0]);
        continue;
      }
      // bottom chunk first: its rows belong to the highest ancestors, which finish first
      for (auto t = of_node[s].rbegin(); t != of_node[s].rend(); ++t) {
        TB.push_back(*t);
        A.pnodes[s].expect_b++;
      }
      for (int i = A.pnodes[s].np - 1; i >= 0; --i) TB.push_back(PTask{s, P_DIAG, i, 0, 0, 0, {0, 0}});
    }
    // EARLIEST-START ORDER (default on one GPU; SPLLT_B200_PIPE_ORDER_EST=0 / 1 forces it off / on): both
    // lists are sorted by a modelled earliest start time of every task -- the longest path through its
    // dependencies with rough task costs (2.3 us per strip, hand-off + bytes at 25 GB/s per below task).
    // CTAs claim tasks in list order, so this is list scheduling: a task is claimed about when its
    // inputs exist, and the slots are not held by tasks that poll while ready work sits further down
    // the list (with the (depth, node) order the trace showed 43 % of the CTA-time of a backward sweep
    // in polling below tasks).  The order stays topological: in the model every dependency ends before
    // its dependant starts (tests/test_pipe_schedule.py replays and executes the lists).  Measured:
    // Poisson 100^3 solve 9.6 -> 7.5 ms (forward 3.7 -> 3.0, backward 5.8 -> 4.5), 64^3 2.13 -> 1.55 ms.
    // The multi-GPU lists keep the (depth, node) order (only the CPU replay has seen them sorted).
    const char* est_env = getenv("SPLLT_B200_PIPE_ORDER_EST");
    if (est_env ? atoi(est_env) != 0 : !multi) {
      auto below_cost = [&](const PTask& t) {
        const HNode& nd = A.nodes[t.node];
        const double rows = t.kind == P_SMALL ? nd.m : t.nrows;
        return 4.0 + rows * nd.n * 8.0 / 25000.0;       // us: hand-off + streaming at ~25 GB/s per CTA
      };
      const double strip_cost = 2.3;
      auto sort_by = [&](std::vector<PTask>& T, const std::vector<double>& est) {
        std::vector<int> perm(T.size());
        std::iota(perm.begin(), perm.end(), 0);
        std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return est[a] < est[b]; });
        std::vector<PTask> out;
        out.reserve(T.size());
        for (int k : perm) out.push_back(T[k]);
        T.swap(out);
      };
      {   // forward: a strip starts after the previous strip of its node and after every contribution
        std::vector<double> est(TF.size(), 0.0), contrib(strip, 0.0), strip_end(strip, 0.0), last_start(nn, 0.0);
        for (size_t k = 0; k < TF.size(); ++k) {
          const PTask& t = TF[k];
          const PNode& pn = A.pnodes[t.node];
          if (t.kind == P_DIAG) {
            const int sidx = pn.strip0 + t.r0;
            est[k] = std::max(contrib[sidx], t.r0 > 0 ? strip_end[sidx - 1] : 0.0);
            strip_end[sidx] = est[k] + strip_cost;
            last_start[t.node] = est[k];
          } else {
            double e0;
            if (t.kind == P_SMALL) e0 = contrib[pn.strip0];
            else e0 = last_start[t.node] + 0.01;           // behind ALL strip tasks of its node in the list
            est[k] = e0;
            const double end = (t.kind == P_SMALL ? e0 : std::max(e0, strip_end[pn.strip0 + pn.np - 1])) + below_cost(t);
            for (int q = 0; q < t.dest_count; ++q) {
              const int d = A.pipe_dest[t.dest_begin + q];
              contrib[d] = std::max(contrib[d], end);
            }
          }
        }
        sort_by(TF, est);
      }
      {   // backward: a below task starts when the ancestor strips it reads are published, the last
          // strip of a node after all its below tasks, strip i after strip i + 1
        std::vector<double> est(TB.size(), 0.0), strip_end(strip, 0.0), below_end(nn, 0.0);
        for (size_t k = 0; k < TB.size(); ++k) {
          const PTask& t = TB[k];
          const PNode& pn = A.pnodes[t.node];
          if (t.kind == P_DIAG) {
            const int sidx = pn.strip0 + t.r0;
            est[k] = t.r0 == pn.np - 1 ? below_end[t.node] : strip_end[sidx + 1];
            strip_end[sidx] = est[k] + strip_cost;
          } else {
            double e0 = 0.0;
            for (int q = 0; q < t.dest_count; ++q) e0 = std::max(e0, strip_end[A.pipe_dest[t.dest_begin + q]]);
            est[k] = e0;
            const double end = e0 + below_cost(t);
            if (t.kind == P_SMALL) strip_end[pn.strip0] = end;
            else below_end[t.node] = std::max(below_end[t.node], end);
          }
        }
        sort_by(TB, est);
      }
    } else {
    // Readiness order (default on one GPU; SPLLT_B200_PIPE_BWD_EARLY=0 / 1 forces it off / on).  Backward, a BELOW task only
    // needs the x of the ancestor strips its rows map to -- usually long before its node's turn in the
    // list.  Claimed at its node's position it would find its inputs ready, but the task trace shows
    // the claim front lagging: every CTA slot holds a task that is still waiting (43 % of the CTA-time
    // of a Poisson 100^3 backward sweep is BELOW tasks polling).  Here every BELOW task is moved right
    // behind the LAST strip task it depends on (the order stays topological: producers first), so it is
    // claimed when its inputs exist and the slots go to tasks that can run.  Measured: Poisson 64^3
    // backward sweep 1.24 -> 0.95 ms (solve 2.13 -> 1.75 ms), Poisson 100^3 unchanged (5.8 -> 5.7 ms).
    // The multi-GPU lists keep the node order (validated only by the CPU replay so far).
    const char* early = getenv("SPLLT_B200_PIPE_BWD_EARLY");
    if (early ? atoi(early) != 0 : !multi) {
      std::vector<int> strip_pos(strip, -1);       // position in TB of the DIAG / SMALL task that publishes a strip
      for (size_t k = 0; k < TB.size(); ++k) {
        const PTask& t = TB[k];
        if (t.kind == P_DIAG) strip_pos[A.pnodes[t.node].strip0 + t.r0] = (int)k;
        else if (t.kind == P_SMALL) strip_pos[A.pnodes[t.node].strip0] = (int)k;
      }
      std::vector<std::pair<int, int>> key(TB.size());   // (anchor position, original position)
      for (size_t k = 0; k < TB.size(); ++k) {
        const PTask& t = TB[k];
        int anchor = (int)k;                              // strips and fused nodes stay where they are
        if (t.kind == P_BELOW) {
          anchor = -1;                                    // nothing in this list to wait for: front of the list
          for (int q = 0; q < t.dest_count; ++q) anchor = std::max(anchor, strip_pos[A.pipe_dest[t.dest_begin + q]]);
        }
        key[k] = {anchor, (int)k};
      }
      std::vector<int> perm(TB.size());
      std::iota(perm.begin(), perm.end(), 0);
      std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) {
        if (key[a].first != key[b].first) return key[a].first < key[b].first;
        // same anchor: the anchor task itself first, then the tasks that wait for it, in list order
        const bool aa = key[a].second == key[a].first, bb = key[b].second == key[b].first;
        if (aa != bb) return aa;
        return key[a].second < key[b].second;
      });
      std::vector<PTask> sorted;
      sorted.reserve(TB.size());
      for (int k : perm) sorted.push_back(TB[k]);
      TB.swap(sorted);
    }
    }
  }
  // multi-GPU: pivot columns this rank keeps when the work vector is summed over the ranks
  A.col_keep.assign(A.n, 1);
  if (multi)
    for (int s = 0; s < nn; ++s) {
      const int own = A.nodes[s].owner;
      const char keep = own == A.rank || (own < 0 && A.rank == 0);
      for (int c = A.nodes[s].sa; c <= A.nodes[s].en; ++c) A.col_keep[c] = keep;
    }
}

void build_solve_schedule(Analysis& A) {
  const int nn = A.nnodes, nb = A.nb;
  // SPLLT_B200_SOLVE_CUT = d: nodes whose first block column sits at schedule depth < d are
  // solved by the level-set launches, the rest by the persistent pipelined kernels.
  // default 0 = everything pipelined; SPLLT_B200_SOLVE_LEVELSET=1 = everything level-set.
  {
    const char* e = getenv("SPLLT_B200_SOLVE_CUT");
    A.solve_cut = e ? atoi(e) : 0;
    const char* l = getenv("SPLLT_B200_SOLVE_LEVELSET");
    if (l && atoi(l)) A.solve_cut = 1 << 30;
    if (A.world > 1) A.solve_cut = 0;   // the multi-GPU solve runs entirely in the persistent kernels
  }
  {
    // Which path serves how many right-hand sides.  The persistent kernels win where the sweep is
    // bound by the dependency chain of the assembly tree (few right-hand sides, most of L in
    // narrow nodes); the level-set update kernels stream the rows below WIDE nodes better (more
    // resident warps), so matrices whose factor sits mostly in wide nodes (large 3D fronts, e.g.
    // the elasticity configuration) keep the level-set launches.  SPLLT_B200_PIPE_MAX_NRHS overrides.
    double wide = 0, total = 0;
    for (int s = 0; s < nn; ++s) {
      const double e = (double)A.nodes[s].m * A.nodes[s].n;
      total += e;
      if (A.nodes[s].n > PIPE_FAT_NP * PS) wide += e;
    }
    A.wide_frac = total > 0 ? wide / total : 0;
    const char* e = getenv("SPLLT_B200_PIPE_MAX_NRHS");
    A.pipe_max_nrhs = e ? atoi(e) : (A.wide_frac > PIPE_WIDE_FRAC_MAX ? 0 : 8);
  }
  A.sbcols.clear();
  A.supds.clear();
  A.supds_t.clear();
  // two level-set schedules in the same lists: the nodes below the cut (companion of the
  // persistent kernels) and all nodes (many right-hand sides)
  for (int pass = 0; pass < 2; ++pass) {
    std::vector<SolveLaunch>& SL = pass == 0 ? A.slaunch : A.slaunch_full;
    const int cut = pass == 0 ? A.solve_cut : (1 << 30);
    SL.assign(A.ndepth, SolveLaunch{0, 0, 0, 0, 0, 0});
    std::vector<std::vector<int>> at(A.ndepth);
    for (int s = 0; s < nn; ++s)
      if (A.nodes[s].depth0 < cut)
        for (int c = 0; c < A.nodes[s].nc; ++c) at[A.nodes[s].depth0 + c].push_back(A.nodes[s].bcol0 + c);
    for (int d = 0; d < A.ndepth; ++d) {
      SolveLaunch& L = SL[d];
      L.diag_begin = A.sbcols.size();
      L.upd_begin = A.supds.size();
      L.updt_begin = A.supds_t.size();
      for (int g : at[d]) {
        const HNode& nd = A.nodes[A.bcol_node[g]];
        SolveBcol b;
        b.off = nd.off;
        b.idx_off = nd.idx_off;
        b.ld = nd.ld;
        b.m = nd.m;
        b.r0 = A.bcol_c[g] * nb;
        b.w = std::min(nb, nd.n - b.r0);
        b.sa = nd.sa;
        b.pad = 0;
        int id = (int)A.sbcols.size();
        A.sbcols.push_back(b);
        for (int r = b.r0 + b.w; r < nd.m; r += SOLVE_ROWS)
          A.supds.push_back({id, r, std::min(SOLVE_ROWS, nd.m - r), 0});
        for (int r = b.r0 + b.w; r < nd.m; r += SOLVE_ROWS_T)
          for (int k0 = 0; k0 < b.w; k0 += 64) A.supds_t.push_back({id, r, std::min(SOLVE_ROWS_T, nd.m - r), k0});
      }
      L.diag_count = (i64)A.sbcols.size() - L.diag_begin;
      L.upd_count = (i64)A.supds.size() - L.upd_begin;
      L.updt_count = (i64)A.supds_t.size() - L.updt_begin;
    }
  }
  build_pipe_schedule(A);
}

}  // namespace spllt
