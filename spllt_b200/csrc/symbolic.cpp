// Host-side symbolic analysis: see symbolic.h for what this replaces in the reference.
#include "symbolic.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>

// Bundled METIS 5 (cuSOLVER's libmetis_static.a) is built with 64-bit idx_t.
extern "C" int METIS_NodeND(int64_t* nvtxs, int64_t* xadj, int64_t* adjncy, int64_t* vwgt,
                            int64_t* options, int64_t* perm, int64_t* iperm);

namespace spllt {
namespace {

// Symmetric adjacency (no diagonal, no duplicates), 0-based.
struct Graph {
  int n = 0;
  std::vector<int64_t> xadj;
  std::vector<int> adj;
};

Graph build_graph(int n, const int* ptr, const int* row) {
  Graph g;
  g.n = n;
  std::vector<int64_t> cnt(n + 1, 0);
  for (int j = 0; j < n; ++j)
    for (int64_t p = ptr[j] - 1; p < ptr[j + 1] - 1; ++p) {
      int i = row[p] - 1;
      if (i == j || i < 0 || i >= n) continue;
      cnt[i + 1]++;
      cnt[j + 1]++;
    }
  g.xadj.assign(n + 1, 0);
  for (int i = 0; i < n; ++i) g.xadj[i + 1] = g.xadj[i] + cnt[i + 1];
  g.adj.resize(g.xadj[n]);
  std::vector<int64_t> fill(g.xadj.begin(), g.xadj.end() - 1);
  for (int j = 0; j < n; ++j)
    for (int64_t p = ptr[j] - 1; p < ptr[j + 1] - 1; ++p) {
      int i = row[p] - 1;
      if (i == j || i < 0 || i >= n) continue;
      g.adj[fill[i]++] = j;
      g.adj[fill[j]++] = i;
    }
  // sort + dedupe each list (an entry given in both triangles would appear twice)
  std::vector<int64_t> nx(n + 1, 0);
  int64_t w = 0;
  for (int i = 0; i < n; ++i) {
    int64_t b = g.xadj[i], e = g.xadj[i + 1];
    std::sort(g.adj.begin() + b, g.adj.begin() + e);
    int64_t start = w;
    for (int64_t p = b; p < e; ++p)
      if (w == start || g.adj[w - 1] != g.adj[p]) g.adj[w++] = g.adj[p];
    nx[i + 1] = w;
  }
  g.adj.resize(w);
  g.xadj = nx;
  return g;
}

// Relabel graph: vertex v becomes lab[v].
Graph relabel(const Graph& g, const std::vector<int>& lab) {
  Graph h;
  h.n = g.n;
  h.xadj.assign(g.n + 1, 0);
  for (int v = 0; v < g.n; ++v) h.xadj[lab[v] + 1] = g.xadj[v + 1] - g.xadj[v];
  for (int i = 0; i < g.n; ++i) h.xadj[i + 1] += h.xadj[i];
  h.adj.resize(g.adj.size());
  for (int v = 0; v < g.n; ++v) {
    int64_t o = h.xadj[lab[v]];
    for (int64_t p = g.xadj[v]; p < g.xadj[v + 1]; ++p) h.adj[o++] = lab[g.adj[p]];
    std::sort(h.adj.begin() + h.xadj[lab[v]], h.adj.begin() + o);
  }
  return h;
}

// Liu's elimination tree with path compression.
std::vector<int> etree(const Graph& g) {
  int n = g.n;
  std::vector<int> parent(n, -1), anc(n, -1);
  for (int j = 0; j < n; ++j) {
    for (int64_t p = g.xadj[j]; p < g.xadj[j + 1]; ++p) {
      int i = g.adj[p];
      if (i >= j) break;  // lists are sorted
      // climb from i to the current root of its tree, compressing to j
      while (i != -1 && i != j) {
        int nxt = anc[i];
        anc[i] = j;
        if (nxt == -1) parent[i] = j;
        i = nxt;
      }
    }
  }
  return parent;
}

// Postorder of a forest given by parent[] (children visited in ascending order).
std::vector<int> postorder(const std::vector<int>& parent) {
  int n = (int)parent.size();
  std::vector<int> head(n, -1), next(n, -1), post;
  post.reserve(n);
  for (int j = n - 1; j >= 0; --j)
    if (parent[j] != -1) {
      next[j] = head[parent[j]];
      head[parent[j]] = j;
    }
  std::vector<int> stack;
  for (int r = 0; r < n; ++r) {
    if (parent[r] != -1) continue;
    stack.push_back(r);
    while (!stack.empty()) {
      int v = stack.back();
      int c = head[v];
      if (c == -1) {
        post.push_back(v);
        stack.pop_back();
      } else {
        head[v] = next[c];
        stack.push_back(c);
      }
    }
  }
  return post;
}

// Column counts of L (diagonal included) for a graph whose vertices are already in
// etree postorder (parent[j] > j).  Skeleton-graph algorithm of Gilbert, Ng & Peyton.
std::vector<int64_t> column_counts(const Graph& g, const std::vector<int>& parent) {
  int n = g.n;
  std::vector<int> first_desc(n), sz(n, 1);
  for (int j = 0; j < n; ++j)
    if (parent[j] != -1) sz[parent[j]] += sz[j];
  for (int j = 0; j < n; ++j) first_desc[j] = j - sz[j] + 1;
  std::vector<int64_t> w(n);
  for (int j = 0; j < n; ++j) w[j] = (sz[j] == 1) ? 1 : 0;  // leaves start at 1
  std::vector<int> max_first(n, -1), prev_leaf(n, -1), uf(n);
  std::iota(uf.begin(), uf.end(), 0);
  auto find = [&](int x) {
    int r = x;
    while (uf[r] != r) r = uf[r];
    while (uf[x] != r) {
      int t = uf[x];
      uf[x] = r;
      x = t;
    }
    return r;
  };
  for (int j = 0; j < n; ++j) {
    if (parent[j] != -1) w[parent[j]]--;
    for (int64_t p = g.xadj[j]; p < g.xadj[j + 1]; ++p) {
      int i = g.adj[p];
      if (i <= j) continue;
      // is j a leaf of the row subtree of i ?
      if (first_desc[j] <= max_first[i]) continue;
      max_first[i] = first_desc[j];
      int pl = prev_leaf[i];
      prev_leaf[i] = j;
      w[j]++;
      if (pl != -1) w[find(pl)]--;  // least common ancestor of consecutive leaves
    }
    if (parent[j] != -1) uf[j] = parent[j];
  }
  for (int j = 0; j < n; ++j)
    if (parent[j] != -1) w[parent[j]] += w[j];
  return w;
}

}  // namespace

int symbolic_analyse(int n, const int* ptr, const int* row, int nemin, int ordering,
                     const int* user_order, Symbolic& out) {
  out = Symbolic();
  out.n = n;
  if (n <= 0) return 0;
  if (nemin < 1) nemin = 32;
  Graph g = build_graph(n, ptr, row);

  // ---- ordering: pos[v] = 0-based pivot position of variable v
  std::vector<int> pos(n);
  if (ordering == ORDER_METIS && n > 1 && !g.adj.empty()) {
    int64_t nv = n;
    std::vector<int64_t> xadj(g.xadj), adj(g.adj.begin(), g.adj.end()), perm(n), iperm(n);
    int rc = METIS_NodeND(&nv, xadj.data(), adj.data(), nullptr, nullptr, perm.data(), iperm.data());
    if (rc != 1) return -1;
    for (int v = 0; v < n; ++v) pos[v] = (int)iperm[v];
  } else if (ordering == ORDER_USER && user_order) {
    std::vector<char> seen(n, 0);
    for (int v = 0; v < n; ++v) {
      int p = user_order[v] - 1;
      if (p < 0 || p >= n || seen[p]) return -2;
      seen[p] = 1;
      pos[v] = p;
    }
  } else {
    std::iota(pos.begin(), pos.end(), 0);
  }

  // ---- elimination tree of the permuted matrix, then postorder it
  Graph gp = relabel(g, pos);
  std::vector<int> par0 = etree(gp);
  std::vector<int> post = postorder(par0);
  std::vector<int> lab(n);
  for (int k = 0; k < n; ++k) lab[post[k]] = k;
  for (int v = 0; v < n; ++v) pos[v] = lab[pos[v]];
  Graph gq = relabel(gp, lab);
  std::vector<int> parent(n, -1);
  for (int v = 0; v < n; ++v)
    if (par0[v] != -1) parent[lab[v]] = lab[par0[v]];
  gp = Graph();

  std::vector<int64_t> cc = column_counts(gq, parent);

  // ---- supernodes: a column group is merged into the group of its etree parent when
  //  (a) the merge creates no fill and the parent is still a single column, or
  //  (b) both groups eliminate fewer than nemin columns (relaxed amalgamation).
  std::vector<int> merged_into(n, -1);
  std::vector<int64_t> grows(cc);      // rows of the group whose top column is j
  std::vector<int> gelim(n, 1);        // columns eliminated by that group
  for (int j = 0; j < n; ++j) {
    int p = parent[j];
    if (p == -1) continue;
    bool nofill = (gelim[p] == 1) && (grows[p] == grows[j] - gelim[j]);
    bool relaxed = (gelim[p] < nemin) && (gelim[j] < nemin);
    if (nofill || relaxed) {
      merged_into[j] = p;
      grows[p] += gelim[j];
      gelim[p] += gelim[j];
    }
  }
  std::vector<int> top(n);
  for (int j = n - 1; j >= 0; --j) top[j] = (merged_into[j] == -1) ? j : top[merged_into[j]];
  // supernode forest over the tops
  std::vector<int> sn_of_top(n, -1), tops;
  for (int j = 0; j < n; ++j)
    if (top[j] == j) {
      sn_of_top[j] = (int)tops.size();
      tops.push_back(j);
    }
  int ns = (int)tops.size();
  std::vector<int> sparent0(ns, -1);
  for (int s = 0; s < ns; ++s) {
    int p = parent[tops[s]];
    if (p != -1) sparent0[s] = sn_of_top[top[p]];
  }
  std::vector<int> spost = postorder(sparent0);  // children ascending by top column
  std::vector<int> snew(ns);
  for (int k = 0; k < ns; ++k) snew[spost[k]] = k;
  // columns of each group in ascending order -> new contiguous numbering
  std::vector<int> ccount(ns, 0);
  for (int j = 0; j < n; ++j) ccount[snew[sn_of_top[top[j]]]]++;
  out.nnodes = ns;
  out.sptr.assign(ns + 1, 1);
  for (int s = 0; s < ns; ++s) out.sptr[s + 1] = out.sptr[s] + ccount[s];
  std::vector<int> nextcol(ns);
  for (int s = 0; s < ns; ++s) nextcol[s] = out.sptr[s] - 1;
  std::vector<int> newcol(n);
  for (int j = 0; j < n; ++j) newcol[j] = nextcol[snew[sn_of_top[top[j]]]]++;
  out.sparent.assign(ns, ns + 1);
  for (int s = 0; s < ns; ++s)
    if (sparent0[s] != -1) out.sparent[snew[s]] = snew[sparent0[s]] + 1;
  for (int v = 0; v < n; ++v) pos[v] = newcol[pos[v]];
  out.order.resize(n);
  for (int v = 0; v < n; ++v) out.order[v] = pos[v] + 1;

  // ---- row lists by supernodal symbolic factorization in the final numbering
  Graph gf = relabel(gq, newcol);
  gq = Graph();
  std::vector<int> mark(n, -1), chead(ns, -1), cnext(ns, -1), b;
  for (int s = ns - 1; s >= 0; --s) {
    int p = out.sparent[s] - 1;
    if (p < ns) {
      cnext[s] = chead[p];
      chead[p] = s;
    }
  }
  out.rptr.assign(ns + 1, 1);
  out.rlist.clear();
  int64_t nfac = 0, nflop = 0;
  for (int s = 0; s < ns; ++s) {
    int sa = out.sptr[s] - 1, en = out.sptr[s + 1] - 2;
    b.clear();
    for (int c = sa; c <= en; ++c)
      for (int64_t p = gf.xadj[c]; p < gf.xadj[c + 1]; ++p) {
        int r = gf.adj[p];
        if (r > en && mark[r] != s) {
          mark[r] = s;
          b.push_back(r);
        }
      }
    for (int c = chead[s]; c != -1; c = cnext[c]) {
      // rows of child c below its own pivot block (already final in rlist)
      int64_t cb = out.rptr[c] - 1 + (out.sptr[c + 1] - out.sptr[c]), ce = out.rptr[c + 1] - 1;
      for (int64_t q = cb; q < ce; ++q) {
        int r = out.rlist[q] - 1;
        if (r > en && mark[r] != s) {
          mark[r] = s;
          b.push_back(r);
        }
      }
    }
    std::sort(b.begin(), b.end());
    int64_t m = (int64_t)b.size() + (en - sa + 1);
    out.rptr[s + 1] = out.rptr[s] + m;
    for (int c = sa; c <= en; ++c) out.rlist.push_back(c + 1);
    for (int r : b) out.rlist.push_back(r + 1);
    for (int64_t k = 0; k < en - sa + 1; ++k) {
      nfac += m - k;
      nflop += (m - k) * (m - k);
    }
  }
  out.num_factor = nfac;
  out.num_flops = nflop;
  return 0;
}

}  // namespace spllt
