// Host-side symbolic analysis (ordering + supernodal symbolic factorization).
//
// Stand-in for the SPRAL `ssids_analyse` call that the reference makes at
// src/spllt_analyse_mod.F90:129-131 (ordering = METIS, nemin amalgamation).  SPRAL is
// a third-party dependency absent from /root/reference, so this is an independent
// implementation of the published algorithms (Liu's elimination tree, the
// Gilbert-Ng-Peyton column counts, relaxed supernode amalgamation with `nemin`).
// Outputs use the same conventions as the SSIDS arrays SpLLT consumes
// (src/spllt_analyse_mod.F90:155-158): all indices 1-based, pivot order.
#pragma once
#include <cstdint>
#include <vector>

namespace spllt {

struct Symbolic {
  int n = 0;
  int nnodes = 0;
  std::vector<int> order;      // order[i-1] = position of variable i in the pivot sequence (1-based)
  std::vector<int> sptr;       // [nnodes+1] first column of each supernode, 1-based
  std::vector<int> sparent;    // [nnodes]   parent supernode, 1-based; roots -> nnodes+1
  std::vector<int64_t> rptr;   // [nnodes+1] 1-based pointers into rlist
  std::vector<int> rlist;      // row indices (pivot order, 1-based); first (sptr[s+1]-sptr[s]) are the columns
  int64_t num_factor = 0;      // entries in L
  int64_t num_flops = 0;       // sum over columns of colcount^2
};

enum OrderingKind {
  ORDER_METIS = 1,      // nested dissection through the bundled METIS 5 (64-bit idx_t)
  ORDER_NATURAL = 0,    // identity
  ORDER_USER = 2        // take `order` as given on input
};

// n, ptr, row: lower triangle, CSC, 1-based (reference convention, example/C/simple.c:38-39).
// Returns 0 on success, <0 on error.
int symbolic_analyse(int n, const int* ptr, const int* row, int nemin, int ordering,
                     const int* user_order, Symbolic& out);

}  // namespace spllt
