// extern "C" door to the SSIDS stand-in (ordering + symbolic factorization) for the ORACLE side.
//
// TEST INFRASTRUCTURE.  The reference obtains order / sptr / sparent / rptr / rlist from SPRAL's
// ssids_analyse (src/spllt_analyse_mod.F90:129-131); SPRAL is absent from /root/reference, so the
// product and the oracle share one independent implementation of that front end
// (spllt_b200/csrc/symbolic.cpp, SURVEY.md 8c: "the symbolic front end is shared by oracle and
// CUDA build").  This file only re-exports it from its own shared object, so that the CPU arm of
// bench.py (--impl reference) and the oracle tests can produce the symbolic inputs WITHOUT loading
// libspllt_b200.so.  Built by oracle/Makefile from the source where it lies.
#include <algorithm>

#include "../spllt_b200/csrc/symbolic.h"

using namespace spllt;

extern "C" {

void* ssi_analyse(int n, const int* ptr, const int* row, int nemin, int ordering) {
  Symbolic* s = new Symbolic();
  if (symbolic_analyse(n, ptr, row, nemin, ordering, nullptr, *s) != 0) {
    delete s;
    return nullptr;
  }
  return s;
}
int ssi_nnodes(void* h) { return ((Symbolic*)h)->nnodes; }
long long ssi_rlist_len(void* h) { return (long long)((Symbolic*)h)->rlist.size(); }
void ssi_get(void* h, int* order, int* sptr, int* sparent, long long* rptr, int* rlist) {
  const Symbolic& s = *(Symbolic*)h;
  std::copy(s.order.begin(), s.order.end(), order);
  std::copy(s.sptr.begin(), s.sptr.end(), sptr);
  std::copy(s.sparent.begin(), s.sparent.end(), sparent);
  std::copy(s.rptr.begin(), s.rptr.end(), rptr);
  std::copy(s.rlist.begin(), s.rlist.end(), rlist);
}
void ssi_free(void* h) { delete (Symbolic*)h; }

}  // extern "C"
