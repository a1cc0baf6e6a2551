#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <chrono>
extern "C" int METIS_NodeND(int64_t*, int64_t*, int64_t*, int64_t*, int64_t*, int64_t*, int64_t*);
int main(int argc,char**argv){
  int N = argc>1? atoi(argv[1]):20;
  int64_t n=(int64_t)N*N*N; std::vector<int64_t> xadj(n+1), adj; adj.reserve(n*6);
  for(int z=0;z<N;z++)for(int y=0;y<N;y++)for(int x=0;x<N;x++){
    int64_t v=((int64_t)z*N+y)*N+x; xadj[v]=adj.size();
    if(x>0)adj.push_back(v-1); if(x<N-1)adj.push_back(v+1);
    if(y>0)adj.push_back(v-N); if(y<N-1)adj.push_back(v+N);
    if(z>0)adj.push_back(v-(int64_t)N*N); if(z<N-1)adj.push_back(v+(int64_t)N*N);
  } xadj[n]=adj.size();
  std::vector<int64_t> perm(n), iperm(n);
  auto t0=std::chrono::steady_clock::now();
  int rc=METIS_NodeND(&n,xadj.data(),adj.data(),nullptr,nullptr,perm.data(),iperm.data());
  auto t1=std::chrono::steady_clock::now();
  std::vector<char> seen(n,0); bool ok=true; for(auto p:perm){ if(p<0||p>=n||seen[p]) ok=false; else seen[p]=1;}
  for(int64_t i=0;i<n&&ok;i++) if(perm[iperm[i]]!=i) ok=false;
  printf("rc=%d valid=%d time=%.2fs\n",rc,(int)ok,std::chrono::duration<double>(t1-t0).count());
}
