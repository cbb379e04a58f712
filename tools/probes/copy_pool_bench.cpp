#include "../../aerobulk_b200/csrc/ab_copy_pool.hpp"
#include <chrono>
#include <cstdio>
#include <vector>
int main(int argc, char** argv){
  int T = argc>1?atoi(argv[1]):8; bool nt = argc>2;
  auto& pool = *new abpool::CopyPool(T, nt);
  size_t n = 14u<<20; // doubles = 117MB
  std::vector<double> a(n,1.0), b(n,0.0);
  std::vector<abpool::CopyPiece> jobs; size_t step=16384;
  for(size_t o=0;o<n;o+=step) jobs.push_back({b.data()+o,a.data()+o,8*std::min(step,n-o)});
  double best=1e9;
  for(int r=0;r<10;++r){auto t0=std::chrono::steady_clock::now(); pool.run(jobs.data(),(int)jobs.size()); double ms=std::chrono::duration<double,std::milli>(std::chrono::steady_clock::now()-t0).count(); if(ms<best)best=ms;}
  printf("T=%d nt=%d: %.2f ms for %.0f MB -> %.1f GB/s\n",T,(int)nt,best,n*8/1e6,n*8/1e6/best);
}
