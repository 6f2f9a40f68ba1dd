#include <chrono>
#include <cstdio>
#include <vector>
#include <cstdint>
#include <cmath>
// Host tree-builder benchmark (no GPU needed): cfg5-sized grid, prints the time of repeated builds.
//   g++ -O3 -std=c++17 -pthread -I. -o tools/builder_bench tools/builder_bench.cpp pymra_b200/csrc/mra_structure.cpp
#include "include/pymra_b200.h"
int main(){
  int n=2000; int64_t N=(int64_t)n*n; std::vector<double> locs(2*N);
  for(int iy=0;iy<n;++iy)for(int ix=0;ix<n;++ix){locs[2*((int64_t)iy*n+ix)]=(ix+1.0)/n; locs[2*((int64_t)iy*n+ix)+1]=(iy+1.0)/n;}
  int r=64,M=7,J=4,cd=8; int max_nodes=21845+10;
  std::vector<uint32_t> key(624); for(int i=0;i<624;++i) key[i]=1812433253u*i+12345; int pos=624;
  std::vector<int32_t> lvl(max_nodes),par(max_nodes),kind(max_nodes),cst(max_nodes),ccnt(max_nodes),dfs(max_nodes),kloc((size_t)max_nodes*r);
  std::vector<int64_t> rs(max_nodes),rc(max_nodes),koff(max_nodes),knots((size_t)max_nodes*r),perm(N);
  int nn,depth; int64_t nk;
  for(int rep=0;rep<6;++rep){
  auto t0=std::chrono::steady_clock::now();
  int rcode=mra_build_structure_2d(locs.data(),N,r,M,J,cd,key.data(),&pos,max_nodes,&nn,&depth,lvl.data(),par.data(),kind.data(),rs.data(),rc.data(),cst.data(),ccnt.data(),koff.data(),knots.data(),kloc.data(),&nk,perm.data(),dfs.data());
  auto t1=std::chrono::steady_clock::now();
  printf("rc=%d nodes=%d %.3f s\n",rcode,nn,std::chrono::duration<double>(t1-t0).count());}
}
