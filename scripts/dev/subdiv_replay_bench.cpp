// Stand-alone timing of the Subdiv2D insertion replay (csrc/host_subdiv.cu) on the seeds of a C3 map, host only:
//   python scripts/dev/dump_c3.py                      # on a GPU box: writes gpurun_out/c3_seeds.npy
//   python -c "import numpy as np; np.load('gpurun_out/c3_seeds.npy').astype('f8').tofile('scripts/dev/sd_bench/seeds.bin')"
//   g++ -O3 -std=c++17 -ffp-contract=off -Iactive-orchard-slam_b200/csrc -x c++ -o scripts/dev/sd_bench/cur \
//       scripts/dev/subdiv_replay_bench.cpp active-orchard-slam_b200/csrc/host_subdiv.cu
//   (cd scripts/dev/sd_bench && ./cur)                 # 7 runs: insert ms, facets ms, FNV hash of the facet vertices
// The hash must not change between variants (bit-identical facets); the container's CPU is too noisy to rank
// variants within 10 %, the GPU box's host is not (gpurun -- 'cd scripts/dev/sd_bench; ./A; ./B').
#include "host_subdiv.h"
#include <string.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <vector>
using namespace aos;
long g_flip=0,g_pred=0,g_loc=0,g_conn=0,g_fallback=0;
int main(int argc,char**argv){
  FILE*f=fopen("seeds.bin","rb"); std::vector<double> s(2*238898); size_t n=fread(s.data(),16,238898,f); fclose(f);
  double minx=1e30,maxx=-1e30,miny=1e30,maxy=-1e30;
  for(size_t i=0;i<n;i++){minx=std::min(minx,s[2*i]);maxx=std::max(maxx,s[2*i]);miny=std::min(miny,s[2*i+1]);maxy=std::max(maxy,s[2*i+1]);}
  minx-=1;maxx+=1;miny-=1;maxy+=1;
  const float rx = (float)(minx - 1.0), ry = (float)(miny - 1.0);
  const float rw = (float)(fabs(maxx - minx) + 2.0), rh = (float)(fabs(maxy - miny) + 2.0);
  if (argc > 1 && !strcmp(argv[1], "scalar")) aos::g_subdiv_simd = 0;  // the scalar flip loop instead of the AVX2 one
  for(int it=0;it<7;it++){
    g_flip=g_pred=g_loc=g_conn=0;
    auto t0=std::chrono::steady_clock::now();
    Subdiv sd; sd.reserve(n); sd.init((int)lrint(rx),(int)lrint(ry),(int)lrint(rw),(int)lrint(rh));
    for(size_t i=0;i<n;i++){ float x=(float)s[2*i],y=(float)s[2*i+1];
      x = std::max(rx + .1f, std::min(rx + rw - .1f, x)); y = std::max(ry + .1f, std::min(ry + rh - .1f, y)); sd.insert(x,y);}
    auto t1=std::chrono::steady_clock::now();
    std::vector<float> xy; std::vector<int32_t> off; sd.voronoi_facets(&xy,&off);
    auto t2=std::chrono::steady_clock::now();
    unsigned long h=1469598103934665603ul; for(float v:xy){unsigned u; memcpy(&u,&v,4); h=(h^u)*1099511628211ul;}
    printf("insert %.1f ms facets %.1f ms  hash %lx  flips %ld pred %ld loc %ld conn %ld\n",
      std::chrono::duration<double,std::milli>(t1-t0).count(),std::chrono::duration<double,std::milli>(t2-t1).count(),h,g_flip,g_pred,g_loc,g_conn);
  }
}
