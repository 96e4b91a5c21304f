// dev_hash.cuh -- device-side open-addressing hash (64-bit key -> slot with an int payload) and the uniform
// point grid built on it (cell -> head of a linked list of point indices).  Shared by k_graph.cu and k_seeds.cu.
#pragma once
#include <stdint.h>

namespace aos {

// ---------------------------------------------------------------------------------------------------
// open-addressing hash: 64-bit key -> slot with an int payload
// ---------------------------------------------------------------------------------------------------
struct DevHash {
  unsigned long long *keys;  // kEmptyKey when free
  int *val;
  unsigned mask;
};
constexpr unsigned long long kEmptyKey = ~0ull;

static __device__ __forceinline__ unsigned hmix(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return (unsigned)k;
}
static __device__ __forceinline__ int hash_insert(const DevHash &h, unsigned long long key) {
  unsigned i = hmix(key) & h.mask;
  for (;;) {
    unsigned long long k = h.keys[i];
    if (k == key) return (int)i;
    if (k == kEmptyKey) {
      unsigned long long old = atomicCAS(&h.keys[i], kEmptyKey, key);
      if (old == kEmptyKey || old == key) return (int)i;
    }
    i = (i + 1) & h.mask;
  }
}
static __device__ __forceinline__ int hash_find(const DevHash &h, unsigned long long key) {
  unsigned i = hmix(key) & h.mask;
  for (;;) {
    unsigned long long k = h.keys[i];
    if (k == key) return (int)i;
    if (k == kEmptyKey) return -1;
    i = (i + 1) & h.mask;
  }
}

// uniform grid over 2-D points: cell -> head of a linked list of point indices
struct PointGrid {
  DevHash h;   // val = list head (-1 = empty)
  int *next;   // per point
  double inv;  // 1 / cell size
};
static __device__ __forceinline__ long long cell_coord(double v, double inv) {
  double f = floor(v * inv);
  f = fmin(fmax(f, -1073741824.0), 1073741824.0);  // far-away Voronoi vertices share the border cells
  return (long long)f;
}
static __device__ __forceinline__ unsigned long long cell_key(long long cx, long long cy) {
  return ((unsigned long long)(cx + (1ll << 31)) << 32) | (unsigned long long)(cy + (1ll << 31));
}


}  // namespace aos
