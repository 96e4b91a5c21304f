// host_seeds.cu -- the two small host steps around seed selection: the sorted /exploration_tree_rows_info of
// aos_seed_gen_node (src/aos_seed_gen_node.cpp:2546-2582) and the greedy 0.5 m merge of
// aos_gvd_node::voronoiSeedsCallback (src/aos_gvd_node.cpp:84-128).  The ray casts and first-come filters of
// seed selection itself run on the device (k_seeds.cu).  Same arithmetic as the reference, different data
// structures: the reference's O(S^2) scan runs on a uniform hash grid, which returns the same answers.
#include <math.h>

#include <algorithm>

#include "aos_common.cuh"

namespace aos {

// publishExplorationTreeRowsInfoFromClusters (seed_gen:2546-2582): sort by centre (y, then x if |dy| < 1e-6);
// stable, so rows with equal keys keep cluster order
void host_rows_info(const std::vector<aos_tree_row> &rows, std::vector<double> *rows_info) {
  std::vector<int> ord(rows.size());
  for (size_t i = 0; i < rows.size(); ++i) ord[i] = (int)i;
  std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) {
    if (fabs(rows[a].center_y - rows[b].center_y) < 1e-6) return rows[a].center_x < rows[b].center_x;
    return rows[a].center_y < rows[b].center_y;
  });
  rows_info->clear();
  for (int i : ord) {
    rows_info->push_back(rows[i].start_x);
    rows_info->push_back(rows[i].start_y);
    rows_info->push_back(rows[i].end_x);
    rows_info->push_back(rows[i].end_y);
  }
}

// voronoiSeedsCallback, gvd:84-128: greedy leader clustering at 0.5 m (<=), centroid in index order.
// Same answers as the reference's O(S^2) scan; the neighbour candidates come from a flat open-addressing grid
// (cell = merge distance) instead.
void host_merge_seeds(const double *seeds, int n, std::vector<double> *out) {
  const double merge_distance = 0.5;
  out->clear();
  if (n <= 0) return;
  size_t cap = 64;
  while (cap < (size_t)n * 2) cap <<= 1;
  const size_t mask = cap - 1;
  std::vector<unsigned long long> keys(cap, ~0ull);
  std::vector<int> head(cap, -1), next((size_t)n, -1);
  std::vector<char> used((size_t)n, 0), finite((size_t)n, 1);
  auto cell = [&](double v) { return (long long)floor(v / merge_distance); };
  auto key = [](long long cx, long long cy) {
    return ((unsigned long long)(cx + (1ll << 30)) << 32) ^ (unsigned long long)(cy + (1ll << 30));
  };
  auto slot_of = [&](unsigned long long k, bool insert) -> long long {
    unsigned long long h = k;
    h ^= h >> 33;
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33;
    size_t i = (size_t)h & mask;
    for (;;) {
      if (keys[i] == k) return (long long)i;
      if (keys[i] == ~0ull) {
        if (!insert) return -1;
        keys[i] = k;
        return (long long)i;
      }
      i = (i + 1) & mask;
    }
  };
  for (int i = 0; i < n; ++i) {
    if (!std::isfinite(seeds[2 * i]) || !std::isfinite(seeds[2 * i + 1])) {
      finite[i] = 0;  // a non-finite seed never merges with anything (norm is NaN/inf) and is dropped at gvd:266
      continue;
    }
    long long s = slot_of(key(cell(seeds[2 * i]), cell(seeds[2 * i + 1])), true);
    next[i] = head[s];
    head[s] = i;
  }
  std::vector<int> members;
  for (int i = 0; i < n; ++i) {
    if (used[i]) continue;
    used[i] = 1;
    if (!finite[i]) continue;  // its own cluster, filtered out by processGraph (gvd:266-270)
    members.clear();
    const long long cx = cell(seeds[2 * i]), cy = cell(seeds[2 * i + 1]);
    for (long long dy = -1; dy <= 1; ++dy)
      for (long long dx = -1; dx <= 1; ++dx) {
        long long s = slot_of(key(cx + dx, cy + dy), false);
        if (s < 0) continue;
        for (int j = head[s]; j >= 0; j = next[j]) {
          if (j <= i || used[j]) continue;
          double ex = seeds[2 * i] - seeds[2 * j], ey = seeds[2 * i + 1] - seeds[2 * j + 1];
          if (sqrt(ex * ex + ey * ey) <= merge_distance) members.push_back(j);
        }
      }
    std::sort(members.begin(), members.end());
    double sx = seeds[2 * i], sy = seeds[2 * i + 1];
    for (int j : members) {
      used[j] = 1;
      sx += seeds[2 * j];
      sy += seeds[2 * j + 1];
    }
    double cnt = (double)(members.size() + 1);
    out->push_back(sx / cnt);
    out->push_back(sy / cnt);
  }
}

}  // namespace aos
