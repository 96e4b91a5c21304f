// k_edt.cu -- exact Euclidean distance transform with nearest-site labels on the bit-packed grid.
//
// Where an exact EDT legitimately enters this path: applyInflation
// (src/aos_seed_gen_node.cpp:933-967) is the threshold d^2 <= R^2 of the exact squared EDT of the raw occupancy
// grid (property-tested against the stencil kernel), and the declared-but-unfilled GvdGraph.edge_clearances
// (msg/GvdGraph.msg:58, gvd:856,890 publish 0.0f) is a distance-to-skeleton query -- offered as an opt-in
// (aos_set_clearance), default off so the published arrays stay bit-identical to the reference.
// aos_gvd_node's Voronoi diagram itself is cv::Subdiv2D's point-site diagram (host_subdiv.cu), not a raster.
//
// Exact, integer-only, the parallel banding algorithm (Cao et al., I3D 2010):
//   phase 1  per row   : nearest set cell of the same row for every cell (bit tricks inside a word, block-wide
//                        max/min scans across the words of the row)                      -> uint16 x of that cell
//   phase 2  per column: lower envelope of the parabolas (x - sx(y))^2 + (u - y)^2 by stack sweeps with exact
//                        integer separators (floor division): per band, then merged per column
//   phase 3  per band  : backward sweep over the merged stack writes the nearest site (x | y << 16) and d^2
// No approximation (this is not jump flooding); ties go to the site met first by the sweeps (lowest row, then
// lowest x), which tests/test_edt_gpu.py checks only through the distances, which are unique.
#include "aos_common.cuh"

namespace aos {

constexpr uint16_t kNoSite16 = 0xffffu;

// ---- phase 1: one block per row --------------------------------------------------------------------------
constexpr int kEdtRowThreads = 256;

__global__ void __launch_bounds__(kEdtRowThreads) edt_rows_kernel(const uint32_t *__restrict__ bits, int w, int h, int pitch,
                                                                  uint16_t *__restrict__ sx) {
  extern __shared__ int sm_edt[];  // hi[nw] then lo[nw]
  const int nw = (w + 31) >> 5;
  int *hi = sm_edt, *lo = sm_edt + nw;
  const int y = blockIdx.x;
  const uint32_t *row = bits + (size_t)y * pitch;
  // per word: highest / lowest set cell (global x), then running max (from the left) / min (from the right)
  for (int k = threadIdx.x; k < nw; k += kEdtRowThreads) {
    uint32_t v = row[k];
    if (k == nw - 1 && (w & 31)) v &= (1u << (w & 31)) - 1u;
    hi[k] = v ? (k << 5) + 31 - __clz(v) : -1;
    lo[k] = v ? (k << 5) + __ffs(v) - 1 : 0x7fffffff;
  }
  __syncthreads();
  // Hillis-Steele scans over nw (<= 2048) entries; nw is small, the row write-out below dominates
  for (int off = 1; off < nw; off <<= 1) {
    int vh[8], vl[8];
    int cnt = 0;
    for (int k = threadIdx.x; k < nw; k += kEdtRowThreads, ++cnt) {
      vh[cnt] = k >= off ? max(hi[k], hi[k - off]) : hi[k];
      vl[cnt] = k + off < nw ? min(lo[k], lo[k + off]) : lo[k];
    }
    __syncthreads();
    cnt = 0;
    for (int k = threadIdx.x; k < nw; k += kEdtRowThreads, ++cnt) {
      hi[k] = vh[cnt];
      lo[k] = vl[cnt];
    }
    __syncthreads();
  }
  // one thread per pair of cells, so a warp stores 128 contiguous bytes; the word and the two scan values it needs
  // come from shared memory (broadcast within the 16 lanes that share a word)
  uint32_t *out2 = reinterpret_cast<uint32_t *>(sx + (size_t)y * w);  // w * 2 bytes per row: 4-byte aligned iff w even
  const bool pair_ok = (w & 1) == 0;
  auto nearest_in_row = [&](int x) -> int {
    const int k = x >> 5, b = x & 31;
    uint32_t v = row[k];
    if (k == nw - 1 && (w & 31)) v &= (1u << (w & 31)) - 1u;
    const uint32_t le = v & (0xffffffffu >> (31 - b));  // bits <= b
    const uint32_t ge = v & (0xffffffffu << b);         // bits >= b
    const int L = le ? (k << 5) + 31 - __clz(le) : (k > 0 ? hi[k - 1] : -1);
    const int R = ge ? (k << 5) + __ffs(ge) - 1 : (k + 1 < nw ? lo[k + 1] : 0x7fffffff);
    if (L < 0 && R == 0x7fffffff) return kNoSite16;
    if (L < 0) return R;
    if (R == 0x7fffffff) return L;
    return (x - L <= R - x) ? L : R;  // ties go to the lower x
  };
  if (pair_ok) {
    for (int p = threadIdx.x; p < (w >> 1); p += kEdtRowThreads) {
      const int x = p << 1;
      out2[p] = (uint32_t)nearest_in_row(x) | ((uint32_t)nearest_in_row(x + 1) << 16);
    }
  } else {
    uint16_t *out = sx + (size_t)y * w;
    for (int x = threadIdx.x; x < w; x += kEdtRowThreads) out[x] = (uint16_t)nearest_in_row(x);
  }
}

// ---- phases 2 + 3: parallel banding over the columns ----------------------------------------------------------
// Every column is cut into bands of kEdtBandRows rows.  (2a) one thread per (column, band) sweeps its rows and keeps
// the sites of the band's own lower envelope on a stack; (2b) one thread per column merges the band stacks in row
// order into the column's envelope (only survivors are visited: a site dominated inside its band is dominated
// globally); (3) one thread per (column, band) locates its last row in the merged stack by binary search and fills
// nearest site and d^2 downwards.  Threads of a warp work on adjacent columns, so every plane access is coalesced.
constexpr int kEdtBandRows = 512;

__device__ __forceinline__ long long floor_div(long long a, long long b) {  // b > 0
  long long q = a / b;
  return (a % b != 0 && a < 0) ? q - 1 : q;
}

struct EnvTop {
  int q;                      // index of the top entry, -1 = empty
  long long s, t, g;          // its row, first row where it wins, squared x-distance
};

// push site (row u, squared x-distance gu) onto the envelope stack stored at plane rows base + q
__device__ __forceinline__ void env_push(uint16_t *__restrict__ S, uint16_t *__restrict__ T, const uint16_t *__restrict__ sx,
                                         size_t base, int w, int h, int x, EnvTop &e, int u, long long gu) {
  while (e.q >= 0) {
    const long long a = e.t - e.s, b = e.t - u;
    if (a * a + e.g > b * b + gu) {  // the new parabola already wins where the top one starts: pop
      --e.q;
      if (e.q >= 0) {
        e.s = S[(base + e.q) * w + x];
        e.t = T[(base + e.q) * w + x];
        const long long d = (long long)x - sx[(size_t)e.s * w + x];
        e.g = d * d;
      }
    } else {
      break;
    }
  }
  if (e.q < 0) {
    e.q = 0;
    e.s = u;
    e.t = 0;
    e.g = gu;
    S[base * w + x] = (uint16_t)u;
    T[base * w + x] = 0;
  } else {
    // first row where u beats the top: 1 + floor((u^2 - s^2 + g(u) - g(s)) / (2 (u - s)))
    const long long num = (long long)u * u - e.s * e.s + gu - e.g;
    const long long wrow = 1 + floor_div(num, 2 * ((long long)u - e.s));
    if (wrow < h) {
      ++e.q;
      e.s = u;
      e.t = wrow;
      e.g = gu;
      S[(base + e.q) * w + x] = (uint16_t)u;
      T[(base + e.q) * w + x] = (uint16_t)wrow;
    }
  }
}

__global__ void edt_band_sweep_kernel(const uint16_t *__restrict__ sx, int w, int h, uint16_t *__restrict__ S,
                                      uint16_t *__restrict__ T, uint16_t *__restrict__ counts) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, band = blockIdx.y;
  if (x >= w) return;
  const int r0 = band * kEdtBandRows, r1 = min(h, r0 + kEdtBandRows);
  EnvTop e{-1, 0, 0, 0};
  for (int u = r0; u < r1; ++u) {
    const uint16_t sxu = sx[(size_t)u * w + x];
    if (sxu == kNoSite16) continue;
    const long long dx = (long long)x - sxu;
    env_push(S, T, sx, (size_t)r0, w, h, x, e, u, dx * dx);
  }
  counts[(size_t)band * w + x] = (uint16_t)(e.q + 1);
}

// (2b) Junction merge.  The band stacks stay where they are; only the junction between the envelope built so far
// and the next band is repaired: tops that the band's first site beats at their start row are popped, leading
// sites of the band that never win against the new predecessor are dropped, and the first survivor gets its
// global start row.  Everything behind it keeps its local start rows (same predecessor as in the band's own
// envelope), so the cost per column is proportional to what is removed, not to what survives.  Per (band, column):
// lo/hi = surviving index range; per column: the list of non-empty bands (for the search in phase 3).
constexpr int kEdtMaxBands = 128;

__global__ void edt_merge_kernel(const uint16_t *__restrict__ sx, int w, int h, int n_bands, const uint16_t *__restrict__ S,
                                 uint16_t *__restrict__ T, const uint16_t *__restrict__ counts, uint16_t *__restrict__ lo_pl,
                                 uint16_t *__restrict__ hi_pl, uint16_t *__restrict__ ne_pl, int *__restrict__ n_ne) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= w) return;
  int n_live = 0;  // non-empty bands so far: ne_pl[i * w + x], their hi in hi_pl
  int top_band = -1, top_hi = 0, top_lo = 0;  // cached range of the last non-empty band
  for (int k = 0; k < n_bands; ++k) {
    const int cnt = counts[(size_t)k * w + x];
    const size_t r0 = (size_t)k * kEdtBandRows;
    int lo = 0;
    bool live = cnt > 0;
    if (cnt > 0 && top_band >= 0) {
      int j = 0;
      for (;;) {
        if (j == cnt) {  // the whole band lost against the envelope below it
          live = false;
          break;
        }
        const long long u = S[(r0 + j) * w + x];
        const long long dxu = (long long)x - sx[(size_t)u * w + x], gu = dxu * dxu;
        // pop tops that u already beats where they start
        long long ts = 0, tt = 0, tg = 0;
        while (top_band >= 0) {
          const size_t tr = (size_t)top_band * kEdtBandRows + top_hi - 1;
          ts = S[tr * w + x];
          tt = T[tr * w + x];
          const long long d = (long long)x - sx[(size_t)ts * w + x];
          tg = d * d;
          const long long a = tt - ts, bb = tt - u;
          if (a * a + tg > bb * bb + gu) {
            if (--top_hi <= top_lo) {  // that band is exhausted: fall back to the previous non-empty one
              hi_pl[(size_t)top_band * w + x] = (uint16_t)top_lo;
              --n_live;
              if (n_live > 0) {
                top_band = ne_pl[(size_t)(n_live - 1) * w + x];
                top_hi = hi_pl[(size_t)top_band * w + x];
                top_lo = lo_pl[(size_t)top_band * w + x];
              } else {
                top_band = -1;
              }
            }
          } else {
            break;
          }
        }
        if (top_band < 0) {  // nothing below survives: u starts the envelope
          T[(r0 + j) * w + x] = 0;
          lo = j;
          break;
        }
        const long long num = u * u - ts * ts + gu - tg;
        const long long wrow = 1 + floor_div(num, 2 * (u - ts));
        const long long next_t = j + 1 < cnt ? (long long)T[(r0 + j + 1) * w + x] : (long long)h;
        if (wrow >= next_t || wrow >= h) {  // u never wins between its new predecessor and its successor
          ++j;
          continue;
        }
        T[(r0 + j) * w + x] = (uint16_t)wrow;
        lo = j;
        break;
      }
    }
    if (top_band >= 0) hi_pl[(size_t)top_band * w + x] = (uint16_t)top_hi;  // publish pops
    lo_pl[(size_t)k * w + x] = (uint16_t)lo;
    hi_pl[(size_t)k * w + x] = (uint16_t)(live ? cnt : lo);
    if (live) {
      ne_pl[(size_t)n_live * w + x] = (uint16_t)k;
      ++n_live;
      top_band = k;
      top_hi = cnt;
      top_lo = lo;
    }
  }
  n_ne[x] = n_live;
}

__global__ void edt_fill_kernel(const uint16_t *__restrict__ sx, int w, int h, const uint16_t *__restrict__ S,
                                const uint16_t *__restrict__ T, const uint16_t *__restrict__ lo_pl, const uint16_t *__restrict__ hi_pl,
                                const uint16_t *__restrict__ ne_pl, const int *__restrict__ n_ne, uint32_t *__restrict__ nearest,
                                int32_t *__restrict__ dist2) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, band = blockIdx.y;
  if (x >= w) return;
  const int r0 = band * kEdtBandRows, r1 = min(h, r0 + kEdtBandRows);
  const int live = n_ne[x];
  if (live == 0) {
    for (int u = r0; u < r1; ++u) {
      nearest[(size_t)u * w + x] = 0xffffffffu;
      if (dist2) dist2[(size_t)u * w + x] = 0x7fffffff;
    }
    return;
  }
  // owner of row r1 - 1: last surviving entry whose start row is <= r1 - 1.  Start rows increase along the merged
  // sequence, so search the non-empty bands by the start row of their first survivor, then inside the band.
  int a = 0, b = live - 1;
  while (a < b) {
    const int mid = (a + b + 1) >> 1;
    const int kb = ne_pl[(size_t)mid * w + x];
    const int t0 = T[((size_t)kb * kEdtBandRows + lo_pl[(size_t)kb * w + x]) * w + x];
    if (t0 <= r1 - 1) a = mid;
    else b = mid - 1;
  }
  int li = a;                                   // index into the non-empty band list
  int kb = ne_pl[(size_t)li * w + x];
  int lo = lo_pl[(size_t)kb * w + x], hi = hi_pl[(size_t)kb * w + x];
  int ja = lo, jb = hi - 1;
  while (ja < jb) {
    const int mid = (ja + jb + 1) >> 1;
    if ((int)T[((size_t)kb * kEdtBandRows + mid) * w + x] <= r1 - 1) ja = mid;
    else jb = mid - 1;
  }
  int j = ja;
  long long top_s = S[((size_t)kb * kEdtBandRows + j) * w + x], top_t = T[((size_t)kb * kEdtBandRows + j) * w + x];
  long long site_x = sx[(size_t)top_s * w + x];
  for (int u = r1 - 1; u >= r0; --u) {
    nearest[(size_t)u * w + x] = (uint32_t)site_x | ((uint32_t)top_s << 16);
    if (dist2) {
      const long long dx = (long long)x - site_x, dy = (long long)u - top_s;
      dist2[(size_t)u * w + x] = (int32_t)(dx * dx + dy * dy);
    }
    if (u == top_t && (j > lo || li > 0)) {  // step to the previous survivor
      if (j > lo) {
        --j;
      } else {
        --li;
        kb = ne_pl[(size_t)li * w + x];
        lo = lo_pl[(size_t)kb * w + x];
        hi = hi_pl[(size_t)kb * w + x];
        j = hi - 1;
      }
      top_s = S[((size_t)kb * kEdtBandRows + j) * w + x];
      top_t = T[((size_t)kb * kEdtBandRows + j) * w + x];
      site_x = sx[(size_t)top_s * w + x];
    }
  }
}

// inflation as an EDT threshold: out bit = d^2 <= R^2
__global__ void edt_threshold_kernel(const int32_t *__restrict__ dist2, int w, int h, int pitch, int r2, uint32_t *__restrict__ out) {
  const size_t total = (size_t)pitch * h;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / (size_t)pitch), cw = (int)(i - (size_t)y * pitch);
    uint32_t v = 0;
    const int x0 = cw << 5;
    for (int b = 0; b < 32 && x0 + b < w; ++b)
      if (dist2[(size_t)y * w + x0 + b] <= r2) v |= 1u << b;
    out[i] = v;
  }
}

aos_status launch_edt(Ctx *c, const uint32_t *bits, int w, int h, uint32_t *nearest, int32_t *dist2) {
  AOS_REQUIRE(c, w > 0 && h > 0 && w < 65535 && h < 65535, "EDT grid must be smaller than 65535 x 65535");
  const int pitch = pitch_words_for(w);
  const size_t cells = (size_t)w * h;
  const int n_bands = (h + kEdtBandRows - 1) / kEdtBandRows;
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  AOS_REQUIRE(c, n_bands <= kEdtMaxBands, "EDT grid has more than 65536 rows");
  const size_t plane = up(cells * 2), cnt_bytes = up((size_t)n_bands * w * 2);
  AOS_CUDA_OK(c, c->edt_buf.reserve(3 * plane + 4 * cnt_bytes + up((size_t)w * 4) + 1024));
  char *base = c->edt_buf.as<char>();
  uint16_t *sx = reinterpret_cast<uint16_t *>(base);
  uint16_t *ss = reinterpret_cast<uint16_t *>(base + plane), *st = reinterpret_cast<uint16_t *>(base + 2 * plane);
  uint16_t *counts = reinterpret_cast<uint16_t *>(base + 3 * plane);
  uint16_t *lo_pl = reinterpret_cast<uint16_t *>(base + 3 * plane + cnt_bytes);
  uint16_t *hi_pl = reinterpret_cast<uint16_t *>(base + 3 * plane + 2 * cnt_bytes);
  uint16_t *ne_pl = reinterpret_cast<uint16_t *>(base + 3 * plane + 3 * cnt_bytes);
  int *totals = reinterpret_cast<int *>(base + 3 * plane + 4 * cnt_bytes);
  const int nw = (w + 31) >> 5;
  AOS_REQUIRE(c, nw <= 8 * kEdtRowThreads, "row too wide for the EDT row kernel");
  edt_rows_kernel<<<h, kEdtRowThreads, sizeof(int) * 2 * nw, c->stream>>>(bits, w, h, pitch, sx);
  ++c->launches;
  c->mark("edt_rows");
  dim3 gband((w + 127) / 128, n_bands);
  edt_band_sweep_kernel<<<gband, 128, 0, c->stream>>>(sx, w, h, ss, st, counts);
  ++c->launches;
  c->mark("edt_band_sweep");
  edt_merge_kernel<<<(w + 63) / 64, 64, 0, c->stream>>>(sx, w, h, n_bands, ss, st, counts, lo_pl, hi_pl, ne_pl, totals);
  ++c->launches;
  c->mark("edt_merge");
  edt_fill_kernel<<<gband, 128, 0, c->stream>>>(sx, w, h, ss, st, lo_pl, hi_pl, ne_pl, totals, nearest, dist2);
  ++c->launches;
  c->mark("edt_fill");
  AOS_CUDA_OK(c, cudaGetLastError());
  return AOS_OK;
}

aos_status launch_edt_threshold(Ctx *c, const int32_t *dist2, int w, int h, int r2, uint32_t *out) {
  const int pitch = pitch_words_for(w);
  size_t total = (size_t)pitch * h;
  int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)kNumSMs * 16);
  edt_threshold_kernel<<<grid < 1 ? 1 : grid, 256, 0, c->stream>>>(dist2, w, h, pitch, r2, out);
  ++c->launches;
  AOS_CUDA_OK(c, cudaGetLastError());
  return AOS_OK;
}

}  // namespace aos
