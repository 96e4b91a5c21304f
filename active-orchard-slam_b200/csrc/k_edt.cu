// k_edt.cu -- exact Euclidean distance transform with nearest-site labels on the bit-packed grid.
//
// Where an exact EDT legitimately enters this path: applyInflation
// (src/aos_seed_gen_node.cpp:933-967) is the threshold d^2 <= R^2 of the exact squared EDT of the raw occupancy
// grid (property-tested against the stencil kernel), and the declared-but-unfilled GvdGraph.edge_clearances
// (msg/GvdGraph.msg:58, gvd:856,890 publish 0.0f) is a distance-to-skeleton query -- offered as an opt-in
// (aos_set_clearance), default off so the published arrays stay bit-identical to the reference.
// aos_gvd_node's Voronoi diagram itself is cv::Subdiv2D's point-site diagram (host_subdiv.cu), not a raster.
//
// Exact, integer-only, separable (the structure of the parallel banding algorithm with one band per line):
//   phase 1  per row   : nearest set cell of the same row for every cell (bit tricks inside a word, block-wide
//                        max/min scans across the words of the row)                      -> uint16 x of that cell
//   phase 2  per column: lower envelope of the parabolas (x - sx(y))^2 + (u - y)^2 by the stack sweep with exact
//                        integer separators (floor division), one thread per column, coalesced across columns
//   phase 3  per column: backward sweep over the stack writes the nearest site (x | y << 16) and d^2
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
  uint16_t *out = sx + (size_t)y * w;
  for (int k = threadIdx.x; k < nw; k += kEdtRowThreads) {
    uint32_t v = row[k];
    if (k == nw - 1 && (w & 31)) v &= (1u << (w & 31)) - 1u;
    const int left_before = k > 0 ? hi[k - 1] : -1;                // nearest set cell in earlier words
    const int right_after = k + 1 < nw ? lo[k + 1] : 0x7fffffff;   // nearest set cell in later words
    const int x0 = k << 5;
    const int nb = min(32, w - x0);
    for (int b = 0; b < nb; ++b) {
      const uint32_t le = v & (0xffffffffu >> (31 - b));  // bits <= b
      const uint32_t ge = v & (0xffffffffu << b);         // bits >= b
      const int L = le ? x0 + 31 - __clz(le) : left_before;
      const int R = ge ? x0 + __ffs(ge) - 1 : right_after;
      const int x = x0 + b;
      int s;
      if (L < 0 && R == 0x7fffffff) s = kNoSite16;
      else if (L < 0) s = R;
      else if (R == 0x7fffffff) s = L;
      else s = (x - L <= R - x) ? L : R;
      out[x] = (uint16_t)s;
    }
  }
}

// ---- phases 2 + 3: one thread per column -------------------------------------------------------------------
__device__ __forceinline__ long long floor_div(long long a, long long b) {  // b > 0
  long long q = a / b;
  return (a % b != 0 && a < 0) ? q - 1 : q;
}

__global__ void edt_columns_kernel(const uint16_t *__restrict__ sx, int w, int h, uint16_t *__restrict__ stack_s,
                                   uint16_t *__restrict__ stack_t, uint32_t *__restrict__ nearest, int32_t *__restrict__ dist2) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= w) return;
  // forward sweep: s[q] = row of the q-th envelope site, t[q] = first row where it wins
  int q = -1;
  long long top_s = 0, top_t = 0, top_g = 0;  // cached top of the stack
  for (int u = 0; u < h; ++u) {
    const uint16_t sxu = sx[(size_t)u * w + x];
    if (sxu == kNoSite16) continue;
    const long long dx = (long long)x - sxu, gu = dx * dx;
    while (q >= 0) {
      const long long a = top_t - top_s, b = top_t - u;
      if (a * a + top_g > b * b + gu) {  // the new parabola already wins where the top one starts: pop
        --q;
        if (q >= 0) {
          top_s = stack_s[(size_t)q * w + x];
          top_t = stack_t[(size_t)q * w + x];
          const long long d = (long long)x - sx[(size_t)top_s * w + x];
          top_g = d * d;
        }
      } else {
        break;
      }
    }
    if (q < 0) {
      q = 0;
      top_s = u;
      top_t = 0;
      top_g = gu;
      stack_s[x] = (uint16_t)u;
      stack_t[x] = 0;
    } else {
      // first row where u beats top_s: 1 + floor((u^2 - s^2 + g(u) - g(s)) / (2 (u - s)))
      const long long num = (long long)u * u - top_s * top_s + gu - top_g;
      const long long wrow = 1 + floor_div(num, 2 * ((long long)u - top_s));
      if (wrow < h) {
        ++q;
        top_s = u;
        top_t = wrow;
        top_g = gu;
        stack_s[(size_t)q * w + x] = (uint16_t)u;
        stack_t[(size_t)q * w + x] = (uint16_t)wrow;
      }
    }
  }
  // backward sweep
  if (q < 0) {
    for (int u = 0; u < h; ++u) {
      nearest[(size_t)u * w + x] = 0xffffffffu;
      if (dist2) dist2[(size_t)u * w + x] = 0x7fffffff;
    }
    return;
  }
  long long site_x = sx[(size_t)top_s * w + x];
  for (int u = h - 1; u >= 0; --u) {
    nearest[(size_t)u * w + x] = (uint32_t)site_x | ((uint32_t)top_s << 16);
    if (dist2) {
      const long long dx = (long long)x - site_x, dy = (long long)u - top_s;
      dist2[(size_t)u * w + x] = (int32_t)(dx * dx + dy * dy);
    }
    if (u == top_t && q > 0) {
      --q;
      top_s = stack_s[(size_t)q * w + x];
      top_t = stack_t[(size_t)q * w + x];
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
  AOS_CUDA_OK(c, c->edt_buf.reserve(cells * 2 * 3 + 1024));
  uint16_t *sx = c->edt_buf.as<uint16_t>();
  uint16_t *ss = sx + cells, *st = ss + cells;
  const int nw = (w + 31) >> 5;
  AOS_REQUIRE(c, nw <= 8 * kEdtRowThreads, "row too wide for the EDT row kernel");
  edt_rows_kernel<<<h, kEdtRowThreads, sizeof(int) * 2 * nw, c->stream>>>(bits, w, h, pitch, sx);
  ++c->launches;
  edt_columns_kernel<<<(w + 63) / 64, 64, 0, c->stream>>>(sx, w, h, ss, st, nearest, dist2);
  ++c->launches;
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
