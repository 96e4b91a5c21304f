// k_inflate.cu -- disc inflation (applyInflation, src/aos_seed_gen_node.cpp:933-967) with the 5-cell
// frame of markBoundariesAsOccupied (:708-757) as a second output, and the 3x3-cross morphological
// opening of skeletonizeOccupancyGrid (:678-680).  Both are bit-parallel stencils on the packed grid:
// a warp owns 32 consecutive words (1024 cells) of a row, tiles are staged in shared memory by TMA
// (out-of-image parts of the box arrive zero-filled, which is exactly the "clipped to the grid"
// rule of the reference), neighbours' words come from warp shuffles.
//
// Inflation by a disc of radius R is  out(y) = OR_d H_{hw[|d|]}(in(y+d)),  hw[d] = floor(sqrt(R^2-d^2)),
// H_k = horizontal dilation by +-k.  Since H_{a+b} = H_a o H_b and hw is non-increasing in d, the rows
// are folded into one chain  acc = U_0;  acc = H_{hw[d-1]-hw[d]}(acc) | U_d  (d = 1..R),  U_d = in(y+d)|in(y-d),
// so the total horizontal work per output row is one dilation by R, not 2R+1 of them.
#include "aos_common.cuh"

namespace aos {

constexpr int kInfRows = 64;   // output rows per CTA
constexpr int kInfThreads = 256;
constexpr int kMaxR = kMaxStencilRadius;

struct InflateParams {
  int w, h, pitch, R;
  int y_off, gh;  // global row of local row 0, global height (frame of markBoundariesAsOccupied)
  unsigned char delta[kMaxR + 1];  // delta[d] = hw[d-1]-hw[d] for d>=1
};

// frame of thickness t (markBoundariesAsOccupied): bits of word column cw at row y that belong to it
__device__ __forceinline__ uint32_t frame_mask(int cw, int y, int w, int h, int t) {
  if (y < t || y >= h - t) return 0xffffffffu;
  uint32_t m = 0;
  int x0 = cw << 5;
  // left band [0, t)
  if (x0 < t) m |= (t - x0 >= 32) ? 0xffffffffu : ((1u << (t - x0)) - 1u);
  // right band [w-t, w)
  int lo = w - t - x0;  // first bit index of the band inside this word
  if (lo < 32) m |= (lo <= 0) ? 0xffffffffu : ~((1u << lo) - 1u);
  return m;
}
__device__ __forceinline__ uint32_t valid_mask(int cw, int w) {
  int x0 = cw << 5;
  if (cw < 0 || x0 >= w) return 0;
  int n = w - x0;
  return n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
}

// OR of all shifts in [-t, t] needs the neighbours' current words: one doubling step widens the
// covered interval from [-c, c] to [-(c+t), c+t] (t <= 2c+1).
__device__ __forceinline__ uint32_t widen(uint32_t v, int t) {
  uint32_t l = __shfl_up_sync(0xffffffffu, v, 1);    // word to the left  (lower x)
  uint32_t r = __shfl_down_sync(0xffffffffu, v, 1);  // word to the right (higher x)
  // cell x moves to x+t : (v << t) | (l >> (32-t));  cell x moves to x-t : (v >> t) | (r << (32-t))
  uint32_t up = __funnelshift_l(l, v, t);
  uint32_t dn = __funnelshift_r(v, r, t);
  return v | up | dn;
}

__global__ void __launch_bounds__(kInfThreads) inflate_kernel(const __grid_constant__ CUtensorMap tmap,
                                                              const __grid_constant__ InflateParams P,
                                                              uint32_t *__restrict__ out,
                                                              uint32_t *__restrict__ out_border) {
  extern __shared__ __align__(128) uint32_t tile[];  // (kInfRows + 2R) x kTileBoxW
  __shared__ __align__(8) uint64_t bar;
  const int R = P.R;
  const int own0 = blockIdx.x * kTileOwnW;  // first owned word column
  const int y0 = blockIdx.y * kInfRows;
  const int box_h = kInfRows + 2 * R;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, (uint32_t)(box_h * kTileBoxW * 4));
    tma_load_2d(tile, &tmap, &bar, own0 - 4, y0 - R);
  }
  mbar_wait(&bar, 0);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cw = own0 - kTileLane0 + lane;
  const bool lane_ok = lane >= kTileLane0 && lane < kTileLane0 + kTileOwnW && cw < P.pitch;
  const uint32_t vmask = valid_mask(cw, P.w);
  for (int ry = warp; ry < kInfRows; ry += kInfThreads / 32) {
    const int y = y0 + ry;
    if (y >= P.h) break;
    const uint32_t *ctr = tile + (ry + R) * kTileBoxW + lane + kTileLane0;
    uint32_t acc = *ctr;
    for (int d = 1; d <= R; ++d) {
      int delta = P.delta[d];
      if (delta && __any_sync(0xffffffffu, acc != 0)) {
        int c = 0;
        while (c < delta) {
          int t = min(2 * c + 1, delta - c);
          acc = widen(acc, t);
          c += t;
        }
      }
      acc |= ctr[d * kTileBoxW] | ctr[-d * kTileBoxW];
    }
    if (lane_ok) {
      acc &= vmask;
      size_t o = (size_t)y * P.pitch + cw;
      out[o] = acc;
      if (out_border) out_border[o] = (acc | frame_mask(cw, y + P.y_off, P.w, P.gh, 5)) & vmask;
    }
  }
}

aos_status launch_inflate(Ctx *c, const uint32_t *in, uint32_t *out, uint32_t *out_border, int w, int h, int R) {
  AOS_REQUIRE(c, R >= 0 && R <= kMaxR, "inflation radius above 64 cells is not supported");
  InflateParams P{};
  P.w = w;
  P.h = h;
  P.pitch = pitch_words_for(w);
  P.R = R;
  P.y_off = c->band_gh ? c->band_y_off : 0;
  P.gh = c->band_gh ? c->band_gh : h;
  int prev = R;  // hw[0] = R
  for (int d = 1; d <= R; ++d) {
    int k = 0;
    while ((k + 1) * (k + 1) + d * d <= R * R) ++k;  // hw[d]
    P.delta[d] = (unsigned char)(prev - k);
    prev = k;
  }
  CUtensorMap tmap;
  int box_h = kInfRows + 2 * R;
  if (!make_bitgrid_tmap(&tmap, in, P.pitch, h, kTileBoxW, box_h)) {
    set_error(c, "cuTensorMapEncodeTiled failed (inflate)");
    return AOS_ERR_CUDA;
  }
  int words_used = (w + 31) >> 5;
  dim3 grid((words_used + kTileOwnW - 1) / kTileOwnW, (h + kInfRows - 1) / kInfRows);
  size_t smem = (size_t)box_h * kTileBoxW * 4;
  // always the maximum (R = kMaxR): contexts on other host threads launch this kernel with other radii, and the
  // attribute is per function, not per launch
  AOS_CUDA_OK(c, cudaFuncSetAttribute(inflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)((size_t)(kInfRows + 2 * kMaxR) * kTileBoxW * 4)));
  inflate_kernel<<<grid, kInfThreads, smem, c->stream>>>(tmap, P, out, out_border);
  ++c->launches;
  AOS_CUDA_OK(c, cudaGetLastError());
  return AOS_OK;
}

// ---- 3x3 cross opening (erode, then dilate) with OpenCV's default morphology border: pixels
//      outside the image never constrain the erosion and never add to the dilation. -----------------
constexpr int kOpenRows = 64;
constexpr int kOpenThreads = 256;

__global__ void __launch_bounds__(kOpenThreads) open_kernel(const __grid_constant__ CUtensorMap tmap, int w, int h,
                                                            int pitch, int y_off, int gh, uint32_t *__restrict__ out) {
  __shared__ __align__(128) uint32_t tile[(kOpenRows + 4) * kTileBoxW];
  __shared__ __align__(8) uint64_t bar;
  const int own0 = blockIdx.x * kTileOwnW;
  const int y0 = blockIdx.y * kOpenRows;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, (uint32_t)((kOpenRows + 4) * kTileBoxW * 4));
    tma_load_2d(tile, &tmap, &bar, own0 - 4, y0 - 2);
  }
  mbar_wait(&bar, 0);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cw = own0 - kTileLane0 + lane;
  const uint32_t vmask = valid_mask(cw, w);
  const uint32_t *tl = tile + lane + kTileLane0;
  // bit that holds x = 0 / x = w-1 in this word column (0 if not here)
  const uint32_t first_bit = (cw == 0) ? 1u : 0u;
  const uint32_t last_bit = (cw == ((w - 1) >> 5)) ? (1u << ((w - 1) & 31)) : 0u;
  const int rows_per_warp = kOpenRows / (kOpenThreads / 32);  // 8
  const int ry0 = warp * rows_per_warp;

  // eroded row at image row y (tile row ty = y - y0 + 2)
  auto eroded = [&](int y) -> uint32_t {
    // eroded image does not exist outside the (global) image: adds nothing to the dilation
    if (y < 0 || y >= h || y + y_off < 0 || y + y_off >= gh) return 0u;
    int ty = y - y0 + 2;
    uint32_t C = tl[ty * kTileBoxW];
    uint32_t N = (y + y_off - 1 < 0) ? 0xffffffffu : tl[(ty - 1) * kTileBoxW];
    uint32_t S = (y + y_off + 1 >= gh) ? 0xffffffffu : tl[(ty + 1) * kTileBoxW];
    uint32_t l = __shfl_up_sync(0xffffffffu, C, 1), r = __shfl_down_sync(0xffffffffu, C, 1);
    uint32_t Wn = __funnelshift_l(l, C, 1) | first_bit;  // value of the x-1 neighbour at x
    uint32_t En = __funnelshift_r(C, r, 1) | last_bit;   // value of the x+1 neighbour at x
    return C & N & S & Wn & En & vmask;
  };
  uint32_t e_prev = eroded(y0 + ry0 - 1), e_cur = eroded(y0 + ry0);
  for (int k = 0; k < rows_per_warp; ++k) {
    int y = y0 + ry0 + k;
    uint32_t e_next = eroded(y + 1);  // all lanes execute the shuffles inside
    uint32_t l = __shfl_up_sync(0xffffffffu, e_cur, 1), r = __shfl_down_sync(0xffffffffu, e_cur, 1);
    uint32_t d = e_cur | e_prev | e_next | __funnelshift_l(l, e_cur, 1) | __funnelshift_r(e_cur, r, 1);
    if (y < h && lane >= kTileLane0 && lane < kTileLane0 + kTileOwnW && cw < pitch) out[(size_t)y * pitch + cw] = d & vmask;
    e_prev = e_cur;
    e_cur = e_next;
  }
}

aos_status launch_open(Ctx *c, const uint32_t *in, uint32_t *out, int w, int h) {
  int pitch = pitch_words_for(w);
  CUtensorMap tmap;
  if (!make_bitgrid_tmap(&tmap, in, pitch, h, kTileBoxW, kOpenRows + 4)) {
    set_error(c, "cuTensorMapEncodeTiled failed (open)");
    return AOS_ERR_CUDA;
  }
  int words_used = (w + 31) >> 5;
  dim3 grid((words_used + kTileOwnW - 1) / kTileOwnW, (h + kOpenRows - 1) / kOpenRows);
  open_kernel<<<grid, kOpenThreads, 0, c->stream>>>(tmap, w, h, pitch, c->band_gh ? c->band_y_off : 0, c->band_gh ? c->band_gh : h,
                                                     out);
  ++c->launches;
  AOS_CUDA_OK(c, cudaGetLastError());
  return AOS_OK;
}

}  // namespace aos
