// k_thin.cu -- Zhang-Suen thinning (cv::ximgproc::thinning(THINNING_ZHANGSUEN), call site
// src/aos_seed_gen_node.cpp:684) on the bit-packed grid.
//
// One launch runs kSub sub-iterations (0,1,0,1,...) on a tile held in shared memory (temporal
// blocking): the tile is loaded by TMA with a halo of kSub rows above/below and one 32-cell word left/
// right, every sub-iteration invalidates one more ring of the halo, and after kSub of them the owned
// interior is exact.  A sub-iteration is the parallel update of the published algorithm -- deletions are
// computed from the pre-sub-iteration image (ping-pong between two shared-memory tiles) -- evaluated
// 32 pixels at a time with bit-sliced logic: the neighbour count B through a carry-save adder tree, the
// transition count A==1 through a one/two accumulator, 32 lanes of a warp = 32 consecutive words of a row,
// left/right words by shuffle.  Image-border pixels (row 0, H-1, column 0, W-1) are never deleted, as in
// opencv_contrib's loops (1..rows-2, 1..cols-2).
//
// Convergence: the reference stops when a full (0,1) pair changes nothing; extra sub-iterations at the
// fixed point are no-ops, so running whole launches until the last pair of one deletes nothing in the owned
// regions gives the same image.  Launches are queued in batches; each launch first reads its predecessor's deletion
// counter and exits at once if that was zero, so the host synchronises once per batch, not per launch.
#include "aos_common.cuh"

namespace aos {

constexpr int kSub = kThinSubIters;           // sub-iterations per launch (even)
constexpr int kThinOwn = 112;                 // owned rows per CTA
constexpr int kThinBox = kThinOwn + 2 * kSub; // 128 tile rows
constexpr int kThinThreads = 256;
constexpr int kThinBatch = 8;                 // launches queued per host synchronisation

__device__ __forceinline__ uint32_t zs_delete_mask(uint32_t C, uint32_t p2, uint32_t p3, uint32_t p4, uint32_t p5,
                                                   uint32_t p6, uint32_t p7, uint32_t p8, uint32_t p9, int iter) {
  // B = p2+...+p9 by carry-save addition
  uint32_t s1 = p2 ^ p3 ^ p4, c1 = (p2 & p3) | (p4 & (p2 ^ p3));
  uint32_t s2 = p5 ^ p6 ^ p7, c2 = (p5 & p6) | (p7 & (p5 ^ p6));
  uint32_t s3 = p8 ^ p9, c3 = p8 & p9;
  uint32_t b0 = s1 ^ s2 ^ s3, c4 = (s1 & s2) | (s3 & (s1 ^ s2));
  uint32_t s5 = c1 ^ c2 ^ c3, c5 = (c1 & c2) | (c3 & (c1 ^ c2));
  uint32_t b1 = s5 ^ c4, c6 = s5 & c4;
  uint32_t b2 = c5 ^ c6, b3 = c5 & c6;
  uint32_t condB = (b1 | b2) & ~b3 & ~(b0 & b1 & b2);  // 2 <= B <= 6
  // A = number of 0->1 transitions in p2,p3,...,p9,p2 ; need exactly one
  uint32_t one = ~p2 & p3, two = 0, t;
  t = ~p3 & p4; two |= one & t; one |= t;
  t = ~p4 & p5; two |= one & t; one |= t;
  t = ~p5 & p6; two |= one & t; one |= t;
  t = ~p6 & p7; two |= one & t; one |= t;
  t = ~p7 & p8; two |= one & t; one |= t;
  t = ~p8 & p9; two |= one & t; one |= t;
  t = ~p9 & p2; two |= one & t; one |= t;
  uint32_t condA = one & ~two;
  uint32_t condM = iter == 0 ? ~(p4 & p6 & (p2 | p8)) : ~(p2 & p8 & (p4 | p6));
  return C & condB & condA & condM;
}

struct ThinParams {
  int w, h, pitch;
  int y_off, gh;       // row-band mode: global row of local row 0, global height
  int cnt_r0, cnt_r1;  // local rows whose deletions count towards convergence (the band without its halo)
  // fused halo exchange (aos_band_thin_launch_p2p): the band's first / last kThinSubIters rows are ALSO stored into
  // the neighbouring GPU's destination buffer (peer memory over NVLink: its halo rows next to its band), and the
  // local halo rows the neighbours provide are not written here.  Null pointers / empty ranges switch it off.
  uint32_t *peer_lo, *peer_hi;       // neighbour's destination grid (same pitch), or null
  int push_lo_r0, push_lo_shift;     // local rows [push_lo_r0, +kSub) go to peer_lo row (local + push_lo_shift)
  int push_hi_r0, push_hi_shift;
  int skip_lo_r0, skip_hi_r0;        // local rows [skip_*_r0, +kSub) are provided by the neighbour (-1: none)
};

__global__ void __launch_bounds__(kThinThreads) thin_kernel(const __grid_constant__ CUtensorMap tmap,
                                                            const __grid_constant__ ThinParams P,
                                                            uint32_t *__restrict__ dst, const int *prev_count,
                                                            int *my_count) {
  __shared__ __align__(128) uint32_t buf[2][kThinBox * kTileBoxW];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int s_deleted;
  if (prev_count && *prev_count == 0) return;  // predecessor already at the fixed point

  const int own0 = blockIdx.x * kTileOwnW;
  const int y0 = blockIdx.y * kThinOwn - kSub;  // image row of tile row 0
  if (threadIdx.x == 0) {
    s_deleted = 0;
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, (uint32_t)(kThinBox * kTileBoxW * 4));
    tma_load_2d(buf[0], &tmap, &bar, own0 - 4, y0);
  }
  mbar_wait(&bar, 0);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cw = own0 - kTileLane0 + lane;
  const int sl = lane + kTileLane0;  // this lane's word inside a staged row
  // pixels that may never be deleted: x == 0, x == w-1 (and rows 0, h-1 below)
  uint32_t xmask = 0xffffffffu;
  if (cw == 0) xmask &= ~1u;
  if (cw == ((P.w - 1) >> 5)) xmask &= ~(1u << ((P.w - 1) & 31));
  const bool lane_owned = lane >= kTileLane0 && lane < kTileLane0 + kTileOwnW && cw >= 0;
  constexpr int kRowsPerWarp = kThinBox / (kThinThreads / 32);  // 16
  const int r_begin = warp * kRowsPerWarp;
  // Per warp, once: which of its 16 tile rows may lose pixels at all (the tile's first and last row lack a neighbour row,
  // the image's first and last row are never thinned) and which count towards convergence.  The row loop below is fully
  // unrolled, so these become one bit test per row instead of six compares.
  uint32_t del_rows = 0, cnt_rows = 0;
#pragma unroll
  for (int j = 0; j < kRowsPerWarp; ++j) {
    const int r = r_begin + j, y = y0 + r;
    if (r > 0 && r < kThinBox - 1 && y + P.y_off > 0 && y + P.y_off < P.gh - 1) del_rows |= 1u << j;
    if (r >= kSub && r < kSub + kThinOwn && y >= P.cnt_r0 && y < P.cnt_r1) cnt_rows |= 1u << j;
  }
  uint32_t dacc = 0;  // pixels deleted in counted rows during the last (0,1) pair of this launch

  int cur = 0;
#pragma unroll 1
  for (int s = 0; s < kSub; ++s) {
    const uint32_t *src = buf[cur] + sl;
    uint32_t *out = buf[cur ^ 1] + sl;
    const int iter = s & 1;
    // convergence = the LAST full (0,1) pair of this launch deleted nothing in the owned rows: that is the reference's
    // stopping rule, and it spares the extra launch that would only confirm the fixed point
    if (s == kSub - 2) dacc = 0;
    // sliding window over rows: (centre, west-neighbour plane, east-neighbour plane) of rows r-1, r, r+1
    auto planes = [&](uint32_t c, uint32_t &wv, uint32_t &ev) {
      uint32_t l = __shfl_up_sync(0xffffffffu, c, 1), rr = __shfl_down_sync(0xffffffffu, c, 1);
      wv = __funnelshift_l(l, c, 1);   // bit i = pixel x-1
      ev = __funnelshift_r(c, rr, 1);  // bit i = pixel x+1
    };
    uint32_t nC, nW, nE, cC, cW, cE, sC, sW, sE;
    nC = r_begin > 0 ? src[(r_begin - 1) * kTileBoxW] : 0u;
    planes(nC, nW, nE);
    cC = src[r_begin * kTileBoxW];
    planes(cC, cW, cE);
#pragma unroll
    for (int j = 0; j < kRowsPerWarp; ++j) {
      const int r = r_begin + j;
      if (j + 1 < kRowsPerWarp) sC = src[(r + 1) * kTileBoxW];
      else sC = r + 1 < kThinBox ? src[(r + 1) * kTileBoxW] : 0u;
      planes(sC, sW, sE);
      uint32_t res = cC;
      if ((del_rows >> j) & 1u) {
        // a pixel with all 8 neighbours set (or a zero pixel) cannot go: skip solid / empty stretches
        uint32_t boundary = cC & ~(nC & nW & nE & cW & cE & sC & sW & sE);
        if (__any_sync(0xffffffffu, boundary != 0)) {
          uint32_t del = zs_delete_mask(cC, nC, nE, cE, sE, sC, sW, cW, nW, iter) & xmask;
          res = cC & ~del;
          if ((cnt_rows >> j) & 1u) dacc |= del;
        }
      }
      out[r * kTileBoxW] = res;
      nC = cC; nW = cW; nE = cE;
      cC = sC; cW = sW; cE = sE;
    }
    cur ^= 1;
    __syncthreads();
  }
  const int deleted_owned = lane_owned && dacc != 0;
  // write back the owned interior (rows kSub .. kSub+kThinOwn-1, lanes 1..30)
  const uint32_t *fin = buf[cur];
  for (int r = kSub + warp; r < kSub + kThinOwn; r += kThinThreads / 32) {
    int y = y0 + r;
    if (y >= P.h) break;
    if (lane_owned && cw < P.pitch) {
      const uint32_t v = fin[r * kTileBoxW + sl];
      const bool from_lo = P.skip_lo_r0 >= 0 && y >= P.skip_lo_r0 && y < P.skip_lo_r0 + kSub;
      const bool from_hi = P.skip_hi_r0 >= 0 && y >= P.skip_hi_r0 && y < P.skip_hi_r0 + kSub;
      if (!from_lo && !from_hi) dst[(size_t)y * P.pitch + cw] = v;
      if (P.peer_lo && y >= P.push_lo_r0 && y < P.push_lo_r0 + kSub) P.peer_lo[(size_t)(y + P.push_lo_shift) * P.pitch + cw] = v;
      if (P.peer_hi && y >= P.push_hi_r0 && y < P.push_hi_r0 + kSub) P.peer_hi[(size_t)(y + P.push_hi_shift) * P.pitch + cw] = v;
    }
  }
  if (__any_sync(0xffffffffu, deleted_owned) && lane == 0) atomicOr(&s_deleted, 1);
  __syncthreads();
  if (threadIdx.x == 0 && s_deleted) atomicAdd(my_count, 1);
}

// One launch (kSub sub-iterations) src -> dst on a LOCAL grid of a row band (aos_band_thin_launch): the caller
// refreshes the kSub halo rows next to the band from the neighbouring GPUs between launches.  d_count receives the
// number of CTAs that deleted something inside rows [cnt_r0, cnt_r1).
aos_status launch_thin_once(Ctx *c, const uint32_t *src, uint32_t *dst, int w, int h, int y_off, int gh, int cnt_r0,
                            int cnt_r1, int *d_count, const ThinHalo *halo) {
  ThinParams P{w, h, pitch_words_for(w), y_off, gh, cnt_r0, cnt_r1, nullptr, nullptr, 0, 0, 0, 0, -1, -1};
  if (halo) {
    P.peer_lo = halo->peer_lo;
    P.peer_hi = halo->peer_hi;
    P.push_lo_r0 = halo->push_lo_r0;
    P.push_lo_shift = halo->push_lo_shift;
    P.push_hi_r0 = halo->push_hi_r0;
    P.push_hi_shift = halo->push_hi_shift;
    P.skip_lo_r0 = halo->skip_lo_r0;
    P.skip_hi_r0 = halo->skip_hi_r0;
  }
  CUtensorMap map_src;
  if (!make_bitgrid_tmap(&map_src, src, P.pitch, h, kTileBoxW, kThinBox)) {
    set_error(c, "cuTensorMapEncodeTiled failed (thin)");
    return AOS_ERR_CUDA;
  }
  int words_used = (w + 31) >> 5;
  dim3 grid((words_used + kTileOwnW - 1) / kTileOwnW, (h + kThinOwn - 1) / kThinOwn);
  AOS_CUDA_OK(c, cudaMemsetAsync(d_count, 0, sizeof(int), c->stream));
  thin_kernel<<<grid, kThinThreads, 0, c->stream>>>(map_src, P, dst, nullptr, d_count);
  ++c->launches;
  AOS_CUDA_OK(c, cudaGetLastError());
  return AOS_OK;
}

// img holds the input; on return *result_in_scratch says which of (img, scratch) holds the skeleton.
aos_status launch_thin(Ctx *c, uint32_t *img, uint32_t *scratch, int w, int h, int *launches, int *subiters) {
  ThinParams P{w, h, pitch_words_for(w), 0, h, 0, h, nullptr, nullptr, 0, 0, 0, 0, -1, -1};
  CUtensorMap map_img, map_scr;
  if (!make_bitgrid_tmap(&map_img, img, P.pitch, h, kTileBoxW, kThinBox) ||
      !make_bitgrid_tmap(&map_scr, scratch, P.pitch, h, kTileBoxW, kThinBox)) {
    set_error(c, "cuTensorMapEncodeTiled failed (thin)");
    return AOS_ERR_CUDA;
  }
  int words_used = (w + 31) >> 5;
  dim3 grid((words_used + kTileOwnW - 1) / kTileOwnW, (h + kThinOwn - 1) / kThinOwn);
  AOS_CUDA_OK(c, c->misc.reserve(4096));
  int *d_counts = c->misc.as<int>() + 64;  // [kThinBatch] deletion counters
  int total = 0;
  bool src_is_img = true;
  for (int batch = 0; batch < 4096; ++batch) {
    AOS_CUDA_OK(c, cudaMemsetAsync(d_counts, 0, sizeof(int) * kThinBatch, c->stream));
    for (int j = 0; j < kThinBatch; ++j) {
      bool from_img = src_is_img ^ (j & 1);
      thin_kernel<<<grid, kThinThreads, 0, c->stream>>>(from_img ? map_img : map_scr, P, from_img ? scratch : img,
                                                        j == 0 ? nullptr : d_counts + j - 1, d_counts + j);
  ++c->launches;
    }
    AOS_CUDA_OK(c, cudaGetLastError());
    AOS_CUDA_OK(c, cudaMemcpyAsync(c->h_flag, d_counts, sizeof(int) * kThinBatch, cudaMemcpyDeviceToHost, c->stream));
    AOS_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    int conv = -1;
    for (int j = 0; j < kThinBatch; ++j)
      if (c->h_flag[j] == 0) {
        conv = j;
        break;
      }
    if (conv >= 0) {
      total += conv + 1;
      // launch `conv` wrote an unchanged copy: its destination holds the skeleton
      bool from_img = src_is_img ^ (conv & 1);
      bool result_in_img = !from_img;
      if (!result_in_img)
        AOS_CUDA_OK(c, cudaMemcpyAsync(img, scratch, (size_t)P.pitch * h * 4, cudaMemcpyDeviceToDevice, c->stream));
      break;
    }
    total += kThinBatch;
    // kThinBatch is even: the next batch reads from the same buffer kind as this one did
  }
  if (launches) *launches = total;
  if (subiters) *subiters = total * kSub;
  return AOS_OK;
}

}  // namespace aos
