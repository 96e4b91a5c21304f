#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
typedef CUresult (*PFN)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int BW, int BH>
__global__ void k(const __grid_constant__ CUtensorMap tmap, int c0, int c1, uint32_t *out) {
  __shared__ __align__(128) uint32_t tile[BW * BH];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(BW * BH * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(tile)), "l"(&tmap), "r"(s32(&bar)), "r"(c0), "r"(c1) : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(s32(&bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = tile[i];
}
template <int BW, int BH>
int run(PFN enc, uint32_t *d, int pitch, int rows, int c0, int c1, const char *name) {
  CUtensorMap m;
  cuuint64_t gdim[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)pitch * 4};
  cuuint32_t box[2] = {BW, BH};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  uint32_t *out;
  cudaMalloc(&out, BW * BH * 4);
  k<BW, BH><<<1, 128>>>(m, c0, c1, out);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<uint32_t> h(BW * BH);
  cudaMemcpy(h.data(), out, BW * BH * 4, cudaMemcpyDeviceToHost);
  printf("%s: enc=%d sync=%s  first=%u,%u row1=%u last=%u\n", name, (int)r, cudaGetErrorString(e), h[0], h[1], h[BW], h[BW * BH - 1]);
  return e != cudaSuccess;
}
int main() {
  void *p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  PFN enc = (PFN)p;
  printf("entry %p q=%d\n", p, (int)q);
  int pitch = 128, rows = 256;
  std::vector<uint32_t> h(pitch * rows);
  for (int i = 0; i < pitch * rows; ++i) h[i] = i + 1;
  uint32_t *d;
  cudaMalloc(&d, pitch * rows * 4);
  cudaMemcpy(d, h.data(), pitch * rows * 4, cudaMemcpyHostToDevice);
  if (run<32, 8>(enc, d, pitch, rows, 0, 0, "32x8 @0,0")) return 1;
  if (run<36, 8>(enc, d, pitch, rows, -4, -3, "36x8 @-4,-3")) return 1;
  if (run<32, 96>(enc, d, pitch, rows, 31, 200, "32x96 @31,200")) return 1;
  if (run<36, 96>(enc, d, 16, rows, -4, -16, "36x96 pitch16 @-4,-16")) return 1;
  return 0;
}
