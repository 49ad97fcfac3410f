// Minimal check of the k_preprocess TMA box: (72, 34, 1) uint16 box of a (w, h, frames) tensor, negative start.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
struct alignas(64) TensorMap { unsigned long long opaque[16]; };
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#ifndef BOXW
#define BOXW 72
#endif
template <bool GLOBAL_DESC>
__global__ void k(const __grid_constant__ TensorMap tmap_p, const TensorMap* tmap_g, int x0, int y0, int f, uint16_t* out) {
  const void* tmap_ptr = GLOBAL_DESC ? (const void*)tmap_g : (const void*)&tmap_p;
  __shared__ __align__(128) uint16_t tile[34][BOXW];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)sizeof(tile)) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(&tile[0][0])),
                 "l"(tmap_ptr), "r"(x0), "r"(y0), "r"(f), "r"(smem_u32(&bar))
                 : "memory");
  }
  __syncthreads();
  asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@!p bra W;\n\t}" ::"r"(smem_u32(&bar)) : "memory");
  for (int i = threadIdx.x; i < 34 * BOXW; i += blockDim.x) out[i] = tile[i / BOXW][i % BOXW];
}
int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int w = 640, h = 480, pitch = 640, frames = 2;
  std::vector<uint16_t> hbuf((size_t)pitch * h * frames);
  for (size_t i = 0; i < hbuf.size(); ++i) hbuf[i] = (uint16_t)(1 + i % 60000);
  uint16_t *d, *o;
  cudaMalloc(&d, hbuf.size() * 2); cudaMalloc(&o, 34 * BOXW * 2);
  cudaMemcpy(d, hbuf.data(), hbuf.size() * 2, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  TensorMap tm;
  const cuuint64_t dims[3] = {w, h, frames}; const cuuint64_t strides[2] = {pitch * 2, (cuuint64_t)pitch * h * 2};
  const cuuint32_t box[3] = {BOXW, 34, 1}, es[3] = {1, 1, 1};
  CUresult r = ((EncodeFn)fn)((CUtensorMap*)&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d (q %d)\n", (int)r, (int)q);
  for (int t = 0; t < 3; ++t) {
    const int x0 = t == 0 ? -4 : t == 1 ? 60 : 636, y0 = t == 0 ? -1 : t == 1 ? 31 : 479, f = t == 2 ? 1 : 0;
    TensorMap* dtm; cudaMalloc(&dtm, sizeof(tm)); cudaMemcpy(dtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    if (variant == 0) k<false><<<1, 256>>>(tm, dtm, x0, y0, f, o); else k<true><<<1, 256>>>(tm, dtm, x0, y0, f, o);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<uint16_t> ho(34 * BOXW);
    cudaMemcpy(ho.data(), o, ho.size() * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < 34; ++r2) for (int c = 0; c < BOXW; ++c) {
      const int x = x0 + c, y = y0 + r2;
      const uint16_t want = (x >= 0 && x < w && y >= 0 && y < h) ? hbuf[((size_t)f * h + y) * pitch + x] : 0;
      bad += ho[r2 * BOXW + c] != want;
    }
    printf("case %d: %s, mismatches %d\n", t, cudaGetErrorString(e), bad);
  }
  return 0;
}
