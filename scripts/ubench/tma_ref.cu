// CUDA programming guide style TMA load (libcu++ wrappers), variants selected by argv[1]:
// 0: 2D int32 32x32 box in bounds; 1: 2D uint16 72x34 box, start (-4,-1); 2: 3D uint16 72x34x1
#include <cuda.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
template <typename T, int BW, int BH, int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int z, T* out) {
  __shared__ alignas(128) T smem[BH][BW];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    if (RANK == 2) cde::cp_async_bulk_tensor_2d_global_to_shared(&smem, &tensor_map, x, y, bar);
    else cde::cp_async_bulk_tensor_3d_global_to_shared(&smem, &tensor_map, x, y, z, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem));
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = smem[i / BW][i % BW];
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  const int v = argc > 1 ? atoi(argv[1]) : 0;
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  const int w = 640, h = 480, frames = 2;
  void* d; cudaMalloc(&d, (size_t)w * h * frames * 4); cudaMemset(d, 1, (size_t)w * h * frames * 4);
  void* o; cudaMalloc(&o, 80 * 34 * 4);
  CUtensorMap tm;
  CUresult r;
  // v bits: 1 = negative start, 2 = box 72 (else 64), 4 = rows 34 (else 32), 8 = start x 636 (crosses the right edge)
  const int x0 = (v & 1) ? -8 : ((v & 8) ? 632 : 56), y0 = (v & 1) ? -1 : 31;
  {
    const cuuint64_t dims[2] = {w, h}; const cuuint64_t st[1] = {w * 2};
    const cuuint32_t box[2] = {(v & 2) ? 80u : 64u, (v & 4) ? 34u : 32u}, es[2] = {1, 1};
    r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                       CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d box %u x %u start %d %d\n", (int)r, box[0], box[1], x0, y0);
    if ((v & 6) == 0) k<uint16_t, 64, 32, 2><<<1, 128>>>(tm, x0, y0, 0, (uint16_t*)o);
    if ((v & 6) == 2) k<uint16_t, 80, 32, 2><<<1, 128>>>(tm, x0, y0, 0, (uint16_t*)o);
    if ((v & 6) == 4) k<uint16_t, 64, 34, 2><<<1, 128>>>(tm, x0, y0, 0, (uint16_t*)o);
    if ((v & 6) == 6) k<uint16_t, 80, 34, 2><<<1, 128>>>(tm, x0, y0, 0, (uint16_t*)o);
  }
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<unsigned char> ho(80 * 34 * 4);
  cudaMemcpy(ho.data(), o, ho.size(), cudaMemcpyDeviceToHost);
  int ones = 0; for (auto c : ho) ones += c == 1;
  printf("variant %d: %s, bytes==1: %d\n", v, cudaGetErrorString(e), ones);
  return 0;
}
