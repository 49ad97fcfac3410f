// Micro-benchmark: issue rate of scalar and packed fp32 FMA forms on sm_100a (warp instructions per cycle per SM
// sub-partition). 16 independent accumulator chains per thread, 4..16 warps per sub-partition.
#include <cuda_runtime.h>
#include <cstdio>
typedef float2 f2;
#define NACC 16
#define ITERS 2048

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float* out, const float* in, float s0, float s1, long long* cyc) {
  float a[NACC], b[NACC], c[NACC];
  f2 A[NACC / 2], B[NACC / 2], Cc[NACC / 2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { a[i] = in[threadIdx.x + 32 * i]; b[i] = in[threadIdx.x + 32 * i + 7]; c[i] = 0.f; }
#pragma unroll
  for (int i = 0; i < NACC / 2; ++i) { A[i] = make_float2(a[2 * i], a[2 * i + 1]); B[i] = make_float2(b[2 * i], b[2 * i + 1]); Cc[i] = make_float2(0.f, 0.f); }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) {          // FFMA, 3 distinct registers
#pragma unroll
      for (int i = 0; i < NACC; ++i) c[i] = __fmaf_rn(a[i], b[i], c[i]);
    } else if (MODE == 1) {   // FFMA, one operand shared by all (register reuse)
#pragma unroll
      for (int i = 0; i < NACC; ++i) c[i] = __fmaf_rn(a[0], b[i], c[i]);
    } else if (MODE == 2) {   // FFMA, uniform/constant operand
#pragma unroll
      for (int i = 0; i < NACC; ++i) c[i] = __fmaf_rn(s0, b[i], c[i]);
    } else if (MODE == 3) {   // FFMA2, 3 distinct pairs  (NACC/2 instr = NACC lane-FMAs)
#pragma unroll
      for (int i = 0; i < NACC / 2; ++i) Cc[i] = __ffma2_rn(A[i], B[i], Cc[i]);
    } else if (MODE == 4) {   // FFMA2, broadcast scalar register
#pragma unroll
      for (int i = 0; i < NACC / 2; ++i) Cc[i] = __ffma2_rn(make_float2(a[0], a[0]), B[i], Cc[i]);
    } else if (MODE == 5) {   // FFMA2, uniform broadcast
#pragma unroll
      for (int i = 0; i < NACC / 2; ++i) Cc[i] = __ffma2_rn(make_float2(s0, s0), B[i], Cc[i]);
    } else if (MODE == 6) {   // FFMA2, swapped operand
#pragma unroll
      for (int i = 0; i < NACC / 2; ++i) Cc[i] = __ffma2_rn(A[i], make_float2(B[i].y, B[i].x), Cc[i]);
    } else if (MODE == 7) {   // FFMA2, same pair twice (E*E)
#pragma unroll
      for (int i = 0; i < NACC / 2; ++i) Cc[i] = __ffma2_rn(A[i], A[i], Cc[i]);
    } else if (MODE == 8) {   // FMUL2 + FADD2 (2 instr per pair)
#pragma unroll
      for (int i = 0; i < NACC / 2; ++i) Cc[i] = __fadd2_rn(__fmul2_rn(A[i], B[i]), Cc[i]);
    } else if (MODE == 9) {   // FMUL + FADD scalar
#pragma unroll
      for (int i = 0; i < NACC; ++i) c[i] = __fadd_rn(__fmul_rn(a[i], b[i]), c[i]);
    } else if (MODE == 10) {  // mixed: FFMA2 (3 pairs) interleaved with FFMA uniform
#pragma unroll
      for (int i = 0; i < NACC / 2; ++i) { Cc[i] = __ffma2_rn(A[i], B[i], Cc[i]); c[i] = __fmaf_rn(s0, b[i], c[i]); }
    } else if (MODE == 11) {  // FFMA 3-reg where the multiplicands rotate through few registers (J_i * J_c pattern)
#pragma unroll
      for (int i = 0; i < NACC; ++i) c[i] = __fmaf_rn(a[i & 3], a[(i >> 2) & 3], c[i]);
    } else if (MODE == 12) {  // FFMA2 pattern of the accumulate: few E pairs, many accumulators
#pragma unroll
      for (int i = 0; i < NACC / 2; ++i) Cc[i] = __ffma2_rn(A[i & 1], (i & 2) ? make_float2(A[2 + (i >> 2)].y, A[2 + (i >> 2)].x) : A[2 + (i >> 2)], Cc[i]);
    }
  }
  const long long t1 = clock64();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) r += c[i];
#pragma unroll
  for (int i = 0; i < NACC / 2; ++i) r += Cc[i].x + Cc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_iter, float* out, float* in, long long* cyc) {
  for (int threads : {128, 256, 512, 1024}) {
    k<MODE><<<148, threads>>>(out, in, 1.0001f, 0.9999f, cyc);
    k<MODE><<<148, threads>>>(out, in, 1.0001f, 0.9999f, cyc);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double warps_per_smsp = threads / 32 / 4.0;
    const double ipc = (double)instr_per_iter * ITERS * warps_per_smsp / (double)c;
    printf("%-44s warps/SMSP %4.1f  cycles %9lld  warp-instr/clk/SMSP %.3f  lane-FMA/clk/SM %.1f\n", name, warps_per_smsp, c, ipc,
           ipc * 4 * 32 * (double)NACC / instr_per_iter);
  }
}

int main() {
  float *out, *in; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&in, 4096 * 4); cudaMalloc(&cyc, 8);
  cudaMemset(in, 0, 4096 * 4);
  run<0>("FFMA r,r,r distinct", NACC, out, in, cyc);
  run<1>("FFMA shared-reg,r,r", NACC, out, in, cyc);
  run<2>("FFMA uniform,r,r", NACC, out, in, cyc);
  run<3>("FFMA2 pair,pair,pair distinct", NACC / 2, out, in, cyc);
  run<4>("FFMA2 bcast-reg,pair,pair", NACC / 2, out, in, cyc);
  run<5>("FFMA2 bcast-uniform,pair,pair", NACC / 2, out, in, cyc);
  run<6>("FFMA2 pair,swap(pair),pair", NACC / 2, out, in, cyc);
  run<7>("FFMA2 pair,same pair,pair", NACC / 2, out, in, cyc);
  run<8>("FMUL2+FADD2", NACC, out, in, cyc);
  run<9>("FMUL+FADD", 2 * NACC, out, in, cyc);
  run<10>("FFMA2 distinct + FFMA uniform interleaved", NACC, out, in, cyc);
  run<11>("FFMA J_i*J_c pattern (few multiplicands)", NACC, out, in, cyc);
  run<12>("FFMA2 accumulate pattern (E pairs, swap)", NACC / 2, out, in, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
