// micro-benchmark (round 2): issue rate per SM sub-partition of the max instructions the kNN epilogue is made of.
// 4 warps per SMSP (16 warps, one CTA per SM), each runs ILP-8 chains; reports cycles per warp-instruction per SMSP.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu_max_rate alu_max_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
__device__ __forceinline__ float fmax3(float a, float b, float c) { float d; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fmax2(float a, float b) { float d; asm volatile("max.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ int imax3(int a, int b, int c) { int d; asm volatile("max.s32 %0, %1, %2;\n\tmax.s32 %0, %0, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ unsigned hmax2(unsigned a, unsigned b) { unsigned d; asm volatile("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ unsigned bmax2(unsigned a, unsigned b) { unsigned d; asm volatile("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float ffmasat(float a, float c) { float d; asm volatile("fma.rn.sat.f32 %0, %1, 0f4B000000, %2;" : "=f"(d) : "f"(a), "f"(c)); return d; }
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, float seed, unsigned long long* out) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = seed + i + threadIdx.x;
  const float y = seed * 0.5f, z = seed * 0.25f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) x[i] = fmax3(x[i], y, z);
      if (MODE == 1) x[i] = fmax2(x[i], y);
      if (MODE == 2) x[i] = __int_as_float(imax3(__float_as_int(x[i]), __float_as_int(y), __float_as_int(z)));
      if (MODE == 3) x[i] = __uint_as_float(hmax2(__float_as_uint(x[i]), __float_as_uint(y)));
      if (MODE == 4) x[i] = __uint_as_float(bmax2(__float_as_uint(x[i]), __float_as_uint(y)));
      if (MODE == 5) x[i] = ffma(x[i], y, z);
      if (MODE == 6) x[i] = ffmasat(x[i], z);
      if (MODE == 7) { x[i] = fmax3(x[i], y, z); x[i] = ffma(x[i], y, z); }   // one ALU + one FMA-pipe op: do they dual-issue?
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 12345.f) out[1] = 1;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
}
template <int MODE> void run(const char* name, int ops_per_iter) {
  unsigned long long* out; cudaMalloc(&out, 16);
  const int iters = 20000;
  k<MODE><<<148, 512>>>(iters, 1.5f, out); k<MODE><<<148, 512>>>(iters, 1.5f, out);
  cudaDeviceSynchronize();
  unsigned long long h = 0; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  // 4 warps per SMSP, each issues iters * ops_per_iter warp-instructions
  printf("%-28s %6.2f cycles per warp-instruction per SMSP (4 warps/SMSP, ILP 8)\n", name, (double)h / ((double)iters * ops_per_iter * 4));
  cudaFree(out);
}
int main() {
  run<0>("FMNMX3 (max.f32 3-input)", 8);
  run<1>("FMNMX  (max.f32 2-input)", 8);
  run<2>("max.s32 x2 (3-input int)", 8);
  run<3>("HMNMX2 (max.f16x2)", 8);
  run<4>("max.bf16x2", 8);
  run<5>("FFMA", 8);
  run<6>("FFMA.SAT imm", 8);
  run<7>("FMNMX3 + FFMA pair", 16);
  return 0;
}
