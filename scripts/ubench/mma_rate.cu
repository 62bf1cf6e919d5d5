// micro-benchmark: tcgen05.mma issue/execute rate per SM: SS vs TS operand A, N = 128 / 256 (sm_100a, kind::f16, M = 128)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr uint32_t LBO = 2048, SBO = 128;
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((LBO >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((SBO >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
template <int N> __host__ __device__ constexpr uint32_t idesc() { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24); }
template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) k(int iters, int nks, unsigned long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0) {
    const uint64_t d0 = umma_desc(0);
    const uint32_t hi = (uint32_t)(d0 >> 32);
    const uint32_t a_lo0 = (uint32_t)d0 | (smem_u32(smem) >> 4);
    const uint32_t b_lo0 = (uint32_t)d0 | (smem_u32(smem + 65536) >> 4);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t tm = (uint32_t)((i & 1) * 256);
      uint32_t da = a_lo0, db = b_lo0;
      for (int j = 0; j < nks; ++j) {
        if (TS) {
          asm volatile("{\n.reg .pred p, pe;\n.reg .b64 db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 db, {%2, %3};\nelect.sync _|pe, 0xffffffff;\n"
                       "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %5, p;\n}\n" ::"r"(tm), "r"(448u + (uint32_t)(j & 3) * 8u), "r"(db), "r"(hi), "r"(j), "r"(idesc<N>()) : "memory");
        } else {
          asm volatile("{\n.reg .pred p, pe;\n.reg .b64 da, db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\nelect.sync _|pe, 0xffffffff;\n"
                       "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(tm), "r"(da), "r"(db), "r"(hi), "r"(j), "r"(idesc<N>()) : "memory");
        }
        da += 256; db += 256;
      }
    }
    asm volatile("{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}
template <int N, bool TS>
void run(const char* name, int nks, int ctas) {
  unsigned long long* out;
  cudaMalloc(&out, 148 * 8);
  const int iters = 4000;
  const size_t smem = 160 * 1024;
  cudaFuncSetAttribute(k<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<N, TS><<<ctas, 128, smem>>>(iters, nks, out);
  k<N, TS><<<ctas, 128, smem>>>(iters, nks, out);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h[148];
  cudaMemcpy(h, out, sizeof(unsigned long long) * ctas, cudaMemcpyDeviceToHost);
  const double cyc = (double)h[0] / ((double)iters * nks);
  printf("%-26s nks=%2d ctas=%3d  %7.1f cyc/MMA  -> %6.0f flops/clk/SM (%s)\n", name, nks, ctas, cyc, 2.0 * 128 * N * 16 / cyc, cudaGetErrorString(e));
  cudaFree(out);
}
int main() {
  for (int ctas : {1, 148}) {
    run<128, false>("SS  M128 N128 K16", 4, ctas);
    run<128, false>("SS  M128 N128 K16", 10, ctas);
    run<256, false>("SS  M128 N256 K16", 4, ctas);
    run<128, true>("TS  M128 N128 K16", 4, ctas);
    run<256, true>("TS  M128 N256 K16", 4, ctas);
  }
  return 0;
}
