// micro-benchmark (round-2 groundwork, NOT yet run): three questions about the tcgen05.mma issue path on sm_100a
//   (1) queue depth      : cycles at which the k-th back-to-back UTCHMMA of an idle tensor pipe returns to the issuer
//   (2) dispatch stall   : does a warp that sits on a blocked UTCHMMA slow down another warp of the SAME SM
//                          sub-partition (warp w and w+4) more than a warp of a different one (w+1)?
//   (3) hand-off latency : tcgen05.commit -> mbarrier phase visible to a spinning waiter, and plain
//                          mbarrier.arrive -> waiter, in cycles
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_dispatch_stall mma_dispatch_stall.cu
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
constexpr uint32_t IDESC = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);  // f16 x f16 -> f32, M = N = 128
__device__ __forceinline__ void mma_elect(uint32_t tm, uint32_t da_lo, uint32_t db_lo, uint32_t hi, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p, pe;\n.reg .b64 da, db;\nsetp.ne.b32 p, %5, 0;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\n"
      "elect.sync _|pe, 0xffffffff;\n@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n}\n" ::"r"(tm), "r"(da_lo),
      "r"(db_lo), "r"(hi), "r"(IDESC), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void commit_elect(uint64_t* bar) {
  asm volatile("{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ long long alu_work(int iters, float& sink) {  // dependent FMA chain: fixed issue-slot demand
  float x = sink;
  const long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < iters; ++i) x = fmaf(x, 1.000001f, 0.5f);
  const long long t1 = clock64();
  sink = x;
  return t1 - t0;
}

// mode 0: queue depth; mode 1: dispatch stall (issuer = warp 0); mode 2: hand-off latency
__global__ void __launch_bounds__(256, 1) k(int mode, int n_mma, int alu_iters, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bars[4];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int b = 0; b < 4; ++b) mbar_init(&bars[b], 1);
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
  const uint64_t d0 = umma_desc(0);
  const uint32_t hi = (uint32_t)(d0 >> 32);
  const uint32_t a_lo = (uint32_t)d0 | (smem_u32(smem) >> 4), b_lo = (uint32_t)d0 | (smem_u32(smem + 32768) >> 4);
  float sink = (float)lane;
  if (mode == 0) {
    if (warp == 0) {
      long long t[33];
      t[0] = clock64();
      for (int i = 0; i < 32; ++i) {
        if (i < n_mma) mma_elect((uint32_t)((i & 1) * 256), a_lo, b_lo, hi, 1u);
        t[i + 1] = clock64();
      }
      commit_elect(&bars[0]);
      mbar_wait(&bars[0], 0);
      const long long tend = clock64();
      if (lane == 0) {
        for (int i = 0; i < 32; ++i) out[i] = t[i + 1] - t[0];
        out[32] = tend - t[0];
      }
    }
  } else if (mode == 1) {
    // warps 0 (issuer), 4 (same sub-partition), 1 (another sub-partition); n_mma == 0 gives the undisturbed baseline
    if (warp == 0) {
      for (int i = 0; i < n_mma; ++i) mma_elect((uint32_t)((i & 1) * 256), a_lo, b_lo, hi, 1u);
      commit_elect(&bars[0]);
      mbar_wait(&bars[0], 0);
    } else if (warp == 4 || warp == 1) {
      const long long d = alu_work(alu_iters, sink);
      if (lane == 0) out[warp == 4 ? 0 : 1] = d;
    }
  } else {
    // warp 0: t0, one MMA, commit(bars[0]); later arrive(bars[1]).  warp 1 spins on both and stamps.
    if (warp == 0) {
      const long long t0 = clock64();
      mma_elect(0u, a_lo, b_lo, hi, 0u);
      commit_elect(&bars[0]);
      const long long t1 = clock64();
      mbar_wait(&bars[2], 0);  // waiter saw the commit
      const long long t2 = clock64();
      if (lane == 0) mbar_arrive(&bars[1]);
      if (lane == 0) { out[0] = t0; out[1] = t1; out[2] = t2; }
    } else if (warp == 1) {
      mbar_wait(&bars[0], 0);
      const long long s0 = clock64();
      if (lane == 0) mbar_arrive(&bars[2]);
      mbar_wait(&bars[1], 0);
      const long long s1 = clock64();
      if (lane == 0) { out[3] = s0; out[4] = s1; }
    }
  }
  if (sink == 12345.678f) out[63] = 1;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

int main() {
  long long* out;
  cudaMalloc(&out, 64 * 8);
  const size_t smem = 64 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long h[64];
  auto run = [&](int mode, int n_mma, int alu) {
    cudaMemset(out, 0, 64 * 8);
    k<<<1, 256, smem>>>(mode, n_mma, alu, out);
    k<<<1, 256, smem>>>(mode, n_mma, alu, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  };
  run(0, 32, 0);
  printf("(1) cycles until the k-th back-to-back MMA (128x128x16, 64 cycles each) returned to the issuer; all 32 done at %lld\n   ", h[32]);
  for (int i = 0; i < 32; ++i) printf(" %lld", h[i]);
  printf("\n");
  for (int n : {0, 64, 256}) {
    run(1, n, 20000);
    printf("(2) %3d MMAs in flight from warp 0: 20000-FMA chain takes %lld cycles on warp 4 (same sub-partition), %lld on warp 1\n", n, h[0], h[1]);
  }
  run(2, 1, 0);
  printf("(3) issue+commit %lld cycles; commit visible to a spinning waiter %lld cycles after issue start (MMA itself: 64); "
         "arrive -> waiter %lld cycles\n", h[1] - h[0], h[3] - h[0], h[4] - h[2]);
  cudaFree(out);
  return 0;
}
