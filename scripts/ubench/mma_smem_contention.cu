// micro-benchmark (round 2): what slows tcgen05.mma in the kNN sweep?  One issuer warp runs the sweep's MMA pattern
// (per visit: 2 query halves x nks k-steps of M128 N128 K16 against ONE staged candidate image) while
//   - a producer warp streams 16 KB candidate images global -> shared through a cp.async.bulk ring (TMA writes), and/or
//   - 8 epilogue warps run tcgen05.ld.32x32b.x32 loops over the accumulators (TMEM reads),
// with operand A read from shared memory (SS) or from TMEM (TS).  Reports cycles per visit (8 MMAs = 512 tensor cycles).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_smem_contention mma_smem_contention.cu
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
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void commit_elect(uint64_t* bar) {
  asm volatile("{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
constexpr int NST = 6;
// flags: bit0 = TMA producer on, bit1 = TMEM-load warps on, bit2 = A from TMEM
template <int N>
__global__ void __launch_bounds__(352, 1) k(int visits, int nks, int flags, const unsigned char* __restrict__ gsrc, size_t gbytes,
                                            unsigned long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t full[NST], empty[NST], done;
  __shared__ volatile int stop;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t stage_b = (uint32_t)N * 64u * 2u;  // one candidate image: N rows x K=64 halves
  unsigned char* As = smem;                          // 2 x 16 KB
  unsigned char* Bs = smem + 32768;                  // NST stages
  for (int i = threadIdx.x; i < (32768 + NST * (int)stage_b) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&done, 1);
    stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const bool tma = flags & 1, ldtm = flags & 2, ts = flags & 4;
  if (warp == 0) {
    if (tma && lane == 0) {
      size_t off = (size_t)blockIdx.x * 1048576u % gbytes;
      int s = 0; uint32_t ph = 1; bool first = true;
      for (int v = 0; v < visits; ++v) {
        if (!first) mbar_wait(&empty[s], ph);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(stage_b) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(Bs + (size_t)s * stage_b)),
                     "l"(gsrc + off), "r"(stage_b), "r"(smem_u32(&full[s])) : "memory");
        off += stage_b; if (off + stage_b > gbytes) off = 0;
        if (++s == NST) { s = 0; ph ^= 1u; first = false; }
      }
    }
  } else if (warp == 1) {
    const uint64_t d0 = umma_desc(0);
    const uint32_t hi = (uint32_t)(d0 >> 32);
    const uint32_t a_lo0 = (uint32_t)d0 | (smem_u32(As) >> 4);
    const uint32_t b_lo0 = (uint32_t)d0 | (smem_u32(Bs) >> 4);
    int s = 0; uint32_t ph = 0;
    const long long t0 = clock64();
    for (int v = 0; v < visits; ++v) {
      if (tma) mbar_wait(&full[s], ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int h = 0; h < (N == 128 ? 2 : 1); ++h) {
        const uint32_t tm = (uint32_t)((v & 1) * 256 + h * 128);
        uint32_t da = a_lo0 + (uint32_t)h * (16384u >> 4), db = b_lo0 + (uint32_t)s * (stage_b >> 4);
        for (int j = 0; j < nks; ++j) {
          if (ts) {
            // A tile in TMEM: 8 columns (16 halves) per k-step, parked past the accumulators is impossible with 512 accumulator
            // columns, so this variant overlays columns 448.. (garbage values are fine for a rate measurement)
            asm volatile("{\n.reg .pred p, pe;\n.reg .b64 db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 db, {%2, %3};\nelect.sync _|pe, 0xffffffff;\n"
                         "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %5, p;\n}\n" ::"r"(tm), "r"(448u + (uint32_t)h * 32u + (uint32_t)(j & 3) * 8u), "r"(db), "r"(hi), "r"(j), "r"(idesc<N>()) : "memory");
          } else {
            asm volatile("{\n.reg .pred p, pe;\n.reg .b64 da, db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\nelect.sync _|pe, 0xffffffff;\n"
                         "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(tm), "r"(da), "r"(db), "r"(hi), "r"(j), "r"(idesc<N>()) : "memory");
          }
          da += 256; db += 256;
        }
      }
      if (tma) commit_elect(&empty[s]);
      if (++s == NST) { s = 0; ph ^= 1u; }
    }
    commit_elect(&done);
    mbar_wait(&done, 0);
    const long long t1 = clock64();
    if (lane == 0) { out[blockIdx.x] = (unsigned long long)(t1 - t0); stop = 1; }
  } else if (warp >= 2 && warp < 10) {
    if (ldtm) {
      uint32_t v[32];
      uint32_t acc = 0;
      unsigned long long n = 0;
      const uint32_t lane_base = ((uint32_t)((warp & 3) * 32) << 16);
      while (!stop) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                       "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                         "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                         "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                       : "r"(lane_base + (uint32_t)(((warp - 2) >> 2) * 128 + c * 32)));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 32; ++i) acc ^= v[i];
          ++n;
        }
      }
      if (acc == 0x12345u) out[200] = n;
      if (lane == 0) out[148 + (blockIdx.x == 0 ? warp : 0)] = n;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}
template <int N>
void run(const char* name, int nks, int flags, int ctas, const unsigned char* g, size_t gbytes) {
  unsigned long long* out;
  cudaMalloc(&out, 256 * 8);
  cudaMemset(out, 0, 256 * 8);
  const int visits = 6000;
  const size_t smem = 32768 + NST * (size_t)N * 128;
  cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<N><<<ctas, 352, smem>>>(visits, nks, flags, g, gbytes, out);
  k<N><<<ctas, 352, smem>>>(visits, nks, flags, g, gbytes, out);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h[256];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  double mx = 0;
  for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? (double)h[i] : mx;
  const double cyc = mx / visits;
  const double mma_per_visit = (N == 128 ? 2 : 1) * nks;
  printf("%-34s nks=%d ctas=%3d  %7.1f cyc/visit  %6.1f cyc/MMA (ideal %d)  ldtm/visit(warp2)=%.2f (%s)\n", name, nks, ctas, cyc, cyc / mma_per_visit, N / 2,
         (double)h[148 + 2] / visits, cudaGetErrorString(e));
  cudaFree(out);
}
int main() {
  unsigned char* g;
  const size_t gbytes = 64u << 20;  // L2-resident source
  cudaMalloc(&g, gbytes);
  cudaMemset(g, 0x3c, gbytes);
  for (int ctas : {1, 148}) {
    run<128>("SS", 4, 0, ctas, g, gbytes);
    run<128>("SS + TMA ring", 4, 1, ctas, g, gbytes);
    run<128>("SS + TMEM loads", 4, 2, ctas, g, gbytes);
    run<128>("SS + TMA + TMEM loads", 4, 3, ctas, g, gbytes);
    run<128>("TS", 4, 4, ctas, g, gbytes);
    run<128>("TS + TMA ring", 4, 5, ctas, g, gbytes);
    run<128>("TS + TMA + TMEM loads", 4, 7, ctas, g, gbytes);
    run<256>("SS N256 (1 half)", 4, 0, ctas, g, gbytes);
    run<256>("SS N256 + TMA ring", 4, 1, ctas, g, gbytes);
    run<256>("SS N256 + TMA + TMEM loads", 4, 3, ctas, g, gbytes);
    run<256>("TS N256 + TMA + TMEM loads", 4, 7, ctas, g, gbytes);
  }
  return 0;
}
