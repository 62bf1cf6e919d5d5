// micro-benchmark: tcgen05.ld throughput per SM for a few shapes / warp counts (sm_100a)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define LD_X(N, SHAPE, REGS, ...) asm volatile("tcgen05.ld.sync.aligned." SHAPE ".x" #N ".b32 {" REGS "}, [%" #__VA_ARGS__ "];" ::: "memory")
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, unsigned long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const uint32_t a = base + (uint32_t)((i * 32) & 255);
    if (MODE == 0) {  // 32x32b.x32
      uint32_t r[32];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
                     "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31])
                   : "r"(a));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc ^= r[0] ^ r[31];
    } else if (MODE == 1) {  // 32x32b.x64
      uint32_t r[64];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
                   : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
                     "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]),
                     "=r"(r[32]),"=r"(r[33]),"=r"(r[34]),"=r"(r[35]),"=r"(r[36]),"=r"(r[37]),"=r"(r[38]),"=r"(r[39]),"=r"(r[40]),"=r"(r[41]),"=r"(r[42]),"=r"(r[43]),"=r"(r[44]),"=r"(r[45]),"=r"(r[46]),"=r"(r[47]),
                     "=r"(r[48]),"=r"(r[49]),"=r"(r[50]),"=r"(r[51]),"=r"(r[52]),"=r"(r[53]),"=r"(r[54]),"=r"(r[55]),"=r"(r[56]),"=r"(r[57]),"=r"(r[58]),"=r"(r[59]),"=r"(r[60]),"=r"(r[61]),"=r"(r[62]),"=r"(r[63])
                   : "r"(a & ~63u));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc ^= r[0] ^ r[63];
    } else if (MODE == 2) {  // 16x256b.x8 (32 regs) : lanes 0-15 of the warp's quadrant + ... one instruction covers 16 lanes x 8*256 bits
      uint32_t r[32];
      asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
                     "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31])
                   : "r"(a));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc ^= r[0] ^ r[31];
    } else if (MODE == 3) {  // two x32 loads in flight before the wait
      uint32_t r[32], q[32];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
                     "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31])
                   : "r"(a));
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(q[0]),"=r"(q[1]),"=r"(q[2]),"=r"(q[3]),"=r"(q[4]),"=r"(q[5]),"=r"(q[6]),"=r"(q[7]),"=r"(q[8]),"=r"(q[9]),"=r"(q[10]),"=r"(q[11]),"=r"(q[12]),"=r"(q[13]),"=r"(q[14]),"=r"(q[15]),
                     "=r"(q[16]),"=r"(q[17]),"=r"(q[18]),"=r"(q[19]),"=r"(q[20]),"=r"(q[21]),"=r"(q[22]),"=r"(q[23]),"=r"(q[24]),"=r"(q[25]),"=r"(q[26]),"=r"(q[27]),"=r"(q[28]),"=r"(q[29]),"=r"(q[30]),"=r"(q[31])
                   : "r"(a ^ 32u));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc ^= r[0] ^ q[31];
    } else if (MODE == 4) {  // 32x32b.x32 with .pack::16b (two 16-bit columns per register): 64 columns per instruction
      uint32_t r[32];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.pack::16b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
                     "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31])
                   : "r"(a & ~63u));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc ^= r[0] ^ r[31];
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}
template <int MODE>
void run(const char* name, int warps, int regs_per_instr, int instr_per_iter) {
  unsigned long long* out; uint32_t* sink;
  cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 4);
  const int iters = 20000;
  k<MODE><<<148, warps * 32>>>(iters, out, sink);
  k<MODE><<<148, warps * 32>>>(iters, out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  double bytes = (double)iters * instr_per_iter * warps * 32.0 * regs_per_instr * 4.0;
  printf("%-34s warps=%2d  %8.1f B/clk/SM  (%s)\n", name, warps, bytes / (double)h[0], cudaGetErrorString(e));
  cudaFree(out); cudaFree(sink);
}
int main() {
  for (int w : {1, 2, 4, 8, 16}) run<0>("32x32b.x32", w, 32, 1);
  for (int w : {4, 8}) run<1>("32x32b.x64", w, 64, 1);
  for (int w : {4, 8}) run<2>("16x256b.x8", w, 32, 1);
  for (int w : {4, 8, 16}) run<3>("2 x 32x32b.x32 in flight", w, 32, 2);
  for (int w : {4, 8}) run<4>("32x32b.pack::16b.x32 (regs)", w, 32, 1);
  return 0;
}
