// util.cu — small device utilities shared by the graph kernels (prefix sums).
#include "common.cuh"

namespace {
constexpr int SCAN_BLOCK = 1024;

__global__ void __launch_bounds__(SCAN_BLOCK)
scan_block_kernel(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out, int64_t* __restrict__ block_sums) {
  __shared__ int64_t warp_sums[32];
  const int64_t i = (int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t v = i < n ? (int64_t)in[i] : 0;
  int64_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int64_t s = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    warp_sums[lane] = s;
  }
  __syncthreads();
  const int64_t incl = x + (warp > 0 ? warp_sums[warp - 1] : 0);
  if (i < n) out[i] = incl - v;  // exclusive within block
  if (threadIdx.x == SCAN_BLOCK - 1) block_sums[blockIdx.x] = incl;
}
__global__ void __launch_bounds__(1024)
scan_sums_kernel(int64_t* __restrict__ block_sums, int64_t nb, int64_t* __restrict__ total) {
  // single block: sequential chunks of 1024 with a running carry
  __shared__ int64_t warp_sums[32];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = 0; base < nb; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int64_t v = i < nb ? block_sums[i] : 0;
    int64_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int64_t s = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t y = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += y;
      }
      warp_sums[lane] = s;
    }
    __syncthreads();
    const int64_t incl = x + (warp > 0 ? warp_sums[warp - 1] : 0) + carry;
    if (i < nb) block_sums[i] = incl - v;  // exclusive offset of block i
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
__global__ void scan_add_kernel(int64_t* __restrict__ out, int64_t n, const int64_t* __restrict__ block_sums,
                                const int64_t* __restrict__ total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += block_sums[i / SCAN_BLOCK];
  if (i == n) out[n] = *total;
}
}  // namespace

int32_t sb2_scan_i32_to_i64(sb2_ctx* ctx, const int32_t* d_in, int64_t n, int64_t* d_out) {
  ScratchScope scr(ctx);
  const int64_t nb = ceil_div64(n > 0 ? n : 1, SCAN_BLOCK);
  int64_t* sums;
  SB2_TRY(scr.alloc(&sums, (size_t)nb + 1));
  scan_block_kernel<<<(unsigned)nb, SCAN_BLOCK, 0, ctx->stream>>>(d_in, n, d_out, sums);
  SB2_LAUNCH_CHECK(ctx);
  scan_sums_kernel<<<1, 1024, 0, ctx->stream>>>(sums, nb, sums + nb);
  SB2_LAUNCH_CHECK(ctx);
  scan_add_kernel<<<(unsigned)ceil_div64(n + 1, 256), 256, 0, ctx->stream>>>(d_out, n, sums, sums + nb);
  SB2_LAUNCH_CHECK(ctx);
  return SB2_OK;
}
