// knn_tc.cu — tensor-core sweeps of the exact kNN (sm_100a: tcgen05.mma + TMEM + bulk-copy ring).
//
// Contract (shared with knn_pass1_kernel in knn.cu): per query keep the list_m (32 or 64) proposals whose score
//     s(q,c) = q.c - |c|^2/2        (d^2 = |q|^2 - 2 s)
// is largest, plus a rigorous bound on the rounding error of s, so that knn_rescore_kernel can certify the exact
// fp64 top-k.  The GEMM-shaped middle term runs on the 5th-gen tensor cores, fp32 accumulation in TMEM, in one of
// two operand formats (coordinates scaled by a power of two so that max|x| is in [100, 200)):
//   terms = 1  fp16 operands            A = [q_hi | 1 1 1],             B = [c_hi | h0 h1 h2]            K = d+3
//   terms = 3  split fp16 (22 bits)     A = [q_hi | q_hi | q_lo | 1 1 1], B = [c_hi | c_lo | c_hi | h0 h1 h2]  K = 3d+3
// with hi = fp16(x), lo = fp16(x - hi) and h0+h1+h2 = -|c|^2/2 (three-way fp16 split); K is padded to a multiple
// of 16 (d = 50: 64 and 160).  knn.cu runs terms = 1 first and re-sweeps the rows it cannot certify with terms = 3.
//
// knn_tc_prep_kernel   X[n,d] -> per 128-point tile one A image and/or one B image, stored exactly as the UMMA
//                      "no-swizzle, K-major" shared-memory layout wants them (8x8 fp16 core matrices, K-chunk-major),
//                      so a tile (or a K-slice of it) is staged by ONE 1-D bulk (TMA) copy; optionally gathers the
//                      rows (re-sweep) and records each point's fp16 residual norm |x - fp16(x)| (terms = 1 bound).
// knn_pass1_tc_kernel  CTA = 256 queries (two M=128 halves; 128 queries for wide K) x all candidate tiles (N=128).
//                      warp 0     : producer, cp.async.bulk ring on mbarriers (as many stages as shared memory holds)
//                      warps 1,10 : MMA issuers, one per query half: whole-warp loop, elect.sync inside the asm block,
//                                   operands on the uniform datapath -> back-to-back UTCHMMA; tcgen05.commit releases
//                                   the smem stage and publishes (accumulator buffer, half); warp 1 owns TMEM
//                                   (alloc, teardown - no CTA-wide barrier after the role branches)
//                      warps 2-9  : epilogue; thread <-> one query row (TMEM lane); two tcgen05.ld of 32 columns in
//                                   flight, a two-level FMNMX3 tree + one compare per 32 values against the row's
//                                   threshold; survivors go to the row's register-resident list (ids to global
//                                   memory, write-only); the accumulator slab is released after its second load
//                                   round trip.  The first visits sweep every 16th tile and only keep the 6 largest
//                                   chunk maxima per row: the 6th seeds the threshold (rank ~96), so the sweep proper
//                                   sees ~60 insertions per row instead of ~340.
// Candidate images are re-read by every CTA from L2 (16 KB per 256x128 score tile for d = 50, 97 % L2 hits).
// DESIGN.md section 4 has the measurements behind each of these choices.
#include <cuda_fp16.h>
#include <float.h>

#include "common.cuh"
#include "knn_internal.cuh"

namespace {

constexpr int TM = 128;           // rows per operand tile (UMMA M and N)
constexpr int MAXST = 8;          // upper bound of the candidate ring depth (runtime nstage <= MAXST)
constexpr int BAR_BYTES = 1024;   // mbarriers + TMEM slot IN FRONT of the operand images (a fixed address: see the MMA warp)
constexpr int TC_THREADS = 352;   // 11 warps: producer, MMA issuer (half 0), 8 epilogue warps, MMA issuer (half 1)
constexpr uint32_t SBO = 128;     // bytes between 8-row groups (core matrices contiguous)
constexpr uint32_t LBO = TM * 16; // bytes between K-chunks (8 fp16) : [kc][row-group][8][8]

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// UMMA shared-memory descriptor, no swizzle, K-major (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);            // start address   [0,14)
  d |= (uint64_t)((LBO >> 4) & 0x3FFFu) << 16;         // leading byte offset (K direction) [16,30)
  d |= (uint64_t)((SBO >> 4) & 0x3FFFu) << 32;         // stride byte offset (M/N direction) [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell) [46,48)
  return d;                                            // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}
// instruction descriptor: D=F32, A=B=F16, both K-major, N=128, M=128 (cute::UMMA::InstrDescriptor bit layout)
constexpr uint32_t IDESC = (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

// MMA issue.  The whole warp runs the (warp-uniform) issue loop and the single issuing thread is chosen INSIDE the
// asm block (elect.sync): ptxas then keeps descriptors, TMEM addresses and loop counters on the uniform datapath
// and emits back-to-back UTCHMMA.  (Issuing from an `if (lane == 0)` branch makes every operand "divergent": each
// MMA is then wrapped in an ELECT / R2UR.BROADCAST waterfall of ~20 dependent instructions, ~130 cycles per MMA -
// twice the 64 cycles the tensor pipe needs for it.)  Descriptors are (lo, hi) halves: only lo moves along K.
template <bool ACC>
__device__ __forceinline__ void umma_f16_elect(uint32_t tmem_d, uint32_t da_lo, uint32_t db_lo, uint32_t desc_hi) {
  asm volatile(
      "{\n"
      ".reg .pred p, pe;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %5, 0;\n"
      "mov.b64 da, {%1, %3};\n"
      "mov.b64 db, {%2, %3};\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(da_lo), "r"(db_lo), "r"(desc_hi), "r"(IDESC), "n"(ACC ? 1 : 0)
      : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred pe;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
[[maybe_unused]] __device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(IDESC), "r"(accumulate)
      : "memory");
}
[[maybe_unused]] __device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// scale: power of two s with s*R in [100, 200]  ->  fp16 range is safe for coordinates and for s^2 R^2/2
__device__ __forceinline__ float tc_scale_from_maxnorm(unsigned int maxnorm_bits) {
  const float R = sqrtf(__uint_as_float(maxnorm_bits));
  if (!(R > 0.0f) || !isfinite(R)) return 1.0f;
  int e = (int)floorf(log2f(200.0f / R));
  e = max(-60, min(60, e));
  return exp2f((float)e);
}

__global__ void __launch_bounds__(256)
knn_tc_prep_kernel(const float* __restrict__ X, int64_t n, int d, int kpad, int terms,
                   const unsigned int* __restrict__ maxnorm_bits, const int32_t* __restrict__ gather, int64_t gather_base,
                   __half* __restrict__ Aimg, __half* __restrict__ Bimg, float* __restrict__ inv_s2,
                   float* __restrict__ dnorm, unsigned int* __restrict__ dmax_bits) {
  // dnorm[p] (optional) = |x_p - fp16(x_p)| in the units of X, rounded up; *dmax_bits = its maximum (float bits)
  // terms = 3: A = [hi | hi | lo | 1 1 1], B = [hi | lo | hi | h0 h1 h2]   (22-bit operands)
  // terms = 1: A = [hi | 1 1 1],           B = [hi | h0 h1 h2]             (11-bit operands, the fast first tier)
  // gather != nullptr: image row r of tile t is point gather_base + gather[t*128 + r] (rows past n are zero rows);
  // Aimg / Bimg may each be null (that image is not written)
  const int64_t t = blockIdx.x;
  const float s = tc_scale_from_maxnorm(*maxnorm_bits);
  if (t == 0 && threadIdx.x == 0) *inv_s2 = 1.0f / (s * s);
  __shared__ __half hn3[TM][3];
  // per-row -|x|^2 s^2 / 2, three-way fp16 split
  const int kd = terms * d;  // coordinates on the K axis before the three norm slots
  if (threadIdx.x < TM) {
    const int64_t p = t * TM + threadIdx.x;
    __half h0 = __float2half_rn(-60000.0f), h1 = __float2half_rn(0.0f), h2 = __float2half_rn(0.0f);
    if (p < n) {
      double acc = 0.0, dacc = 0.0;
      const int64_t ps = gather ? gather_base + gather[p] : p;
      for (int k = 0; k < d; ++k) {
        const float xs = X[ps * d + k] * s;
        const double v = (double)xs;
        acc += v * v;
        const double dl = v - (double)__half2float(__float2half_rn(xs));
        dacc += dl * dl;
      }
      if (dnorm) {
        const float dn = __double2float_ru(sqrt(dacc) * (1.0 + 1e-12) / (double)s);
        dnorm[p] = dn;
        atomicMax(dmax_bits, __float_as_uint(dn));
      }
      const double hn = -0.5 * acc;
      h0 = __float2half_rn((float)hn);
      const double r1 = hn - (double)__half2float(h0);
      h1 = __float2half_rn((float)r1);
      h2 = __float2half_rn((float)(r1 - (double)__half2float(h1)));
    }
    hn3[threadIdx.x][0] = h0; hn3[threadIdx.x][1] = h1; hn3[threadIdx.x][2] = h2;
  }
  __syncthreads();
  const int nkc = kpad / 8;
  const size_t img = (size_t)TM * kpad;  // halves per image
  __half* Aout = Aimg ? Aimg + (size_t)t * img : nullptr;
  __half* Bout = Bimg ? Bimg + (size_t)t * img : nullptr;
  for (int i = threadIdx.x; i < TM * nkc; i += blockDim.x) {
    const int kc = i / TM, r = i % TM;
    const int64_t p = t * TM + r;
    const int64_t ps = (gather && p < n) ? gather_base + gather[p] : p;
    __align__(16) __half a8[8];
    __align__(16) __half b8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = kc * 8 + j;
      __half av = __float2half_rn(0.0f), bv = av;
      if (p < n) {
        if (e < kd) {
          const int seg = e / d, k = e - seg * d;
          const float xs = X[ps * d + k] * s;
          const __half hi = __float2half_rn(xs);
          const __half lo = __float2half_rn(xs - __half2float(hi));
          av = seg == 2 ? lo : hi;   // A = [hi | hi | lo]
          bv = seg == 1 ? lo : hi;   // B = [hi | lo | hi]
        } else if (e < kd + 3) {
          av = __float2half_rn(1.0f);
          bv = hn3[r][e - kd];
        }
      } else if (e == kd) {
        bv = hn3[r][0];  // padding candidates: score -60000 (padding queries are all-zero rows)
      }
      a8[j] = av; b8[j] = bv;
    }
    const size_t off = ((size_t)kc * LBO + (size_t)(r >> 3) * SBO + (size_t)(r & 7) * 16) / 2;  // in halves
    if (Aout) *reinterpret_cast<uint4*>(Aout + off) = *reinterpret_cast<const uint4*>(a8);
    if (Bout) *reinterpret_cast<uint4*>(Bout + off) = *reinterpret_cast<const uint4*>(b8);
  }
}

// ------------------------------------------------------------------------------------------------
// Per-row proposal list: 32 scores in REGISTERS (4 groups of 8 with a running minimum per group), ids in
// global memory (write-only).  tau = min of the list = the score a candidate must beat.  All indexing is
// static (macro-expanded), so nothing spills and an insertion never waits on memory.
template <int NG>  // NG groups of 8 proposals: 32 (NG = 4) or 64 (NG = 8)
struct RowList {
  float ls[8 * NG];
  float gm[NG];
  float tau;
};
__device__ __forceinline__ float min8(const float* x) {
  return fminf(fminf(fminf(x[0], x[1]), fminf(x[2], x[3])), fminf(fminf(x[4], x[5]), fminf(x[6], x[7])));
}
#define SB2_GROUP_INSERT(G)                                                   \
  {                                                                           \
    bool placed = false;                                                      \
    _Pragma("unroll") for (int i = 0; i < 8; ++i) {                           \
      const bool hset = !placed && (L.ls[(G) * 8 + i] == L.tau);              \
      L.ls[(G) * 8 + i] = hset ? v : L.ls[(G) * 8 + i];                       \
      pos = hset ? ((G) * 8 + i) : pos;                                       \
      placed |= hset;                                                         \
    }                                                                         \
    L.gm[(G)] = min8(&L.ls[(G) * 8]);                                         \
  }
template <int NG>
__device__ __forceinline__ void list_insert(RowList<NG>& L, int32_t* __restrict__ id, float v, int32_t cand) {
  int pos = 0;
  if (L.gm[0] == L.tau) SB2_GROUP_INSERT(0)
  else if (L.gm[1] == L.tau) SB2_GROUP_INSERT(1)
  else if (L.gm[2] == L.tau) SB2_GROUP_INSERT(2)
  else if (NG == 4 || L.gm[3] == L.tau) SB2_GROUP_INSERT(3)
  else if (L.gm[4 % NG] == L.tau) SB2_GROUP_INSERT(4 % NG)
  else if (L.gm[5 % NG] == L.tau) SB2_GROUP_INSERT(5 % NG)
  else if (L.gm[6 % NG] == L.tau) SB2_GROUP_INSERT(6 % NG)
  else SB2_GROUP_INSERT(7 % NG)
  id[pos] = cand;
  float t = L.gm[0];
#pragma unroll
  for (int g = 1; g < NG; ++g) t = fminf(t, L.gm[g]);
  L.tau = t;
}
// examine one 32-column chunk of the row (values already in registers).  Fast path: a two-level 3-input max
// tree (four independent quarter maxima, so the FMNMX3 latencies overlap) and one compare.  Rare path (about
// 340 times per row over a 1.3M sweep): visit the quarters whose maximum beats tau, copy the quarter's eight
// values aside and pull its maxima out one by one.  The quarter loop is NOT unrolled: one copy of the extraction
// and insertion code per call site keeps the kernel inside the instruction caches.
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float max8u(const uint32_t* v) {
  return fmax3(fmax3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2])),
               fmax3(__uint_as_float(v[3]), __uint_as_float(v[4]), __uint_as_float(v[5])),
               fmaxf(__uint_as_float(v[6]), __uint_as_float(v[7])));
}
template <int NG>
__device__ __forceinline__ void scan_chunk(RowList<NG>& L, int32_t* __restrict__ id, const uint32_t (&v)[32], int32_t cand0,
                                           int32_t n_points) {
  const float h0 = max8u(&v[0]), h1 = max8u(&v[8]), h2 = max8u(&v[16]), h3 = max8u(&v[24]);
  if (fmaxf(fmax3(h0, h1, h2), h3) > L.tau) {
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
      float hq = q == 0 ? h0 : (q == 1 ? h1 : (q == 2 ? h2 : h3));
      if (hq > L.tau) {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = q == 0 ? v[i] : (q == 1 ? v[8 + i] : (q == 2 ? v[16 + i] : v[24 + i]));
        do {
          int j = 0;
          bool found = false;
          const uint32_t mb = __float_as_uint(hq);
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const bool hset = !found && (w[jj] == mb);
            j = hset ? jj : j;
            w[jj] = hset ? 0xff800000u : w[jj];  // knock the maximum out (-inf)
            found |= hset;
          }
          const int32_t cand = cand0 + q * 8 + j;
          if (cand < n_points) list_insert(L, id, hq, cand);
          hq = max8u(w);
        } while (hq > L.tau);
      }
    }
  }
}
constexpr int EST_R = 6;
// estimate phase: keep the EST_R largest chunk maxima (sorted descending), branch-free
__device__ __forceinline__ void est_chunk(float (&est)[EST_R], const uint32_t (&v)[32]) {
  float cm = fmaxf(fmax3(max8u(&v[0]), max8u(&v[8]), max8u(&v[16])), max8u(&v[24]));
#pragma unroll
  for (int i = 0; i < EST_R; ++i) {
    const float hi = fmaxf(est[i], cm);
    cm = fminf(est[i], cm);
    est[i] = hi;
  }
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

// QH = query halves per CTA (2: 256 queries, 1: 128 queries); nsplit = pieces the K axis of a candidate image is
// staged in (1 when a whole image fits a ring stage; 2 or 4 for wide embeddings, d up to 150).
// SB2_KNN_DBG=1: cycle stamps of CTA 0's hand-offs (visits [DBG_C0, DBG_C0 + DBG_N)) for pipeline analysis
__device__ long long g_knn_dbg[8 * 64];
__device__ long long g_knn_dbg2[3 * 8 * 16];  // [tfull seen | release | done][epilogue warp][visit - DBG_C0]
__device__ int g_knn_dbg_mode;  // DBG build only: 1 = epilogue skips the scans, 2 = epilogue also skips the TMEM loads
constexpr int DBG_C0 = 200, DBG_N = 64;
template <int NG, int QH, bool DBG = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
knn_pass1_tc_kernel(const __half* __restrict__ Aimg, const __half* __restrict__ Bimg, int nks, int nsplit, int nstage, uint32_t mma_part16,
                    uint32_t mma_tile16, int mma_visits, int64_t n_btiles,
                    int64_t n_est, int64_t est_stride,
                    int64_t qtile0, int64_t n_query, int32_t n_points, float* __restrict__ cand_score,
                    int32_t* __restrict__ cand_idx) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // nks = k-steps (16 columns of the K axis) per staged slice.  Everything the MMA warp derives its operands from is
  // a kernel parameter combined by add / multiply only: a division here would move the values to the vector
  // datapath and every tcgen05.mma would need its operands copied back (R2UR) - see umma_f16_elect.
  const uint32_t part_b = (uint32_t)(TM * 16 * 2) * (uint32_t)nks;  // bytes per staged K-slice of a candidate image
  const uint32_t tile_b = part_b * (uint32_t)nsplit;                // bytes per 128-row image
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);   // fixed offset: barrier addresses stay uniform values
  unsigned char* As = smem_raw + BAR_BYTES;                 // QH images (128*QH queries)
  unsigned char* Bs0 = As + QH * tile_b;                    // nstage slices
  uint64_t* full = bars;                  // [MAXST] producer -> MMA
  uint64_t* empty = bars + MAXST;         // [MAXST] MMA (commit) -> producer
  uint64_t* afull = bars + 2 * MAXST;     // [1]
  uint64_t* tfull = bars + 2 * MAXST + 1;    // [2 buffers][2 halves] MMA (commit) -> epilogue warps of that half
  uint64_t* tempty = bars + 2 * MAXST + 5;   // [2 buffers][2 halves] epilogue warps of that half -> its MMA issuer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAXST + 17);
  static_assert((2 * MAXST + 17) * 8 + 4 <= BAR_BYTES, "barrier block");

  // warp index through a shuffle: tells the compiler it is warp-uniform, so each role's branch is a converged region
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nstage; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], QH); }
    mbar_init(afull, 1);
    for (int b = 0; b < 4; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tmem_base != 0) __trap();  // a 512-column allocation can only start at lane 0, column 0 (the MMA warp relies on it)

  if (warp == 0) {
    // ---------------- producer ----------------
    if (lane == 0) {
      mbar_expect_tx(afull, QH * tile_b);
      for (int h = 0; h < QH; ++h)
        bulk_g2s(As + (size_t)h * tile_b,
                 reinterpret_cast<const unsigned char*>(Aimg) + (size_t)(qtile0 + QH * (int64_t)blockIdx.x + h) * tile_b, tile_b, afull);
      // visit order: n_est sample tiles (every est_stride-th) for the threshold estimate, then every tile
      int s = 0;
      uint32_t use_phase = 1;  // parity of the previous use of stage s (first lap: nothing to wait for)
      bool first_lap = true;
      for (int64_t v = 0; v < n_est + n_btiles; ++v) {
        const int64_t c = v < n_est ? v * est_stride : v - n_est;
        for (int p = 0; p < nsplit; ++p) {
          if (!first_lap) mbar_wait(&empty[s], use_phase);
          mbar_expect_tx(&full[s], part_b);
          bulk_g2s(Bs0 + (size_t)s * part_b, reinterpret_cast<const unsigned char*>(Bimg) + (size_t)c * tile_b + (size_t)p * part_b,
                   part_b, &full[s]);
          if (++s == nstage) { s = 0; use_phase ^= 1u; first_lap = false; }
        }
      }
    }
  } else if (warp == 1 || (warp == 10 && QH == 2)) {
    // ---------------- MMA issuers (whole warp, elected thread issues) ----------------
    // One issuer per query half, on different SM sub-partitions (warp 1 -> SMSP 1, warp 10 -> SMSP 2), each with
    // its own accumulator hand-off (tfull / tempty[2*buffer + half], one waiter per barrier): the epilogue warps of
    // half 0 start on their slab while the MMAs of half 1 are still running.  (Measured: a tcgen05.mma that finds
    // the tensor queue full stalls its sub-partition's dispatch, so the two epilogue warps living next to an issuer
    // fall behind the others; rotating whole visits over three issuer warps was tried and is not faster.)
    {
      const int h = warp == 1 ? 0 : 1;
      mbar_wait(afull, 0);
      // descriptor halves: lo = start address >> 4 | LBO field (bits 16..29); hi = SBO field | version; one k-step
      // (two 8-column K-chunks) advances the start address by 2*LBO bytes
      const uint64_t d0 = umma_desc(0);
      const uint32_t desc_hi = (uint32_t)(d0 >> 32);
      const uint32_t a_lo0 = (uint32_t)d0 | (((smem_u32(smem_raw) + BAR_BYTES) & 0x3FFFFu) >> 4);  // As
      const uint32_t b_lo0 = a_lo0 + (uint32_t)QH * mma_tile16;                       // Bs0 = As + QH images
      constexpr uint32_t KSTEP16 = (2u * LBO) >> 4;
      int s = 0;
      uint32_t ring_phase = 0;
      // mma_part16 / mma_tile16 / mma_visits repeat part_b >> 4, tile_b >> 4 and n_est + n_btiles as parameters of
      // their own, so that this warp's copies stay on the uniform datapath (the other warps use the vector ones)
      for (int c = 0; c < mma_visits; ++c) {  // c counts visits here (the operands come through the ring)
        const int b = c & 1;
        const int useb = c >> 1;
        for (int p = 0; p < nsplit; ++p) {
          const bool dbg = DBG && blockIdx.x == 0 && warp == 1 && lane == 0 && c >= DBG_C0 && c < DBG_C0 + DBG_N;
          if (dbg) g_knn_dbg[0 * 64 + (c - DBG_C0)] = clock64();
          mbar_wait(&full[s], ring_phase);
          if (dbg) g_knn_dbg[1 * 64 + (c - DBG_C0)] = clock64();
          if (p == 0 && useb > 0) mbar_wait(&tempty[2 * b + h], (uint32_t)((useb - 1) & 1));
          if (dbg) g_knn_dbg[2 * 64 + (c - DBG_C0)] = clock64();
          tc_fence_after();
          // the CTA owns all 512 TMEM columns, so its allocation starts at column 0 (checked after the alloc)
          const uint32_t tm = (uint32_t)(b * (128 * QH) + h * 128);
          uint32_t da = a_lo0 + (uint32_t)h * mma_tile16 + (uint32_t)(p * nks) * KSTEP16;
          uint32_t db = b_lo0 + (uint32_t)s * mma_part16;
          if (p == 0) umma_f16_elect<false>(tm, da, db, desc_hi); else umma_f16_elect<true>(tm, da, db, desc_hi);
#pragma unroll 4
          for (int j = 1; j < nks; ++j) {
            da += KSTEP16;
            db += KSTEP16;
            umma_f16_elect<true>(tm, da, db, desc_hi);
          }
          tc_commit_elect(&empty[s]);   // smem stage reusable once the MMAs of both halves have read it (count QH)
          if (++s == nstage) { s = 0; ring_phase ^= 1u; }
        }
        tc_commit_elect(&tfull[2 * b + h]);  // accumulators of (buffer b, half h) complete
        if (DBG && blockIdx.x == 0 && warp == 1 && lane == 0 && c >= DBG_C0 && c < DBG_C0 + DBG_N) g_knn_dbg[3 * 64 + (c - DBG_C0)] = clock64();
      }
      // teardown: the epilogue's release of the last visit means every TMEM read of this CTA has completed
      if (warp == 1) {
        if (mma_visits > 0) {
          const int bl = (mma_visits - 1) & 1;
          const uint32_t pl = (uint32_t)(((mma_visits - 1) >> 1) & 1);
          mbar_wait(&tempty[2 * bl], pl);
          if (QH == 2) mbar_wait(&tempty[2 * bl + 1], pl);
        }
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(0u), "r"(512u) : "memory");
      }
    }
  } else if (warp >= 2 && warp < 10 && ((warp - 2) >> 2) < QH) {
    // ---------------- epilogue: thread <-> query row ----------------
    const int e = warp - 2;                 // 0..7
    const int h = e >> 2;                   // query half
    const int lgrp = warp & 3;              // TMEM lane group this warp may access
    const int row = lgrp * 32 + lane;
    const int64_t ql = ((int64_t)blockIdx.x * QH + h) * TM + row;  // local query index
    const bool valid = ql < n_query;
    constexpr int LM = 8 * NG;
    float* sc = cand_score + (valid ? ql : 0) * LM;
    int32_t* id = cand_idx + (valid ? ql : 0) * LM;
    RowList<NG> L;
    L.tau = INFINITY;
    if (valid) {
#pragma unroll
      for (int i = 0; i < LM; ++i) id[i] = -1;
    }
    // threshold estimate: the EST_R largest chunk maxima over the sample tiles.  The EST_R-th of them is, in
    // expectation, the score of rank EST_R * est_stride among all candidates; the list starts "full" of sentinels
    // at that score, so the sweep only ever handles the ~100 candidates per row that beat it (instead of the
    // LM ln(n/LM) insertions of a cold start).  Any value is safe: the re-score certificate uses the final tau.
    float est[EST_R];
#pragma unroll
    for (int i = 0; i < EST_R; ++i) est[i] = -INFINITY;
    // software pipeline over the 32-column chunks of the accumulator slabs, two register buffers (va, vb), two
    // tcgen05.ld in flight: a slab costs two load round trips, and it is handed back to the MMA warp after the
    // second one - before its last two chunks are examined.
    uint32_t va[32], vb[32];
    const uint32_t lane_base = tmem_base + ((uint32_t)(lgrp * 32) << 16) + (uint32_t)(h * 128);
    mbar_wait(&tfull[h], 0);
    tc_fence_after();
    tmem_ld32_nowait(lane_base, va);
    tmem_ld32_nowait(lane_base + 32u, vb);
    const int64_t n_visit = n_est + n_btiles;
    const int dbg_mode = DBG ? g_knn_dbg_mode : 0;
    for (int64_t c = 0; c < n_visit; ++c) {
      const int b = (int)(c & 1);
      const uint32_t tbase = lane_base + (uint32_t)(b * (128 * QH));
      const uint32_t nbase = lane_base + (uint32_t)((b ^ 1) * (128 * QH));
      const bool estimating = c < n_est;
      const bool more = c + 1 < n_visit;
      const int32_t cbase = (int32_t)((estimating ? 0 : c - n_est) * TM);
      if (c == n_est) {
        const float t0 = valid ? est[EST_R - 1] : INFINITY;  // rows past n_query never accept anything
#pragma unroll
        for (int i = 0; i < LM; ++i) L.ls[i] = t0;
#pragma unroll
        for (int g = 0; g < NG; ++g) L.gm[g] = t0;
        L.tau = t0;
      }
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const bool dbg = DBG && blockIdx.x == 0 && warp == 2 && lane == 0 && c >= DBG_C0 && c < DBG_C0 + DBG_N;
        if (dbg && half == 0) g_knn_dbg[4 * 64 + (c - DBG_C0)] = clock64();
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");  // va = chunk 2*half, vb = chunk 2*half + 1
        if (half == 1) {
          // the whole row slab is in registers: hand the accumulator buffer back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[2 * b + h]);
          if (dbg) g_knn_dbg[5 * 64 + (c - DBG_C0)] = clock64();
          if (DBG && blockIdx.x == 0 && lane == 0 && c >= DBG_C0 && c < DBG_C0 + 16) g_knn_dbg2[(1 * 8 + e) * 16 + (c - DBG_C0)] = clock64();
        }
        if (dbg_mode == 0) { if (estimating) est_chunk(est, va); else scan_chunk(L, id, va, cbase + half * 64, n_points); }
        if (half == 0) {
          if (dbg_mode < 2) tmem_ld32_nowait(tbase + 64u, va);
        } else if (more) {
          mbar_wait(&tfull[2 * (b ^ 1) + h], (uint32_t)(((c + 1) >> 1) & 1));
          if (dbg) g_knn_dbg[6 * 64 + (c - DBG_C0)] = clock64();  // tfull of visit c+1 seen
          if (DBG && blockIdx.x == 0 && lane == 0 && c + 1 >= DBG_C0 && c + 1 < DBG_C0 + 16) g_knn_dbg2[(0 * 8 + e) * 16 + (c + 1 - DBG_C0)] = clock64();
          tc_fence_after();
          if (dbg_mode < 2) tmem_ld32_nowait(nbase, va);
        }
        if (dbg_mode == 0) { if (estimating) est_chunk(est, vb); else scan_chunk(L, id, vb, cbase + half * 64 + 32, n_points); }
        if (half == 0) {
          if (dbg_mode < 2) tmem_ld32_nowait(tbase + 96u, vb);
        } else if (more) {
          if (dbg_mode < 2) tmem_ld32_nowait(nbase + 32u, vb);
        }
        if (dbg && half == 1) g_knn_dbg[7 * 64 + (c - DBG_C0)] = clock64();  // visit c fully examined
        if (DBG && half == 1 && blockIdx.x == 0 && lane == 0 && c >= DBG_C0 && c < DBG_C0 + 16) g_knn_dbg2[(2 * 8 + e) * 16 + (c - DBG_C0)] = clock64();
      }
    }
    if (valid) {
#pragma unroll
      for (int i = 0; i < LM; ++i) sc[i] = L.ls[i];
    }
  }
  // no CTA-wide barrier down here: code after the role branches would make ptxas treat the MMA warp's region as
  // divergent (its operands then leave the uniform datapath, ~2x slower issue).  The MMA warp tears down by itself.
}

}  // namespace

// widest K axis the kernel can stage: one query image + three quarter-slices of a candidate image in 227 KB
bool knn_tc_supported(int d) { return 3 * d + 3 <= 512; }

bool knn_tc_shape(const sb2_ctx* ctx, int d, int terms, KnnTcShape* out) {
  // 256 queries per CTA with whole candidate images per ring stage when that fits; otherwise 128 queries per CTA
  // and the candidate K axis staged in 1, 2 or 4 slices.  The ring takes whatever shared memory is left (<= MAXST).
  const size_t cap = ctx->prop.sharedMemPerBlockOptin;
  const int cand[4][2] = {{2, 1}, {1, 1}, {1, 2}, {1, 4}};
  for (int t = 0; t < 4; ++t) {
    const int unit = 16 * cand[t][1];
    const int kp = ((terms * d + 3 + unit - 1) / unit) * unit;
    const size_t tb = (size_t)TM * kp * 2, part = tb / cand[t][1];
    if (cand[t][0] * tb + 3 * part + BAR_BYTES > cap) continue;
    int nst = (int)((cap - BAR_BYTES - cand[t][0] * tb) / part);
    if (nst > MAXST) nst = MAXST;
    out->qh = cand[t][0]; out->nsplit = cand[t][1]; out->nstage = nst; out->kpad = kp; out->terms = terms;
    out->smem = cand[t][0] * tb + (size_t)nst * part + BAR_BYTES;
    return true;
  }
  return false;
}

size_t knn_tc_image_halves(const KnnTcShape& sh, int64_t n_rows) {
  const int64_t n_tiles = ceil_div64(n_rows, TM);
  return (size_t)(n_tiles + (n_tiles & 1) + 2) * TM * sh.kpad;  // A images may be consumed in pairs
}

int32_t knn_tc_build_images(sb2_ctx* ctx, const KnnTcShape& sh, const float* d_x, int64_t n_rows, int d,
                            const unsigned int* d_maxnorm_bits, const int32_t* d_gather, int64_t gather_base,
                            __half* Aimg, __half* Bimg, float* d_inv_s2, float* d_dnorm, unsigned int* d_dmax_bits) {
  const int64_t n_tiles = ceil_div64(n_rows, TM);
  const int64_t n_tiles_alloc = n_tiles + (n_tiles & 1) + 2;
  knn_tc_prep_kernel<<<(unsigned)n_tiles_alloc, 256, 0, ctx->stream>>>(d_x, n_rows, d, sh.kpad, sh.terms, d_maxnorm_bits, d_gather,
                                                                        gather_base, Aimg, Bimg, d_inv_s2, d_dnorm, d_dmax_bits);
  SB2_LAUNCH_CHECK(ctx);
  return SB2_OK;
}

namespace {
template <int NG, int QH>
cudaError_t launch_tc(unsigned grid, const KnnTcShape& sh, cudaStream_t st, const __half* A, const __half* B, int64_t n_tiles,
                      int64_t n_est, int64_t est_stride, int64_t qtile0, int64_t n_query, int32_t n_points, float* cs,
                      int32_t* ci) {
  cudaError_t e = cudaFuncSetAttribute(knn_pass1_tc_kernel<NG, QH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem);
  if (e != cudaSuccess) return e;
  knn_pass1_tc_kernel<NG, QH><<<grid, TC_THREADS, sh.smem, st>>>(A, B, sh.kpad / 16 / sh.nsplit, sh.nsplit, sh.nstage, (uint32_t)(TM * sh.kpad * 2 / sh.nsplit) >> 4,
                                                                 (uint32_t)(TM * sh.kpad * 2) >> 4, (int)(n_tiles + n_est), n_tiles, n_est, est_stride, qtile0,
                                                                 n_query, n_points, cs, ci);
  return cudaSuccess;
}
}  // namespace

int32_t knn_tc_sweep(sb2_ctx* ctx, const KnnTcShape& sh, const __half* Aimg, int64_t a_tile0, const __half* Bimg,
                     int64_t n_points, int64_t n_query, int list_m, float* cand_score, int32_t* cand_idx, double* issued_flops,
                     bool estimate) {
  cudaStream_t st = ctx->stream;
  const int64_t n_tiles = ceil_div64(n_points, TM);
  const int64_t q_ctas = ceil_div64(n_query, (int64_t)sh.qh * TM);
  // threshold estimate: EST_R chunk maxima over every est_stride-th tile put the starting tau at about rank
  // EST_R * est_stride; aim at ~3 * list_m so that the list still fills (and ends at its usual rank) for nearly all rows
  int64_t est_stride = (3 * list_m) / EST_R, n_est = n_tiles / est_stride;
  const char* est_env = getenv("SB2_KNN_EST");
  if (est_env) est_stride = atoi(est_env) > 0 ? atoi(est_env) : est_stride, n_est = atoi(est_env) > 0 ? n_tiles / est_stride : 0;
  if (n_tiles < 64 * est_stride || !estimate) n_est = 0;  // small problems / re-sweeps: cold start
  if (issued_flops) *issued_flops += 2.0 * (double)(q_ctas * sh.qh * TM) * (double)((n_tiles + n_est) * TM) * (double)sh.kpad;
  SB2_CHECK_ARG(list_m == 32 || list_m == 64, "list_m must be 32 or 64");
  cudaError_t le;
  const unsigned grid = (unsigned)q_ctas;
  const int32_t np = (int32_t)n_points;
  const char* dbg_env = getenv("SB2_KNN_DBG");
  if (dbg_env && list_m == 32 && sh.qh == 2) {
    const int mode = atoi(dbg_env) - 1;  // SB2_KNN_DBG=1: stamps only; 2: no scans; 3: no TMEM loads either
    SB2_CUDA(cudaMemcpyToSymbol(g_knn_dbg_mode, &mode, sizeof(int)));
    cudaEvent_t d0, d1;
    cudaEventCreate(&d0); cudaEventCreate(&d1);
    cudaEventRecord(d0, st);
    le = cudaFuncSetAttribute(knn_pass1_tc_kernel<4, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem);
    knn_pass1_tc_kernel<4, 2, true><<<grid, TC_THREADS, sh.smem, st>>>(Aimg, Bimg, sh.kpad / 16 / sh.nsplit, sh.nsplit, sh.nstage, (uint32_t)(TM * sh.kpad * 2 / sh.nsplit) >> 4,
                                                                      (uint32_t)(TM * sh.kpad * 2) >> 4, (int)(n_tiles + n_est), n_tiles, n_est, est_stride,
                                                                      a_tile0, n_query, np, cand_score, cand_idx);
    long long h[8 * 64];
    cudaEventRecord(d1, st);
    SB2_CUDA(cudaStreamSynchronize(st));
    float dms = 0;
    cudaEventElapsedTime(&dms, d0, d1);
    fprintf(stderr, "[dbg mode %d] kernel %.3f ms, grid %u, visits %lld\n", mode, dms, grid, (long long)(n_tiles + n_est));
    SB2_CUDA(cudaMemcpyFromSymbol(h, g_knn_dbg, sizeof(h)));
    fprintf(stderr, "visit  mma:wait_full  wait_tempty  issue   | epi: tfull_seen->consume  release  next_tfull_seen  done   (cycles, relative to visit's mma start)\n");
    long long h2[3 * 8 * 16];
    SB2_CUDA(cudaMemcpyFromSymbol(h2, g_knn_dbg2, sizeof(h2)));
    for (int i = 4; i < 8; ++i) {
      const long long t0 = h[0 * 64 + i];
      fprintf(stderr, "visit %d (mma start %lld, issued +%lld): per epilogue warp  tfull_seen / release / done  rel. to mma start\n", DBG_C0 + i,
              t0 - h[0 * 64 + 2], h[3 * 64 + i] - t0);
      for (int w = 0; w < 8; ++w)
        fprintf(stderr, "   warp %d: %6lld %6lld %6lld\n", w + 2, h2[(0 * 8 + w) * 16 + i] - t0, h2[(1 * 8 + w) * 16 + i] - t0, h2[(2 * 8 + w) * 16 + i] - t0);
    }
    for (int i = 2; i < 6; ++i) {
      const long long t0 = h[0 * 64 + i];
      fprintf(stderr, "%4d  start=%8lld  full+%5lld tempty+%5lld issued+%5lld | consume+%6lld release+%6lld nexttfull+%6lld done+%6lld\n", DBG_C0 + i,
              t0 - h[0 * 64 + 2], h[1 * 64 + i] - t0, h[2 * 64 + i] - t0, h[3 * 64 + i] - t0, h[4 * 64 + i] - t0, h[5 * 64 + i] - t0,
              h[6 * 64 + i] - t0, h[7 * 64 + i] - t0);
    }
  } else if (list_m == 32 && sh.qh == 2) le = launch_tc<4, 2>(grid, sh, st, Aimg, Bimg, n_tiles, n_est, est_stride, a_tile0, n_query, np, cand_score, cand_idx);
  else if (list_m == 32) le = launch_tc<4, 1>(grid, sh, st, Aimg, Bimg, n_tiles, n_est, est_stride, a_tile0, n_query, np, cand_score, cand_idx);
  else if (sh.qh == 2) le = launch_tc<8, 2>(grid, sh, st, Aimg, Bimg, n_tiles, n_est, est_stride, a_tile0, n_query, np, cand_score, cand_idx);
  else le = launch_tc<8, 1>(grid, sh, st, Aimg, Bimg, n_tiles, n_est, est_stride, a_tile0, n_query, np, cand_score, cand_idx);
  SB2_CUDA(le);
  SB2_LAUNCH_CHECK(ctx);
  return SB2_OK;
}

void knn_tc_error_coefs(const KnnTcShape& sh, double* c_q, double* c_n) {
  // |s_computed - s_true| <= c_q |q| R + c_n R^2 / 2   (R = largest norm in the data set, everything after scaling)
  //   fp32 accumulation in the tensor pipe: kpad/16 accumulator updates of one ulp each plus the alignment loss
  //   inside a 16-product group, bounded by 1.6 * kpad * 2^-24 of sum |a_i b_i| <= |q| R + R^2 / 2;
  //   three-way fp16 split of the norm: 2^-33 R^2/2; sub-normal halves: < 2^-27 R^2/2 since R >= 100 after scaling
  //   terms = 3: operand split 3 * 2^-24 |q| R, dropped lo*lo term 2^-22 * 2^-2 |q| R
  //   terms = 1: |q.c - q_hi.c_hi| <= |q - q_hi| |c| + |q_hi| |c - c_hi|: the re-score kernel adds this term from
  //              the measured residual norms (dnorm / dmax of knn_tc_build_images), not from the 2^-11 worst case
  const double u24 = 5.9604644775390625e-08;
  const double acc = 1.6 * sh.kpad * u24;
  *c_n = acc + 8.0 * u24;
  *c_q = acc + (sh.terms == 3 ? 8.0 * u24 : 0.0);
}
