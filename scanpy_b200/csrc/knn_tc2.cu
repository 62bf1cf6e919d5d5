// knn_tc2.cu — tensor-core sweep of the exact kNN, second generation (sm_100a: tcgen05.mma N = 256 + TMEM + bulk-copy ring).
//
// Same contract as the first-generation sweep (knn_tc.cu): per query keep the proposals whose score
//     s(q,c) = q.c - |c|^2/2        (d^2 = |q|^2 - 2 s)
// is largest, plus a rigorous bound on the rounding error of s, so that knn_rescore_kernel can certify the exact fp64
// top-k.  Operand formats (terms = 1: fp16, K = d+3; terms = 3: split fp16, K = 3d+3) are unchanged.
//
// What round 2 measured on the first generation (profiles/README.md, scripts/ubench/mma_smem_contention.cu,
// alu_max_rate.cu, the SB2_KNN_DBG / SB2_KNN2_NOSCAN cycle stamps):
//   * a 128x128x16 tcgen05.mma with both operands in shared memory reads 8 KB per 64 tensor cycles = 128 B/clk, ALL of
//     the shared-memory bandwidth; next to the bulk-copy ring the pipe delivered one MMA per ~105 cycles (850 cycles per
//     256x128 tile even with an EMPTY epilogue, 512 ideal).  128x256x16 MMAs (12 KB per 128 cycles) hold 131-138 cycles
//     under the same traffic: with N = 256 and scanners that only load and release, this kernel runs at 521-546 cycles
//     per 128x256 tile (tensor floor 512; power-capped at ~1000 W: 1450-1630 TFLOP/s issued);
//   * the first generation's epilogue (two warps per scheduler, one dependent max tree each) had too little instruction
//     level parallelism: ~1200 cycles per tile.
// Hence this shape (one CTA per SM, 18 warps):
//   CTA = 128 queries (UMMA M) x all candidate tiles of 256 points (UMMA N = 256), two 256-column accumulator buffers.
//     warps 0-15  epilogue, FOUR warps per TMEM lane quadrant: warp (quadrant q, column quarter cq) owns rows
//                 32q..32q+31 x columns 64cq..64cq+63 of every tile: tcgen05.ld.32x32b.x32 of its first chunk, examine,
//                 load the second, release the buffer (16th arrival), examine.  Each (row, column quarter) keeps its
//                 OWN register-resident proposal sub-list (16 or 32 scores in groups of 8 with group minima; ids
//                 write-only to global memory): no cross-warp traffic in the loop.  A point outside all four sub-lists
//                 of a row scored at most max(tau_0..tau_3), which is what the re-score certificate uses.
//     warp 16     producer: 1-D cp.async.bulk of pre-arranged candidate images (or K-slices of them) into an mbarrier
//                 ring that takes all the shared memory left
//     warp 17     MMA issuer (whole-warp loop, elect.sync inside the asm block, operands on the uniform datapath);
//                 tcgen05.commit frees the ring stage and publishes the accumulator buffer; owns TMEM alloc/dealloc
//   Service warps carry the highest warp ids (the sub-partition arbiter prefers them among eligible warps).
//   Starting threshold: the first n/16 visits sweep every 16th tile and keep the 6 largest chunk maxima per (row,
//   quarter); the four quarters of a row are merged through shared memory once (named barrier over the 16 epilogue
//   warps) and the 6th largest of the 24 seeds all four sub-lists (rank ~96 of the row, as before).
//   Sweep front: CTAs walk the candidate tiles cyclically, starting where the running CTAs currently are (g_knn2_front).
// Where the time goes now (1.3M x 50, k = 15): 1000-1100 cycles per tile against the 512 of the tensor pipe.  The
// examine phases cost each warp ~600 cycles per tile (~290 of them the ALU pipe's floor for the max trees: FMNMX3 issues
// every 2 cycles per scheduler, FMNMX every cycle), and ~6 % of the chunk scans take the insertion path (~150 dependent
// instructions for one lane); since the release of an accumulator buffer needs all 16 warps, whichever warp is inserting
// gates the hand-off of nearly every tile.  Two decouplings were built, measured and dropped (git history):
//   - scanners that push (row, id, score) into a shared-memory queue drained by inserter warps holding one list per row:
//     exact, but the claim / release-acquire traffic cost more than it saved (1330 cycles per tile);
//   - both chunks loaded before anything is examined and the sub-lists kept in shared memory (early release): the
//     insertion walk over shared memory and the 96-register cap (18 warps) made each examine phase slower (1180-1240).
#include <cuda_fp16.h>
#include <float.h>

#include "common.cuh"
#include "knn_internal.cuh"
#include "tc_ptx.cuh"

namespace {
using namespace tcptx;

constexpr int QM = 128;            // queries per CTA (UMMA M, TMEM lanes)
constexpr int CN = 256;            // candidates per tile (UMMA N, TMEM columns per accumulator buffer)
constexpr int MAXST = 8;           // upper bound of the ring depth
constexpr int BAR_BYTES = 1024;    // mbarriers + TMEM slot IN FRONT of the operand images (fixed addresses)
constexpr int NEPI = 16;           // epilogue warps
constexpr int NQUART = 4;          // column quarters (epilogue warps per TMEM lane quadrant)
constexpr int THREADS = (2 + NEPI) * 32;
// Service warps carry the HIGHEST warp ids: the sub-partition arbiter prefers the highest warp id among eligible warps
// (B300_MICROARCH: "highest-wid-first"), so the producer and the MMA issuer are never starved by the four epilogue
// warps they share a scheduler with.
constexpr int W_PROD = NEPI, W_MMA = NEPI + 1;
constexpr int EST_R = 6;
constexpr int EST_BYTES = QM * NQUART * EST_R * 4;   // exchange area of the threshold estimate
constexpr uint32_t SBO = 128;                  // bytes between 8-row groups (core matrices contiguous)
constexpr uint32_t LBO_A = QM * 16;            // bytes between K-chunks of a 128-row image  [kc][row-group][8][8]
constexpr uint32_t LBO_B = CN * 16;            // ... of a 256-row image
constexpr uint32_t IDESC = umma_idesc_f16_m128(CN);

// scale: power of two s with s*R in [100, 200]  ->  fp16 range is safe for coordinates and for s^2 R^2/2
__device__ __forceinline__ float tc_scale_from_maxnorm(unsigned int maxnorm_bits) {
  const float R = sqrtf(__uint_as_float(maxnorm_bits));
  if (!(R > 0.0f) || !isfinite(R)) return 1.0f;
  int e = (int)floorf(log2f(200.0f / R));
  e = max(-60, min(60, e));
  return exp2f((float)e);
}

// X[n,d] -> operand images of `tile_rows` points each (128: query images, 256: candidate images), stored exactly as
// the UMMA "no-swizzle, K-major" shared-memory layout wants them (8x8 fp16 core matrices, K-chunk-major), so a tile
// (or a K-slice of it) is staged by ONE 1-D bulk copy.
//   is_a = 1: A = [hi | hi | lo | 1 1 1]  (terms = 3)   or  [hi | 1 1 1]      (terms = 1)
//   is_a = 0: B = [hi | lo | hi | h0 h1 h2]             or  [hi | h0 h1 h2]   with h0+h1+h2 = -|x|^2 s^2 / 2
// gather != nullptr: image row r of tile t is point gather_base + gather[t*tile_rows + r]; rows past n are zero rows
// (A) or score -60000 candidates (B).  dnorm[p] (optional) = |x_p - fp16(x_p)| in the units of X, rounded up;
// *dmax_bits = its maximum (float bits).
__global__ void __launch_bounds__(256)
knn_tc2_prep_kernel(const float* __restrict__ X, int64_t n, int d, int kpad, int terms, int tile_rows, int is_a,
                    const unsigned int* __restrict__ maxnorm_bits, const int32_t* __restrict__ gather, int64_t gather_base,
                    __half* __restrict__ img, float* __restrict__ inv_s2, float* __restrict__ dnorm,
                    unsigned int* __restrict__ dmax_bits) {
  const int64_t t = blockIdx.x;
  const float s = tc_scale_from_maxnorm(*maxnorm_bits);
  if (t == 0 && threadIdx.x == 0 && inv_s2) *inv_s2 = 1.0f / (s * s);
  __shared__ __half hn3[CN][3];
  const int kd = terms * d;  // coordinates on the K axis before the three norm slots
  if ((int)threadIdx.x < tile_rows) {
    const int64_t p = t * tile_rows + threadIdx.x;
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
  __half* out = img + (size_t)t * tile_rows * kpad;
  const size_t lbo = (size_t)tile_rows * 16;
  for (int i = threadIdx.x; i < tile_rows * nkc; i += blockDim.x) {
    const int kc = i / tile_rows, r = i % tile_rows;
    const int64_t p = t * tile_rows + r;
    const int64_t ps = (gather && p < n) ? gather_base + gather[p] : p;
    __align__(16) __half o8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = kc * 8 + j;
      __half v = __float2half_rn(0.0f);
      if (p < n) {
        if (e < kd) {
          const int seg = e / d, k = e - seg * d;
          const float xs = X[ps * d + k] * s;
          const __half hi = __float2half_rn(xs);
          const __half lo = __float2half_rn(xs - __half2float(hi));
          v = is_a ? (seg == 2 ? lo : hi) : (seg == 1 ? lo : hi);   // A = [hi | hi | lo], B = [hi | lo | hi]
        } else if (e < kd + 3) {
          v = is_a ? __float2half_rn(1.0f) : hn3[r][e - kd];
        }
      } else if (e == kd && !is_a) {
        v = hn3[r][0];  // padding candidates: score -60000 (padding queries are all-zero rows)
      }
      o8[j] = v;
    }
    const size_t off = ((size_t)kc * lbo + (size_t)(r >> 3) * SBO + (size_t)(r & 7) * 16) / 2;  // in halves
    *reinterpret_cast<uint4*>(out + off) = *reinterpret_cast<const uint4*>(o8);
  }
}

// ------------------------------------------------------------------------------------------------
// Per-(row, column quarter) proposal sub-list: 8*NG scores in REGISTERS (groups of 8 with a running minimum per
// group), ids in global memory (write-only).  tau = min of the sub-list = the score a candidate must beat.  All
// indexing is static (macro-expanded), so nothing spills and an insertion never waits on memory.
template <int NG>
struct SubList {
  float ls[8 * NG];
  float gm[NG];
  float tau;
};
__device__ __forceinline__ float min8(const float* x) {
  return fminf(fminf(fminf(x[0], x[1]), fminf(x[2], x[3])), fminf(fminf(x[4], x[5]), fminf(x[6], x[7])));
}
#define SB2_GROUP_INSERT2(G)                                                  \
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
__device__ __forceinline__ void list_insert(SubList<NG>& L, int32_t* __restrict__ id, float v, int32_t cand) {
  static_assert(NG == 2 || NG == 4, "sub-lists hold 16 or 32 proposals");
  int pos = 0;
  if (L.gm[0] == L.tau) SB2_GROUP_INSERT2(0)
  else if (NG == 2 || L.gm[1] == L.tau) SB2_GROUP_INSERT2(1)
  else if (L.gm[2 % NG] == L.tau) SB2_GROUP_INSERT2(2 % NG)
  else SB2_GROUP_INSERT2(3 % NG)
  id[pos] = cand;
  float t = L.gm[0];
#pragma unroll
  for (int g = 1; g < NG; ++g) t = fminf(t, L.gm[g]);
  L.tau = t;
}
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
// examine one 32-column chunk of the row (values already in registers).  Fast path: a two-level 3-input max tree (four
// independent quarter maxima, so the FMNMX3 latencies overlap) and one compare.  Rare path: visit the quarters whose
// maximum beats tau, copy the quarter's eight values aside and pull its maxima out one by one.  The quarter loop is
// NOT unrolled: one copy of the extraction and insertion code per call site keeps the kernel in the instruction caches.
template <int NG>
__device__ __forceinline__ void scan_chunk(SubList<NG>& L, int32_t* __restrict__ id, const uint32_t (&v)[32], int32_t cand0,
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
// estimate phase: keep the EST_R largest values seen (sorted descending), branch-free
template <int LEN>
__device__ __forceinline__ void est_push(float (&est)[LEN], float cm) {
#pragma unroll
  for (int i = 0; i < EST_R; ++i) {
    const float hi = fmaxf(est[i], cm);
    cm = fminf(est[i], cm);
    est[i] = hi;
  }
}
template <int LEN>
__device__ __forceinline__ void est_chunk(float (&est)[LEN], const uint32_t (&v)[32]) {
  est_push(est, fmaxf(fmax3(max8u(&v[0]), max8u(&v[8]), max8u(&v[16])), max8u(&v[24])));
}

// nks = k-steps (16 columns of the K axis) per staged slice, nsplit = slices per candidate image.  Everything the MMA
// warp derives its operands from is a kernel parameter combined by add / multiply only (a division would move the
// values to the vector datapath and every tcgen05.mma would need its operands copied back through R2UR).
// NOSCAN (timing experiments only, SB2_KNN2_NOSCAN=1): the epilogue loads and releases but does not examine.
// SB2_KNN2_STAMP=1: CTA 0, warp 2 records clock64 / globaltimer around its visit loop (cycles per visit and the SM clock
// actually delivered under this kernel's load); read back by knn_tc2_sweep and printed to stderr
// Sweep front: CTAs walk the candidate tiles cyclically, each starting where the CTAs already running currently are
// (the order of the candidates is irrelevant to the result).  All resident CTAs then stream the same few tiles at the
// same time and the candidate images (166 MB at 1.3M x 50, more than the L2) are read from HBM about once per wave of
// CTAs instead of once per CTA.  "Where the running CTAs are" is the MEAN position of the pack, (sweep visits so far) /
// (resident CTAs): the first version started a new CTA at the position last published by ANY running CTA, so the
// spread of the pack was inherited and widened from wave to wave (a random walk over 68 waves: ncu showed 114 GB of
// DRAM reads per launch, L2 hit rate 77 %, profiles/r2_ncu_metrics_knn_sweep2.txt); re-centring every new CTA on the
// mean bounds the spread by the drift of a single lap.
__device__ unsigned long long g_knn2_front;   // sweep-proper visits of all CTAs of this launch, in units of 16 tiles
__device__ unsigned long long g_knn2_stamp[4];
__device__ unsigned long long g_knn2_phase[2][8];
__device__ __forceinline__ unsigned long long gtimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
template <int NG, int MODE>
__global__ void __launch_bounds__(THREADS, 1)
knn_sweep2_kernel(const __half* __restrict__ Aimg, const __half* __restrict__ Bimg, int nks, int nsplit, int nstage,
                  uint32_t mma_part16, uint32_t mma_atile16, int mma_visits, int64_t n_btiles, int64_t n_est,
                  int64_t est_stride, int64_t qtile0, int64_t n_query, int32_t n_points, float* __restrict__ cand_score,
                  int32_t* __restrict__ cand_idx) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t part_b = (uint32_t)(CN * 16 * 2) * (uint32_t)nks;      // bytes per staged K-slice of a candidate image
  const uint32_t tile_b = part_b * (uint32_t)nsplit;                    // bytes per 256-row candidate image
  const uint32_t tile_a = tile_b >> 1;                                  // bytes per 128-row query image
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);   // fixed offset: barrier addresses stay uniform values
  unsigned char* As = smem_raw + BAR_BYTES;
  unsigned char* Bs0 = As + tile_a;
  uint64_t* full = bars;                  // [MAXST] producer -> MMA
  uint64_t* empty = bars + MAXST;         // [MAXST] MMA (commit) -> producer
  uint64_t* afull = bars + 2 * MAXST;     // [1]
  uint64_t* tfull = bars + 2 * MAXST + 1;    // [2] MMA (commit) -> epilogue
  uint64_t* tempty = bars + 2 * MAXST + 3;   // [2] epilogue (16 warps) -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAXST + 5);
  int32_t* start_tile = reinterpret_cast<int32_t*>(bars + 2 * MAXST + 6);
  static_assert((2 * MAXST + 6) * 8 + 4 <= BAR_BYTES, "barrier block");
  float* est_x = reinterpret_cast<float*>(Bs0 + (size_t)nstage * part_b);  // [QM][NQUART][EST_R], present iff n_est > 0

  // warp index through a shuffle: tells the compiler it is warp-uniform, so each role's branch is a converged region
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nstage; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(afull, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], NEPI); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const unsigned long long f = *reinterpret_cast<volatile unsigned long long*>(&g_knn2_front);
    unsigned int nsm;
    asm volatile("mov.u32 %0, %%nsmid;" : "=r"(nsm));
    const unsigned long long resident = min((unsigned long long)gridDim.x, (unsigned long long)max(nsm, 1u));
    // mean tile of the pack now, plus the n_est visits this CTA spends on its threshold estimate before it joins
    *start_tile = (int32_t)((f * 16ull / resident + (unsigned long long)n_est) % (unsigned long long)n_btiles);
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tmem_base != 0) __trap();  // a 512-column allocation can only start at lane 0, column 0 (the MMA warp relies on it)

  if (warp == W_PROD) {
    // ---------------- producer ----------------
    if (lane == 0) {
      mbar_expect_tx(afull, tile_a);
      bulk_g2s(As, reinterpret_cast<const unsigned char*>(Aimg) + (size_t)(qtile0 + (int64_t)blockIdx.x) * tile_a, tile_a, afull);
      // visit order: n_est sample tiles (every est_stride-th) for the threshold estimate, then every tile
      int s = 0;
      uint32_t use_phase = 1;  // parity of the previous use of stage s (first lap: nothing to wait for)
      bool first_lap = true;
      int64_t cs = *start_tile;   // cyclic walk of the sweep proper
      for (int64_t v = 0; v < n_est + n_btiles; ++v) {
        int64_t c = v * est_stride;
        if (v >= n_est) { c = cs; if (++cs == n_btiles) cs = 0; }
        for (int p = 0; p < nsplit; ++p) {
          if (!first_lap) mbar_wait(&empty[s], use_phase);
          mbar_expect_tx(&full[s], part_b);
          bulk_g2s(Bs0 + (size_t)s * part_b, reinterpret_cast<const unsigned char*>(Bimg) + (size_t)c * tile_b + (size_t)p * part_b,
                   part_b, &full[s]);
          if (++s == nstage) { s = 0; use_phase ^= 1u; first_lap = false; }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ---------------- MMA issuer (whole warp, elected thread issues) ----------------
    mbar_wait(afull, 0);
    // descriptor halves: lo = start address >> 4 | LBO field (bits 16..29); hi = SBO field | version; one k-step (two
    // 8-column K-chunks) advances the start address by 2*LBO bytes
    const uint64_t da0 = umma_desc(0, LBO_A, SBO), db0 = umma_desc(0, LBO_B, SBO);
    const uint32_t desc_hi = (uint32_t)(da0 >> 32);  // identical for A and B (same SBO)
    const uint32_t as16 = ((smem_u32(smem_raw) + BAR_BYTES) & 0x3FFFFu) >> 4;
    const uint32_t a_lo0 = (uint32_t)da0 | as16;
    const uint32_t b_lo0 = ((uint32_t)db0 | as16) + mma_atile16;   // Bs0 = As + one query image
    constexpr uint32_t KSTEP_A16 = (2u * LBO_A) >> 4, KSTEP_B16 = (2u * LBO_B) >> 4;
    int s = 0;
    uint32_t ring_phase = 0;
    // mma_part16 / mma_atile16 / mma_visits repeat part_b >> 4, tile_a >> 4 and n_est + n_btiles as parameters of their
    // own, so that this warp's copies stay on the uniform datapath (the other warps use the vector ones)
    constexpr bool PHASES_M = MODE == 2;
    long long mph[3] = {0, 0, 0}, mtp = 0;
    if (PHASES_M) mtp = clock64();
#define SB2_MPH(i) if (PHASES_M) { const long long tn = clock64(); mph[i] += tn - mtp; mtp = tn; }
    for (int c = 0; c < mma_visits; ++c) {
      const int b = c & 1;
      const int useb = c >> 1;
      for (int p = 0; p < nsplit; ++p) {
        mbar_wait(&full[s], ring_phase);
        SB2_MPH(0)
        if (p == 0 && useb > 0) mbar_wait(&tempty[b], (uint32_t)((useb - 1) & 1));
        SB2_MPH(1)
        tc_fence_after();
        const uint32_t tm = (uint32_t)(b * CN);  // the CTA owns all 512 TMEM columns: allocation starts at column 0
        uint32_t da = a_lo0 + (uint32_t)(p * nks) * KSTEP_A16;
        uint32_t db = b_lo0 + (uint32_t)s * mma_part16;
        if (p == 0) umma_f16_elect<false>(tm, da, db, desc_hi, IDESC); else umma_f16_elect<true>(tm, da, db, desc_hi, IDESC);
#pragma unroll 4
        for (int j = 1; j < nks; ++j) {
          da += KSTEP_A16;
          db += KSTEP_B16;
          umma_f16_elect<true>(tm, da, db, desc_hi, IDESC);
        }
        tc_commit_elect(&empty[s]);   // ring stage reusable once these MMAs have read it
        if (++s == nstage) { s = 0; ring_phase ^= 1u; }
      }
      tc_commit_elect(&tfull[b]);     // accumulator buffer b complete
      SB2_MPH(2)
    }
#undef SB2_MPH
    if (PHASES_M && blockIdx.x == gridDim.x / 2 && lane == 0) {
      g_knn2_phase[0][6] = (unsigned long long)mph[0]; g_knn2_phase[0][7] = (unsigned long long)mph[1]; g_knn2_phase[1][6] = (unsigned long long)mph[2];
    }
    // teardown: every epilogue warp walks the visits in order, so the release of the last visit by all of them means
    // every TMEM read of this CTA has completed
    if (mma_visits > 0) {
      const int bl = (mma_visits - 1) & 1;
      mbar_wait(&tempty[bl], (uint32_t)(((mma_visits - 1) >> 1) & 1));
    }
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(0u), "r"(512u) : "memory");
  } else {
    // ---------------- epilogue: thread <-> (query row, column quarter) ----------------
    const int lgrp = warp & 3;              // TMEM lane quadrant this warp may access
    const int cq = warp >> 2;               // column quarter 0..3 (the four warps of a quadrant: w, w+4, w+8, w+12)
    const int row = lgrp * 32 + lane;
    const int64_t ql = (int64_t)blockIdx.x * QM + row;  // local query index
    const bool valid = ql < n_query;
    constexpr int LS = 8 * NG, LM = NQUART * LS;
    float* sc = cand_score + (valid ? ql : 0) * LM + cq * LS;
    int32_t* id = cand_idx + (valid ? ql : 0) * LM + cq * LS;
    SubList<NG> L;
    L.tau = INFINITY;
    if (valid) {
#pragma unroll
      for (int i = 0; i < LS; ++i) id[i] = -1;
    }
    // During the estimate phase the EST_R running chunk maxima live in L.ls[0..EST_R) (the sub-list proper does not exist
    // yet): one register array serves both phases.
    static_assert(EST_R <= LS, "estimate aliases the list registers");
#pragma unroll
    for (int i = 0; i < LS; ++i) L.ls[i] = -INFINITY;
    uint32_t v[32];
    const uint32_t lane_base = tmem_base + ((uint32_t)(lgrp * 32) << 16) + (uint32_t)(cq * 64);
    const int n_visit = mma_visits, n_est_i = (int)n_est;
    int32_t ctile = *start_tile;                 // candidate tile of the current sweep visit (cyclic walk)
    int32_t cbase = ctile * CN + cq * 64;
    const int32_t n_bt = (int32_t)n_btiles;
    const bool stamp = blockIdx.x == gridDim.x / 2 && warp == 2 && lane == 0;
    unsigned long long st_c = 0, st_t = 0;
    if (stamp) { st_c = clock64(); st_t = gtimer_ns(); }
    constexpr bool NOSCAN = MODE == 1;
    constexpr bool PHASES = MODE == 2;
    long long ph[6] = {0, 0, 0, 0, 0, 0}, tp = 0;
    const bool phw = PHASES && blockIdx.x == gridDim.x / 2 && (warp == 2 || warp == 5) && lane == 0  /* SMSP 2 (no service warp) and SMSP 1 (MMA issuer) */;
    if (PHASES) tp = clock64();
#define SB2_PH(i) if (PHASES) { const long long tn = clock64(); ph[i] += tn - tp; tp = tn; }
    for (int c = 0; c < n_visit; ++c) {
      const int b = c & 1;
      mbar_wait(&tfull[b], (uint32_t)((c >> 1) & 1));
      tc_fence_after();
      SB2_PH(0)
      const uint32_t taddr = lane_base + (uint32_t)(b * CN);
      tmem_ld32_nowait(taddr, v);
      tmem_wait_ld();
      SB2_PH(1)
      if (c == n_est_i) {
        // start of the sweep proper: seed the sub-list.  With an estimate phase the four column quarters of a row are
        // merged (the 6th largest of their 24 chunk maxima sits at about rank 6 * est_stride = 96 of the row);
        // without one the lists start cold (-inf).  Any value is safe: the certificate uses the final tau.
        if (n_est_i > 0) {
#pragma unroll
          for (int i = 0; i < EST_R; ++i) est_x[(row * NQUART + cq) * EST_R + i] = L.ls[i];
          named_bar_sync(1, NEPI * 32);
#pragma unroll
          for (int o = 1; o < NQUART; ++o) {
            const int oq = (cq + o) & (NQUART - 1);
#pragma unroll
            for (int i = 0; i < EST_R; ++i) est_push(L.ls, est_x[(row * NQUART + oq) * EST_R + i]);
          }
        }
        const float t0 = valid ? L.ls[EST_R - 1] : INFINITY;  // rows past n_query never accept anything
#pragma unroll
        for (int i = 0; i < LS; ++i) L.ls[i] = t0;
#pragma unroll
        for (int g = 0; g < NG; ++g) L.gm[g] = t0;
        L.tau = t0;
      }
      const bool estimating = c < n_est_i;
      if (!NOSCAN) { if (estimating) est_chunk(L.ls, v); else scan_chunk(L, id, v, cbase, n_points); }
      SB2_PH(2)
      tmem_ld32_nowait(taddr + 32u, v);
      tmem_wait_ld();
      SB2_PH(3)
      // this warp's 64 columns have left TMEM: hand the buffer back (the 16th arrival lets the MMA warp refill it)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[b]);
      SB2_PH(4)
      if (!NOSCAN) { if (estimating) est_chunk(L.ls, v); else scan_chunk(L, id, v, cbase + 32, n_points); }
      SB2_PH(5)
      if (!estimating) {
        if ((c & 15) == 0 && warp == 0 && lane == 0) atomicAdd(&g_knn2_front, 1ull);   // 16 more visits of the pack
        cbase += CN;
        if (++ctile == n_bt) { ctile = 0; cbase = cq * 64; }
      }
    }
#undef SB2_PH
    if (phw) {
#pragma unroll
      for (int i = 0; i < 6; ++i) g_knn2_phase[warp == 2 ? 0 : 1][i] = (unsigned long long)ph[i];
    }
    if (stamp) {
      g_knn2_stamp[0] = clock64() - st_c;
      g_knn2_stamp[1] = gtimer_ns() - st_t;
      g_knn2_stamp[2] = (unsigned long long)n_visit;
    }
    if (valid) {
#pragma unroll
      for (int i = 0; i < LS; ++i) sc[i] = L.ls[i];
    }
  }
  // no CTA-wide barrier down here: code after the role branches would make ptxas treat the MMA warp's region as
  // divergent (its operands then leave the uniform datapath).  The MMA warp tears TMEM down by itself.
}

}  // namespace

bool knn_tc2_supported(int d) { return 3 * d + 3 <= 512; }

bool knn_tc2_shape(const sb2_ctx* ctx, int d, int terms, bool want_estimate, KnnTc2Shape* out) {
  // one 128-row query image + a ring of candidate-image K-slices (256 rows each) in 227 KB; the K axis is staged in
  // 1, 2, 4 or 8 slices, fewest first; at least 3 ring stages (2 when nothing else fits)
  const size_t cap = ctx->prop.sharedMemPerBlockOptin;
  for (int min_st = 3; min_st >= 2; --min_st) {
    for (int ns = 1; ns <= 8; ns *= 2) {
      const int unit = 16 * ns;
      const int kp = ((terms * d + 3 + unit - 1) / unit) * unit;
      const size_t ta = (size_t)QM * kp * 2, part = (size_t)CN * kp * 2 / ns;
      for (int est = want_estimate ? 1 : 0; est >= 0; --est) {
        const size_t fixed = BAR_BYTES + ta + (est ? EST_BYTES : 0);
        if (fixed + (size_t)min_st * part > cap) continue;
        int nst = (int)((cap - fixed) / part);
        if (nst > MAXST) nst = MAXST;
        out->nsplit = ns; out->nstage = nst; out->kpad = kp; out->terms = terms; out->est = est;
        out->smem = fixed + (size_t)nst * part;
        return true;
      }
    }
  }
  return false;
}

size_t knn_tc2_a_halves(const KnnTc2Shape& sh, int64_t n_rows) { return (size_t)(ceil_div64(n_rows, QM) + 1) * QM * sh.kpad; }
size_t knn_tc2_b_halves(const KnnTc2Shape& sh, int64_t n_rows) { return (size_t)(ceil_div64(n_rows, CN) + 1) * CN * sh.kpad; }

int32_t knn_tc2_build_images(sb2_ctx* ctx, const KnnTc2Shape& sh, const float* d_x, int64_t n_rows, int d,
                             const unsigned int* d_maxnorm_bits, const int32_t* d_gather, int64_t gather_base,
                             __half* Aimg, __half* Bimg, float* d_inv_s2, float* d_dnorm, unsigned int* d_dmax_bits) {
  if (Aimg) {
    const int64_t nt = ceil_div64(n_rows, QM) + 1;
    knn_tc2_prep_kernel<<<(unsigned)nt, 256, 0, ctx->stream>>>(d_x, n_rows, d, sh.kpad, sh.terms, QM, 1, d_maxnorm_bits, d_gather,
                                                                gather_base, Aimg, d_inv_s2, Bimg ? nullptr : d_dnorm,
                                                                Bimg ? nullptr : d_dmax_bits);
    SB2_LAUNCH_CHECK(ctx);
  }
  if (Bimg) {
    const int64_t nt = ceil_div64(n_rows, CN) + 1;
    knn_tc2_prep_kernel<<<(unsigned)nt, 256, 0, ctx->stream>>>(d_x, n_rows, d, sh.kpad, sh.terms, CN, 0, d_maxnorm_bits, d_gather,
                                                                gather_base, Bimg, d_inv_s2, d_dnorm, d_dmax_bits);
    SB2_LAUNCH_CHECK(ctx);
  }
  return SB2_OK;
}

namespace {
template <int NG, int MODE>
cudaError_t launch2(unsigned grid, const KnnTc2Shape& sh, cudaStream_t st, const __half* A, const __half* B, int64_t n_tiles,
                    int64_t n_est, int64_t est_stride, int64_t qtile0, int64_t n_query, int32_t n_points, float* cs, int32_t* ci) {
  cudaError_t e = cudaFuncSetAttribute(knn_sweep2_kernel<NG, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem);
  if (e != cudaSuccess) return e;
  const int nks = sh.kpad / 16 / sh.nsplit;
  knn_sweep2_kernel<NG, MODE><<<grid, THREADS, sh.smem, st>>>(A, B, nks, sh.nsplit, sh.nstage, (uint32_t)(CN * sh.kpad * 2 / sh.nsplit) >> 4,
                                                                (uint32_t)(QM * sh.kpad * 2) >> 4, (int)(n_tiles + n_est), n_tiles, n_est,
                                                                est_stride, qtile0, n_query, n_points, cs, ci);
  return cudaSuccess;
}
}  // namespace

int32_t knn_tc2_sweep(sb2_ctx* ctx, const KnnTc2Shape& sh, const __half* Aimg, int64_t a_tile0, const __half* Bimg,
                      int64_t n_points, int64_t n_query, int list_m, float* cand_score, int32_t* cand_idx, double* issued_flops,
                      bool estimate) {
  cudaStream_t st = ctx->stream;
  const int64_t n_tiles = ceil_div64(n_points, CN);
  const int64_t q_ctas = ceil_div64(n_query, QM);
  SB2_CHECK_ARG(list_m == 64 || list_m == 128, "list_m must be 64 or 128 (four sub-lists of 16 or 32)");
  // threshold estimate: the 6th largest of the row's chunk maxima over every est_stride-th tile sits at about rank
  // EST_R * est_stride of the row; aim at ~96 (1.5 x the 64 proposals kept), i.e. a 1/16 sample
  int64_t est_stride = 96 / EST_R, n_est = n_tiles / est_stride;
  const char* est_env = getenv("SB2_KNN_EST");
  if (est_env) est_stride = atoi(est_env) > 0 ? atoi(est_env) : est_stride, n_est = atoi(est_env) > 0 ? n_tiles / est_stride : 0;
  if (n_tiles < 32 * est_stride || !estimate || !sh.est) n_est = 0;  // small problems / re-sweeps: cold start
  if (issued_flops) *issued_flops += 2.0 * (double)(q_ctas * QM) * (double)((n_tiles + n_est) * CN) * (double)sh.kpad;
  const unsigned grid = (unsigned)q_ctas;
  const int32_t np = (int32_t)n_points;
  const char* ns_env = getenv("SB2_KNN2_NOSCAN");   // timing experiments: 1 = epilogue loads but does not examine, 2 = per-phase cycle stamps
  const int mode = ns_env ? atoi(ns_env) : 0;
  {
    void* fp = nullptr;
    SB2_CUDA(cudaGetSymbolAddress(&fp, g_knn2_front));
    SB2_CUDA(cudaMemsetAsync(fp, 0, sizeof(unsigned long long), st));
  }
  cudaError_t le;
#define SB2_L2(NGV, MODEV) launch2<NGV, MODEV>(grid, sh, st, Aimg, Bimg, n_tiles, n_est, est_stride, a_tile0, n_query, np, cand_score, cand_idx)
  if (list_m == 64) le = mode == 1 ? SB2_L2(2, 1) : (mode == 2 ? SB2_L2(2, 2) : SB2_L2(2, 0));
  else le = SB2_L2(4, 0);
#undef SB2_L2
  SB2_CUDA(le);
  SB2_LAUNCH_CHECK(ctx);
  if (getenv("SB2_KNN2_STAMP")) {
    unsigned long long h[4] = {0, 0, 0, 0};
    SB2_CUDA(cudaStreamSynchronize(st));
    SB2_CUDA(cudaMemcpyFromSymbol(h, g_knn2_stamp, sizeof(h)));
    if (h[2] > 0)
      fprintf(stderr, "[knn2 stamp] grid %u, CTA %u: %llu visits, %.1f cycles/visit (tensor floor %d), SM clock under load %.0f MHz\n", grid,
              grid / 2, h[2], (double)h[0] / (double)h[2], 4 * (sh.kpad / 16) * (CN / 2) / 4, 1e3 * (double)h[0] / (double)h[1]);
    if (mode == 2 && h[2] > 0) {
      unsigned long long p[2][8];
      SB2_CUDA(cudaMemcpyFromSymbol(p, g_knn2_phase, sizeof(p)));
      for (int w = 0; w < 2; ++w)
        fprintf(stderr, "[knn2 phases] warp %d cycles/visit: wait tfull %.0f | ld0 %.0f | scan0 %.0f | ld1 %.0f | release %.0f | scan1 %.0f\n", w ? 5 : 2,
                (double)p[w][0] / h[2], (double)p[w][1] / h[2], (double)p[w][2] / h[2], (double)p[w][3] / h[2], (double)p[w][4] / h[2], (double)p[w][5] / h[2]);
      ;
    }
    if (mode == 2 && h[2] > 0) {
      unsigned long long p[2][8];
      SB2_CUDA(cudaMemcpyFromSymbol(p, g_knn2_phase, sizeof(p)));
      fprintf(stderr, "[knn2 phases] MMA issuer cycles/visit: wait full (ring) %.0f | wait tempty (epilogue) %.0f | issue + commits %.0f\n",
              (double)p[0][6] / h[2], (double)p[0][7] / h[2], (double)p[1][6] / h[2]);
    }
  }
  return SB2_OK;
}

void knn_tc2_error_coefs(const KnnTc2Shape& sh, double* c_q, double* c_n) {
  // |s_computed - s_true| <= c_q |q| R + c_n R^2 / 2   (R = largest norm in the data set, everything after scaling)
  //   fp32 accumulation in the tensor pipe: kpad/16 accumulator updates of one ulp each plus the alignment loss
  //   inside a 16-product group, bounded by 1.6 * kpad * 2^-24 of sum |a_i b_i| <= |q| R + R^2 / 2;
  //   three-way fp16 split of the norm: 2^-33 R^2/2; sub-normal halves: < 2^-27 R^2/2 since R >= 100 after scaling
  //   terms = 3: operand split 3 * 2^-24 |q| R, dropped lo*lo term 2^-22 * 2^-2 |q| R
  //   terms = 1: |q.c - q_hi.c_hi| <= |q - q_hi| |c| + |q_hi| |c - c_hi|: the re-score kernel adds this term from
  //              the measured residual norms (dnorm / dmax of knn_tc2_build_images), not from the 2^-11 worst case
  // tests/test_gpu_parity.py::test_knn_tensor_score_error_within_bound measures the actual error against this bound.
  const double u24 = 5.9604644775390625e-08;
  const double acc = 1.6 * sh.kpad * u24;
  *c_n = acc + 8.0 * u24;
  *c_q = acc + (sh.terms == 3 ? 8.0 * u24 : 0.0);
}
