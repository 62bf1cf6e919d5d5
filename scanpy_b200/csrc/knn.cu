// knn.cu — exact brute-force euclidean kNN on the n x d float32 embedding (sm_100a).
//
// Replaces sklearn's KNeighborsTransformer(algorithm='brute') as built by the reference at
// src/scanpy/neighbors/__init__.py:754-768 (ArgKmin: GEMM-shaped middle term with fp64
// accumulation + per-row heaps).  Pipeline:
//
//   knn_prep_kernel     X[n,d] row-major  ->  Xt: tiles of 128 points, k-major [d+1][128]; the extra
//                       row holds hn_j = -|x_j|^2/2, so one 1-D bulk (TMA) copy stages a whole tile.
//   knn_pass1_kernel    CTA = 128 queries, sweeps all candidate tiles through a 3-stage
//                       cp.async.bulk + mbarrier ring.  8 compute warps, each owns 16 query rows
//                       exclusively; a lane's register micro-tile is 16 queries x 4 candidates
//                       (64 FFMA per k-step).  score = q.c - |c|^2/2 (larger = closer; d^2 = |q|^2 - 2 score)
//                       lands in the accumulators with no epilogue arithmetic (accumulators start
//                       at hn).  Per query row the warp keeps the best 32 scores as a sorted list
//                       distributed over its lanes (one entry per lane, in registers); a candidate
//                       is examined only if it beats the row's current 32nd best (one FSETP per
//                       element + one ballot per row and tile; insertions are O(32 ln n) per row).
//   knn_rescore_kernel  warp per query: exact fp64 |q-c|^2 of the 32 proposals, warp bitonic sort by
//                       (d^2, id), self forced to column 0, top-k out.  Certificate: every point
//                       not proposed has fp32 score <= s32, hence true d^2 >= |q|^2 - 2 (s32 + eps);
//                       if the exact k-th d^2 is below that bound the row is provably exact.
//   knn_fallback_kernel rows without a certificate (ties, duplicates, pathological scales) are
//                       recomputed exactly in fp64 against all points.
#include <float.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "knn_internal.cuh"

namespace {

constexpr int TILE = 128;      // points per tile (queries per CTA, candidates per stage)
constexpr int LISTM = 32;      // proposals kept per query (one per lane)
constexpr int NWARP_C = 8;     // compute warps
constexpr int PASS1_THREADS = (NWARP_C + 1) * 32;

// ------------------------------------------------------------------------------------------------
__global__ void knn_prep_kernel(const float* __restrict__ X, int64_t n, int d, float* __restrict__ Xt,
                                unsigned int* __restrict__ maxnorm_bits) {
  extern __shared__ float s[];  // [128][d] staged chunk
  const int64_t t = blockIdx.x;
  const int64_t p0 = t * TILE;
  const int rows = (n - p0 < TILE) ? (int)(n - p0) : TILE;
  const float* src = X + p0 * d;
  for (int i = threadIdx.x; i < TILE * d; i += blockDim.x) s[i] = (i < rows * d) ? src[i] : 0.0f;
  __syncthreads();
  float* dst = Xt + t * (int64_t)(d + 1) * TILE;
  for (int i = threadIdx.x; i < TILE * d; i += blockDim.x) {
    int k = i / TILE, j = i % TILE;
    dst[i] = s[j * d + k];
  }
  if (threadIdx.x < TILE) {
    int j = threadIdx.x;
    double acc = 0.0;
    for (int k = 0; k < d; ++k) {
      double v = s[j * d + k];
      acc += v * v;
    }
    float hn = (j < rows) ? (float)(-0.5 * acc) : -INFINITY;
    dst[d * TILE + j] = hn;
    if (j < rows) atomicMax(maxnorm_bits, __float_as_uint((float)acc * (1.0f + 1e-6f)));
  }
}

// ------------------------------------------------------------------------------------------------
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
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int NST>
__global__ void __launch_bounds__(PASS1_THREADS, 1)
knn_pass1_kernel(const float* __restrict__ Xt, int d, int64_t n_tiles, int64_t qtile0, int64_t n_query,
                 float* __restrict__ cand_score, int32_t* __restrict__ cand_idx) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int chunk_f = (d + 1) * TILE;              // floats per tile chunk
  const uint32_t chunk_b = (uint32_t)chunk_f * 4u;  // bytes (multiple of 512)
  float* As = reinterpret_cast<float*>(smem_raw);
  float* Bs0 = As + chunk_f;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Bs0 + (size_t)NST * chunk_f);
  uint64_t* full = bars;            // [NST]
  uint64_t* empty = bars + NST;     // [NST]
  uint64_t* afull = bars + 2 * NST;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NWARP_C);
    }
    mbar_init(afull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int64_t qt = qtile0 + blockIdx.x;
  if (warp == NWARP_C) {
    // ---------------- producer warp: one elected lane drives the bulk-copy ring ----------------
    if (lane == 0) {
      mbar_expect_tx(afull, chunk_b);
      bulk_g2s(As, Xt + qt * chunk_f, chunk_b, afull);
      for (int64_t c = 0; c < n_tiles; ++c) {
        const int s = (int)(c % NST);
        const int64_t use = c / NST;
        if (use > 0) mbar_wait(&empty[s], (uint32_t)((use - 1) & 1));
        mbar_expect_tx(&full[s], chunk_b);
        bulk_g2s(Bs0 + (size_t)s * chunk_f, Xt + c * chunk_f, chunk_b, &full[s]);
      }
    }
    return;
  }

  // ---------------- compute warps ----------------
  float lv[16];
  int32_t li[16];
  float th[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    lv[r] = -INFINITY;
    li[r] = -1;
    th[r] = -INFINITY;
  }
  mbar_wait(afull, 0);
  const float* Aw = As + 16 * warp;

  for (int64_t c = 0; c < n_tiles; ++c) {
    const int s = (int)(c % NST);
    mbar_wait(&full[s], (uint32_t)((c / NST) & 1));
    const float* Bs = Bs0 + (size_t)s * chunk_f + 4 * lane;

    float acc[16][4];
    {
      const float4 h = *reinterpret_cast<const float4*>(Bs + d * TILE);
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        acc[r][0] = h.x; acc[r][1] = h.y; acc[r][2] = h.z; acc[r][3] = h.w;
      }
    }
#pragma unroll 2
    for (int k = 0; k < d; ++k) {
      const float4 b = *reinterpret_cast<const float4*>(Bs + k * TILE);
      const float4 a0 = *reinterpret_cast<const float4*>(Aw + k * TILE);
      const float4 a1 = *reinterpret_cast<const float4*>(Aw + k * TILE + 4);
      const float4 a2 = *reinterpret_cast<const float4*>(Aw + k * TILE + 8);
      const float4 a3 = *reinterpret_cast<const float4*>(Aw + k * TILE + 12);
      const float a[16] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w,
                           a2.x, a2.y, a2.z, a2.w, a3.x, a3.y, a3.z, a3.w};
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        acc[r][0] = fmaf(a[r], b.x, acc[r][0]);
        acc[r][1] = fmaf(a[r], b.y, acc[r][1]);
        acc[r][2] = fmaf(a[r], b.z, acc[r][2]);
        acc[r][3] = fmaf(a[r], b.w, acc[r][3]);
      }
    }
    // release the stage as early as possible (all smem reads of this stage are done)
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);

    const int32_t base = (int32_t)(c * TILE);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const bool any = (acc[r][0] > th[r]) | (acc[r][1] > th[r]) | (acc[r][2] > th[r]) | (acc[r][3] > th[r]);
      if (__ballot_sync(0xffffffffu, any)) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          unsigned mj = __ballot_sync(0xffffffffu, acc[r][j] > th[r]);
          while (mj) {
            const int src = __ffs(mj) - 1;
            mj &= mj - 1;
            const float v = __shfl_sync(0xffffffffu, acc[r][j], src);
            if (v > th[r]) {  // warp-uniform: th may have risen since the ballot
              const int32_t id = base + 4 * src + j;
              const int pos = __popc(__ballot_sync(0xffffffffu, lv[r] >= v));
              const float up_v = __shfl_up_sync(0xffffffffu, lv[r], 1);
              const int32_t up_i = __shfl_up_sync(0xffffffffu, li[r], 1);
              if (lane > pos) { lv[r] = up_v; li[r] = up_i; }
              else if (lane == pos) { lv[r] = v; li[r] = id; }
              th[r] = __shfl_sync(0xffffffffu, lv[r], 31);
            }
          }
        }
      }
    }
  }

  // proposals out: row-major [n_query][32], lane = rank (0 = best)
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const int64_t q = (int64_t)blockIdx.x * TILE + 16 * warp + r;
    if (q < n_query) {
      cand_score[q * LISTM + lane] = lv[r];
      cand_idx[q * LISTM + lane] = li[r];
    }
  }
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool lex_less(double ka, int ia, double kb, int ib) {
  return ka < kb || (ka == kb && ia < ib);
}

// R = proposals per lane: the list has 32*R entries.  The list is made of sub-lists of `sub` consecutive slots (sub = 32*R
// for a single list; the second-generation tensor sweep keeps four sub-lists per row, one per column quarter of a
// tile).  Every sub-list covers a disjoint part of the candidates and its smallest stored score (a real proposal or
// the sentinel it started from) bounds the score of every candidate of that part that is NOT in it; hence no
// unproposed candidate scored above tau = max over sub-lists of their minima.
template <int R>
__global__ void knn_rescore_kernel(const float* __restrict__ X, int64_t n_points, int d, int64_t q0, int64_t ql_base,
                                   const int32_t* __restrict__ row_map, int64_t n_rows,
                                   int k, int sub, const float* __restrict__ cand_score, const int32_t* __restrict__ cand_idx,
                                   const unsigned int* __restrict__ maxnorm_bits, const float* __restrict__ inv_s2,
                                   double c_q, double c_n, const float* __restrict__ dnorm,
                                   const unsigned int* __restrict__ dmax_bits, int32_t* __restrict__ idx_out,
                                   double* __restrict__ dist_out, int32_t* __restrict__ work_q,
                                   double* __restrict__ work_ub, unsigned long long* __restrict__ work_cnt) {
  constexpr int LM = 32 * R;
  const int lane = threadIdx.x & 31;
  // launch row i holds the proposals of local query ql (= row_map[i] for a gathered re-sweep, else ql_base + i);
  // outputs and queue entries are indexed by ql, the point itself is q0 + ql
  const int64_t li = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (li >= n_rows) return;
  const int64_t ql = row_map ? (int64_t)row_map[li] : ql_base + li;
  const int64_t q = q0 + ql;
  const float* xq = X + q * d;
  double key[R];
  int32_t id[R];
  float cs[R];
  double qn = 0.0;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int32_t ci = cand_idx[li * LM + r * 32 + lane];
    cs[r] = cand_score[li * LM + r * 32 + lane];  // unused slots hold the sweep's starting threshold (-inf for a cold start)
    const float* xc = X + (int64_t)(ci < 0 ? 0 : ci) * d;
    double acc = 0.0, qq = 0.0;
    for (int j = 0; j < d; ++j) {
      const double a = xq[j];
      const double df = a - (double)xc[j];
      acc = fma(df, df, acc);
      qq = fma(a, a, qq);
    }
    qn = qq;
    key[r] = ci < 0 ? DBL_MAX : ((ci == q) ? -1.0 : acc);
    id[r] = ci < 0 ? INT32_MAX : ci;
  }
  // tau = max over sub-lists of (min over the sub-list's slots)
  float tau = -INFINITY;
  if (sub >= 32) {
    const int regs_per = sub >> 5;
#pragma unroll
    for (int r0 = 0; r0 < R; ++r0) {
      if (r0 % regs_per == 0) {
        float m = INFINITY;
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (r >= r0 && r < r0 + regs_per) m = fminf(m, cs[r]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
        tau = fmaxf(tau, m);
      }
    }
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float m = cs[r];
      for (int o = sub >> 1; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));   // min inside the sub-list's lanes
      for (int o = 16; o >= sub; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));      // max across the sub-lists of this register
      tau = fmaxf(tau, m);
    }
  }
  // bitonic sort of the 32*R elements (element e = r*32 + lane) ascending by (key, id)
#pragma unroll
  for (int kk = 2; kk <= LM; kk <<= 1) {
#pragma unroll
    for (int j = kk >> 1; j > 0; j >>= 1) {
      if (j >= 32) {  // partner lives in the same lane, register r ^ (j / 32)
        const int rj = j >> 5;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if ((r & rj) == 0 && (r | rj) < R) {
            const int r2 = r | rj;
            const bool asc = ((r * 32) & kk) == 0;
            const bool swap = asc ? lex_less(key[r2], id[r2], key[r], id[r]) : lex_less(key[r], id[r], key[r2], id[r2]);
            if (swap) {
              const double tk = key[r]; key[r] = key[r2]; key[r2] = tk;
              const int32_t ti = id[r]; id[r] = id[r2]; id[r2] = ti;
            }
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int e = r * 32 + lane;
          const double ok = __shfl_xor_sync(0xffffffffu, key[r], j);
          const int32_t oi = __shfl_xor_sync(0xffffffffu, id[r], j);
          const bool want_min = ((e & j) == 0) == ((e & kk) == 0);
          const bool other_less = lex_less(ok, oi, key[r], id[r]);
          const bool take = want_min ? other_less : lex_less(key[r], id[r], ok, oi);
          if (take) { key[r] = ok; id[r] = oi; }
        }
      }
    }
  }
  // the query itself must sit in column 0 (src/scanpy/neighbors/_common.py:74-98); if it was not proposed at
  // all (a flood of exact duplicates can push it out of the list) the row is not certified.
  const bool self_first = __shfl_sync(0xffffffffu, id[0], 0) == (int32_t)q;
  double kth = 0.0;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const double t = __shfl_sync(0xffffffffu, key[r], (k - 1) & 31);
    if (r == ((k - 1) >> 5)) kth = t;
  }
  bool certified;
  if (tau == -INFINITY) {
    certified = true;  // cold start and no sub-list ever filled: the proposal list is the whole data set
  } else {
    const double Rn = sqrt((double)__uint_as_float(*maxnorm_bits));
    double eps = c_n * 0.5 * Rn * Rn + c_q * sqrt(qn) * Rn;
    if (dnorm) {  // fp16 tier: |q.c - q_hi.c_hi| <= |dq| R + (|q| + |dq|) max|dc|
      const double dq = (double)dnorm[q], dc = (double)__uint_as_float(*dmax_bits);
      eps += (dq * Rn + (sqrt(qn) + dq) * dc) * (1.0 + 1e-6);
    }
    const double bound = qn - 2.0 * ((double)tau * (inv_s2 ? (double)*inv_s2 : 1.0) + eps);
    certified = self_first && (kth < bound);
    if (inv_s2) {
      // the tensor-path bounds are derived for a scaled largest norm in [100, 200) (the prep kernel's power-of-two
      // scale); data so tiny or so huge that the clamped scale cannot reach it (largest norm below ~1e-16 or above
      // ~2e20) is left to the exact scan.  Also catches non-finite scores.
      const double r2s2 = Rn * Rn / (double)*inv_s2;
      if (!(r2s2 >= 9801.0 && r2s2 <= 67600.0)) certified = false;
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int e = r * 32 + lane;
    if (e < k) {
      idx_out[ql * k + e] = id[r];
      dist_out[ql * k + e] = key[r] < 0.0 ? 0.0 : sqrt(key[r]);
    }
  }
  if (!certified && lane == 0) {
    const unsigned long long w = atomicAdd(work_cnt, 1ull);
    work_q[w] = (int32_t)ql;
    work_ub[w] = self_first ? kth : DBL_MAX;
  }
}

// a handful of open rows is cheaper to scan exactly than to sweep again: move queue 1 behind queue 2
__global__ void knn_append_queue_kernel(const int32_t* __restrict__ q_src, const double* __restrict__ ub_src, int64_t n,
                                        int32_t* __restrict__ q_dst, double* __restrict__ ub_dst,
                                        unsigned long long* __restrict__ cnt_dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long w = atomicAdd(cnt_dst, 1ull);
  q_dst[w] = q_src[i];
  ub_dst[w] = ub_src[i];
}

// exact fallback: one CTA per uncertified query, fp64 distances to every point
constexpr int FB_THREADS = 256;
constexpr int FB_BUF = 2048;

__device__ void block_bitonic_sort(double* keys, int32_t* ids, int n_pow2) {
  for (int kk = 2; kk <= n_pow2; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const bool asc = (i & kk) == 0;
          const bool gt = lex_less(keys[ixj], ids[ixj], keys[i], ids[i]);
          if (gt == asc) {
            double tk = keys[i]; keys[i] = keys[ixj]; keys[ixj] = tk;
            int32_t ti = ids[i]; ids[i] = ids[ixj]; ids[ixj] = ti;
          }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(FB_THREADS)
knn_fallback_kernel(const float* __restrict__ X, const float* __restrict__ Xt, int64_t n_points, int d, int64_t q0,
                    int k, const int32_t* __restrict__ work_q, const double* __restrict__ work_ub,
                    const unsigned long long* __restrict__ work_cnt, int32_t* __restrict__ idx_out,
                    double* __restrict__ dist_out) {
  __shared__ double keys[FB_BUF];
  __shared__ int32_t ids[FB_BUF];
  __shared__ double qx[256];
  __shared__ int cnt;
  __shared__ double thr_key;
  __shared__ int32_t thr_id;
  const unsigned long long nwork = *work_cnt;
  for (unsigned long long w = blockIdx.x; w < nwork; w += gridDim.x) {
    const int64_t ql = work_q[w];
    const int64_t q = q0 + ql;
    __syncthreads();
    for (int j = threadIdx.x; j < d; j += blockDim.x) qx[j] = X[q * d + j];
    if (threadIdx.x == 0) {
      cnt = 0;
      thr_key = work_ub[w];
      thr_id = INT32_MAX;
    }
    __syncthreads();
    for (int64_t base = 0; base < n_points; base += FB_THREADS) {
      const int64_t p = base + threadIdx.x;
      if (p < n_points) {
        const float* col = Xt + (p / TILE) * (int64_t)(d + 1) * TILE + (p % TILE);
        double acc = 0.0;
        for (int j = 0; j < d; ++j) {
          const double df = qx[j] - (double)col[(int64_t)j * TILE];
          acc = fma(df, df, acc);
        }
        const double key = (p == q) ? -1.0 : acc;
        if (key < thr_key || (key == thr_key && (int32_t)p <= thr_id)) {
          const int pos = atomicAdd(&cnt, 1);
          keys[pos] = key;
          ids[pos] = (int32_t)p;
        }
      }
      __syncthreads();
      const bool compact = cnt > FB_BUF - FB_THREADS;  // snapshot: every thread takes the same branch
      __syncthreads();                                 // ... before anybody's next atomicAdd can change cnt
      if (compact) {
        const int c0 = cnt;
        for (int i = c0 + threadIdx.x; i < FB_BUF; i += blockDim.x) { keys[i] = DBL_MAX; ids[i] = INT32_MAX; }
        __syncthreads();
        block_bitonic_sort(keys, ids, FB_BUF);
        if (threadIdx.x == 0) {
          cnt = k;
          thr_key = keys[k - 1];
          thr_id = ids[k - 1];
        }
        __syncthreads();
      }
    }
    const int c0 = cnt;
    for (int i = c0 + threadIdx.x; i < FB_BUF; i += blockDim.x) { keys[i] = DBL_MAX; ids[i] = INT32_MAX; }
    __syncthreads();
    block_bitonic_sort(keys, ids, FB_BUF);
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
      idx_out[ql * k + i] = ids[i];
      dist_out[ql * k + i] = keys[i] < 0.0 ? 0.0 : sqrt(keys[i]);
    }
  }
}

}  // namespace

namespace {

struct EventPairs {  // CUDA-event brackets around the sweep kernels of one call (info != NULL only)
  cudaEvent_t ev[8][2];
  int n = 0;
  bool on = false;
  cudaStream_t st = nullptr;
  int32_t begin() {
    if (!on || n >= 8) return SB2_OK;
    SB2_CUDA(cudaEventCreate(&ev[n][0]));
    SB2_CUDA(cudaEventCreate(&ev[n][1]));
    SB2_CUDA(cudaEventRecord(ev[n][0], st));
    return SB2_OK;
  }
  int32_t end() {
    if (!on || n >= 8) return SB2_OK;
    SB2_CUDA(cudaEventRecord(ev[n][1], st));
    ++n;
    return SB2_OK;
  }
  float total_ms() {
    float t = 0.0f;
    for (int i = 0; i < n; ++i) {
      float ms = 0.0f;
      if (cudaEventElapsedTime(&ms, ev[i][0], ev[i][1]) == cudaSuccess) t += ms;
    }
    return t;
  }
  ~EventPairs() {
    for (int i = 0; i < n; ++i) { cudaEventDestroy(ev[i][0]); cudaEventDestroy(ev[i][1]); }
  }
};

int32_t launch_rescore(sb2_ctx* ctx, int list_m, int sub, const float* d_x, int64_t n_points, int d, int64_t q0, int64_t ql_base,
                       const int32_t* row_map, int64_t n_rows, int k, const float* cand_score, const int32_t* cand_idx,
                       const unsigned int* maxnorm, const float* inv_s2, double c_q, double c_n, const float* dnorm,
                       const unsigned int* dmax_bits, int32_t* d_idx,
                       double* d_dist, int32_t* wq, double* wub, unsigned long long* wcnt) {
  if (n_rows == 0) return SB2_OK;
  const int wpb = 8;
  const unsigned grid = (unsigned)ceil_div64(n_rows, wpb);
#define SB2_RESCORE(RR)                                                                                                          \
  knn_rescore_kernel<RR><<<grid, wpb * 32, 0, ctx->stream>>>(d_x, n_points, d, q0, ql_base, row_map, n_rows, k, sub, cand_score, \
                                                             cand_idx, maxnorm, inv_s2, c_q, c_n, dnorm, dmax_bits, d_idx, d_dist, wq, wub, wcnt)
  if (list_m == 32) SB2_RESCORE(1);
  else if (list_m == 64) SB2_RESCORE(2);
  else SB2_RESCORE(4);
#undef SB2_RESCORE
  SB2_LAUNCH_CHECK(ctx);
  return SB2_OK;
}

int32_t read_count(sb2_ctx* ctx, const unsigned long long* d_cnt, int64_t* out) {
  unsigned long long h = 0;
  SB2_CUDA(cudaMemcpyAsync(&h, d_cnt, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  SB2_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = (int64_t)h;
  return SB2_OK;
}

// ---- the two generations of the tensor-core sweep behind one interface (tensor_tiers below) ----
struct SweepV1 {   // knn_tc.cu: 256-query CTAs x 128-candidate tiles, one 32/64-proposal list per row (SB2_KNN_V=1)
  KnnTcShape sh;
  bool shape(const sb2_ctx* ctx, int d, int terms) { return knn_tc_shape(ctx, d, terms, &sh); }
  static int list_m(int k, bool force64) { return (k > 24 || force64) ? 64 : 32; }
  static int sub(int list_m) { return list_m; }
  int wave_rows(const sb2_ctx* ctx) const { return ctx->prop.multiProcessorCount * sh.qh * TILE; }
  size_t a_halves(int64_t n) const { return knn_tc_image_halves(sh, n); }
  size_t b_halves(int64_t n) const { return knn_tc_image_halves(sh, n); }
  int32_t build(sb2_ctx* ctx, const float* x, int64_t n, int d, const unsigned int* mn, const int32_t* g, int64_t gb, __half* A,
                __half* B, float* inv_s2, float* dnorm, unsigned int* dmax) const {
    return knn_tc_build_images(ctx, sh, x, n, d, mn, g, gb, A, B, inv_s2, dnorm, dmax);
  }
  int32_t sweep(sb2_ctx* ctx, const __half* A, int64_t a_tile0, const __half* B, int64_t n_points, int64_t n_query, int lm,
                float* cs, int32_t* ci, double* fl, bool est) const {
    return knn_tc_sweep(ctx, sh, A, a_tile0, B, n_points, n_query, lm, cs, ci, fl, est);
  }
  void coefs(double* cq, double* cn) const { knn_tc_error_coefs(sh, cq, cn); }
};
struct SweepV2 {   // knn_tc2.cu: 128-query CTAs x 256-candidate tiles (UMMA N = 256), four sub-lists per row (default)
  KnnTc2Shape sh;
  bool shape(const sb2_ctx* ctx, int d, int terms) { return knn_tc2_shape(ctx, d, terms, terms == 1, &sh); }
  static int list_m(int k, bool force_big) { return (k > 24 || force_big) ? 128 : 64; }
  static int sub(int list_m) { return list_m / 4; }
  int wave_rows(const sb2_ctx* ctx) const { return ctx->prop.multiProcessorCount * TILE; }
  size_t a_halves(int64_t n) const { return knn_tc2_a_halves(sh, n); }
  size_t b_halves(int64_t n) const { return knn_tc2_b_halves(sh, n); }
  int32_t build(sb2_ctx* ctx, const float* x, int64_t n, int d, const unsigned int* mn, const int32_t* g, int64_t gb, __half* A,
                __half* B, float* inv_s2, float* dnorm, unsigned int* dmax) const {
    return knn_tc2_build_images(ctx, sh, x, n, d, mn, g, gb, A, B, inv_s2, dnorm, dmax);
  }
  int32_t sweep(sb2_ctx* ctx, const __half* A, int64_t a_tile0, const __half* B, int64_t n_points, int64_t n_query, int lm,
                float* cs, int32_t* ci, double* fl, bool est) const {
    return knn_tc2_sweep(ctx, sh, A, a_tile0, B, n_points, n_query, lm, cs, ci, fl, est);
  }
  void coefs(double* cq, double* cn) const { knn_tc2_error_coefs(sh, cq, cn); }
};

struct TierState {
  sb2_ctx* ctx;
  ScratchScope* scr;
  const float* d_x;
  int64_t n_points, q0, n_query;
  int d, k;
  unsigned int* maxnorm;
  int32_t *wq1, *wq2;
  double *wub1, *wub2;
  unsigned long long* wcnt;
  int32_t* d_idx;
  double* d_dist;
  EventPairs* evs;
  double issued_flops = 0.0;
  int64_t n_resweep = 0;
};

// Tensor path, in tiers that each end in the same fp64 re-score + certificate:
//   tier 1  one fp16 sweep (K = d+3): proposals whose rounding bound (2^-10 |q| R) still certifies the exact top-k are final
//   tier 2  rows tier 1 could not certify are gathered and swept again in split precision (K = 3d+3, bound ~2^-16 |q| R)
//   tier 3  rows still open (exact ties at the k-th distance, floods of duplicates): fp64 scan of every point (caller)
// A pilot of one wave of CTAs measures tier 1's certification rate; if most rows fail (data far from the origin,
// tiny neighbour gaps) the remaining rows go straight to the split-precision sweep.
template <class SW>
int32_t tensor_tiers(TierState& t, bool split_only, bool force_big_list) {
  sb2_ctx* ctx = t.ctx;
  ScratchScope& scr = *t.scr;
  cudaStream_t st = ctx->stream;
  const int64_t n_points = t.n_points, n_query = t.n_query, q0 = t.q0;
  const int d = t.d, k = t.k;
  SW s1, s3;
  SB2_CHECK_ARG(s1.shape(ctx, d, 1) && s3.shape(ctx, d, 3), "tensor-core kNN tile does not fit shared memory");
  const int list_m = SW::list_m(k, force_big_list), sub = SW::sub(list_m);
  float* cand_score;
  int32_t* cand_idx;
  SB2_TRY(scr.alloc(&cand_score, (size_t)n_query * list_m));
  SB2_TRY(scr.alloc(&cand_idx, (size_t)n_query * list_m));
  float* inv_s2;
  SB2_TRY(scr.alloc(&inv_s2, 4));
  double cq1, cn1, cq3, cn3;
  s1.coefs(&cq1, &cn1);
  s3.coefs(&cq3, &cn3);
  __half *A1 = nullptr, *B1 = nullptr, *A3 = nullptr, *B3 = nullptr;
  float* dnorm = nullptr;  // |x - fp16(x)| per point; its maximum lives in maxnorm[1]
  unsigned int* maxnorm = t.maxnorm;
  const int64_t qt0 = q0 / TILE;
  int64_t done = 0;  // local rows [0, done) have been through their first sweep
  bool direct_split = split_only;
  if (!split_only) {
    SB2_TRY(scr.alloc(&A1, s1.a_halves(n_points)));
    SB2_TRY(scr.alloc(&B1, s1.b_halves(n_points)));
    SB2_TRY(scr.alloc(&dnorm, (size_t)n_points));
    SB2_TRY(s1.build(ctx, t.d_x, n_points, d, maxnorm, nullptr, 0, A1, B1, inv_s2, dnorm, maxnorm + 1));
    // pilot: one wave of CTAs
    const int64_t wave = s1.wave_rows(ctx);
    const int64_t first = n_query >= 4 * wave ? wave : n_query;
    SB2_TRY(t.evs->begin());
    SB2_TRY(s1.sweep(ctx, A1, qt0, B1, n_points, first, list_m, cand_score, cand_idx, &t.issued_flops, true));
    SB2_TRY(t.evs->end());
    SB2_TRY(launch_rescore(ctx, list_m, sub, t.d_x, n_points, d, q0, 0, nullptr, first, k, cand_score, cand_idx, maxnorm, inv_s2, cq1,
                           cn1, dnorm, maxnorm + 1, t.d_idx, t.d_dist, t.wq1, t.wub1, t.wcnt));
    done = first;
    if (first < n_query) {
      int64_t open1 = 0;
      SB2_TRY(read_count(ctx, t.wcnt, &open1));
      direct_split = open1 * 2 > first;
      if (!direct_split) {
        const int64_t rest = n_query - done;
        SB2_TRY(t.evs->begin());
        SB2_TRY(s1.sweep(ctx, A1, qt0 + done / TILE, B1, n_points, rest, list_m, cand_score + done * list_m,
                         cand_idx + done * list_m, &t.issued_flops, true));
        SB2_TRY(t.evs->end());
        SB2_TRY(launch_rescore(ctx, list_m, sub, t.d_x, n_points, d, q0, done, nullptr, rest, k, cand_score + done * list_m,
                               cand_idx + done * list_m, maxnorm, inv_s2, cq1, cn1, dnorm, maxnorm + 1, t.d_idx, t.d_dist, t.wq1, t.wub1, t.wcnt));
        done = n_query;
      }
    }
  }
  if (done < n_query) {
    // split-precision sweep of the remaining rows, straight from the full image arrays
    SB2_TRY(scr.alloc(&A3, s3.a_halves(n_points)));
    SB2_TRY(scr.alloc(&B3, s3.b_halves(n_points)));
    SB2_TRY(s3.build(ctx, t.d_x, n_points, d, maxnorm, nullptr, 0, A3, B3, inv_s2, nullptr, nullptr));
    const int64_t rest = n_query - done;
    SB2_TRY(t.evs->begin());
    SB2_TRY(s3.sweep(ctx, A3, qt0 + done / TILE, B3, n_points, rest, list_m, cand_score + done * list_m,
                     cand_idx + done * list_m, &t.issued_flops, true));
    SB2_TRY(t.evs->end());
    SB2_TRY(launch_rescore(ctx, list_m, sub, t.d_x, n_points, d, q0, done, nullptr, rest, k, cand_score + done * list_m,
                           cand_idx + done * list_m, maxnorm, inv_s2, cq3, cn3, nullptr, nullptr, t.d_idx, t.d_dist, t.wq2, t.wub2, t.wcnt + 1));
  }
  if (!split_only) {
    // tier 2: gather the rows tier 1 left open and sweep them in split precision
    SB2_TRY(read_count(ctx, t.wcnt, &t.n_resweep));
    // a re-sweep occupies one CTA per 128 open rows for a whole pass over the candidates (a few ms at 1.3M points
    // however few rows there are); the exact scan streams the whole data set once per ROW (~45 us each at
    // 1.3M x 50 once HBM saturates), so it only wins for a handful of rows
    int64_t scan_slots = 16;
    if (const char* se = getenv("SB2_KNN_SCAN_SLOTS")) scan_slots = atoll(se);  // 0 forces the re-sweep (tests)
    const int64_t n_resweep = t.n_resweep;
    if (n_resweep > 0 && n_resweep <= scan_slots) {
      knn_append_queue_kernel<<<(unsigned)ceil_div64(n_resweep, 256), 256, 0, st>>>(t.wq1, t.wub1, n_resweep, t.wq2, t.wub2, t.wcnt + 1);
      SB2_LAUNCH_CHECK(ctx);
    } else if (n_resweep > 0) {
      if (!B3) {
        SB2_TRY(scr.alloc(&B3, s3.b_halves(n_points)));
        SB2_TRY(s3.build(ctx, t.d_x, n_points, d, maxnorm, nullptr, 0, nullptr, B3, inv_s2, nullptr, nullptr));
      }
      __half* Ag;
      SB2_TRY(scr.alloc(&Ag, s3.a_halves(n_resweep)));
      SB2_TRY(s3.build(ctx, t.d_x, n_resweep, d, maxnorm, t.wq1, q0, Ag, nullptr, inv_s2, nullptr, nullptr));
      SB2_TRY(t.evs->begin());
      // cold start: rows whose sampled threshold was too tight (fewer than k candidates beat it) must not fail twice
      SB2_TRY(s3.sweep(ctx, Ag, 0, B3, n_points, n_resweep, list_m, cand_score, cand_idx, &t.issued_flops, false));
      SB2_TRY(t.evs->end());
      SB2_TRY(launch_rescore(ctx, list_m, sub, t.d_x, n_points, d, q0, 0, t.wq1, n_resweep, k, cand_score, cand_idx, maxnorm, inv_s2, cq3,
                             cn3, nullptr, nullptr, t.d_idx, t.d_dist, t.wq2, t.wub2, t.wcnt + 1));
    }
  }
  return SB2_OK;
}

}  // namespace

// Exact kNN: a fast first pass proposes candidates per query, an fp64 re-score certifies the top-k against a rigorous
// rounding-error bound, rows without a certificate are recomputed exactly.  First pass: the tensor-core tiers above
// (default: knn_tc2.cu; SB2_KNN_V=1: the first-generation kernel of knn_tc.cu) or the fp32 CUDA-core sweep (SB2_KNN_PASS1=ffma).
extern "C" int32_t sb2_knn_l2_f32(sb2_ctx* ctx, int64_t n_points, int32_t d, const float* d_x, int64_t q0,
                                  int64_t n_query, int32_t k, int32_t* d_idx, double* d_dist, sb2_knn_info* info) {
  SB2_CHECK_ARG(ctx && d_x && d_idx && d_dist, "null pointer");
  SB2_CHECK_ARG(n_points >= 1 && n_points < (int64_t)INT32_MAX - 2 * TILE, "n_points");
  SB2_CHECK_ARG(d >= 1 && d <= 150, "d must be in [1,150]");
  SB2_CHECK_ARG(k >= 1 && k <= 56 && k <= n_points, "k must be in [1,56] and <= n_points");
  SB2_CHECK_ARG(q0 >= 0 && n_query >= 0 && q0 + n_query <= n_points, "query range");
  SB2_CHECK_ARG(q0 % TILE == 0, "q0 must be a multiple of 128");
  if (n_query == 0) return SB2_OK;
  SB2_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  ScratchScope scr(ctx);
  const int64_t n_tiles = ceil_div64(n_points, TILE);
  const int64_t chunk_f = (int64_t)(d + 1) * TILE;
  float* Xt;
  unsigned int* maxnorm;
  int32_t *wq1, *wq2;
  double *wub1, *wub2;
  unsigned long long* wcnt;  // [0] rows tier 1 left open, [1] rows for the exact scan
  SB2_TRY(scr.alloc(&Xt, (size_t)(n_tiles * chunk_f)));
  SB2_TRY(scr.alloc(&maxnorm, 4));
  const char* force = getenv("SB2_KNN_PASS1");
  const char* tiers = getenv("SB2_KNN_TIERS");
  const char* gen = getenv("SB2_KNN_V");
  const bool use_tc = knn_tc_supported(d) && !(force && strcmp(force, "ffma") == 0);
  const bool use_v1 = gen && strcmp(gen, "1") == 0;
  const bool split_only = tiers && strcmp(tiers, "3") == 0;
  const char* list_env = getenv("SB2_KNN_LIST");
  const bool big_list = list_env && (strcmp(list_env, "64") == 0 || strcmp(list_env, "128") == 0);
  SB2_CHECK_ARG(use_tc || k <= LISTM - 2, "k > 30 needs the tensor-core pass (SB2_KNN_PASS1=ffma limits k to 30)");
  SB2_TRY(scr.alloc(&wq1, (size_t)n_query));
  SB2_TRY(scr.alloc(&wub1, (size_t)n_query));
  SB2_TRY(scr.alloc(&wq2, (size_t)n_query));
  SB2_TRY(scr.alloc(&wub2, (size_t)n_query));
  SB2_TRY(scr.alloc(&wcnt, 2));
  SB2_CUDA(cudaMemsetAsync(maxnorm, 0, 16, st));
  SB2_CUDA(cudaMemsetAsync(wcnt, 0, 16, st));
  {
    size_t smem = (size_t)TILE * d * sizeof(float);
    SB2_CUDA(cudaFuncSetAttribute(knn_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_prep_kernel<<<(unsigned)n_tiles, 256, smem, st>>>(d_x, n_points, d, Xt, maxnorm);
    SB2_LAUNCH_CHECK(ctx);
  }
  EventPairs evs;
  evs.on = info != nullptr;
  evs.st = st;
  double issued_flops = 0.0;
  int64_t n_resweep = 0;
  if (use_tc) {
    TierState t{ctx, &scr, d_x, n_points, q0, n_query, d, k, maxnorm, wq1, wq2, wub1, wub2, wcnt, d_idx, d_dist, &evs};
    if (use_v1) SB2_TRY(tensor_tiers<SweepV1>(t, split_only, big_list));
    else SB2_TRY(tensor_tiers<SweepV2>(t, split_only, big_list));
    issued_flops = t.issued_flops;
    n_resweep = t.n_resweep;
  } else {
    float* cand_score;
    int32_t* cand_idx;
    SB2_TRY(scr.alloc(&cand_score, (size_t)n_query * LISTM));
    SB2_TRY(scr.alloc(&cand_idx, (size_t)n_query * LISTM));
    issued_flops = 2.0 * (double)n_query * (double)n_points * (double)d;
    const double c = 1.5 * (double)(d + 2) * 5.9604644775390625e-08;
    const int64_t q_tiles = ceil_div64(n_query, TILE);
    const size_t chunk_b = (size_t)chunk_f * 4;
    const size_t smem3 = chunk_b * 4 + 64, smem2 = chunk_b * 3 + 64;
    const size_t lim = ctx->prop.sharedMemPerBlockOptin;
    SB2_TRY(evs.begin());
    if (smem3 <= lim) {
      SB2_CUDA(cudaFuncSetAttribute(knn_pass1_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
      knn_pass1_kernel<3><<<(unsigned)q_tiles, PASS1_THREADS, smem3, st>>>(Xt, d, n_tiles, q0 / TILE, n_query,
                                                                           cand_score, cand_idx);
    } else {
      SB2_CHECK_ARG(smem2 <= lim, "d too large for shared memory");
      SB2_CUDA(cudaFuncSetAttribute(knn_pass1_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      knn_pass1_kernel<2><<<(unsigned)q_tiles, PASS1_THREADS, smem2, st>>>(Xt, d, n_tiles, q0 / TILE, n_query,
                                                                           cand_score, cand_idx);
    }
    SB2_LAUNCH_CHECK(ctx);
    SB2_TRY(evs.end());
    SB2_TRY(launch_rescore(ctx, LISTM, LISTM, d_x, n_points, d, q0, 0, nullptr, n_query, k, cand_score, cand_idx, maxnorm, nullptr, c, c,
                           nullptr, nullptr, d_idx, d_dist, wq2, wub2, wcnt + 1));
  }
  {
    const int grid = ctx->prop.multiProcessorCount * 4;
    knn_fallback_kernel<<<grid, FB_THREADS, 0, st>>>(d_x, Xt, n_points, d, q0, k, wq2, wub2, wcnt + 1, d_idx, d_dist);
    SB2_LAUNCH_CHECK(ctx);
  }
  if (info) {
    unsigned long long h_cnt = 0;
    unsigned int h_bits = 0;
    SB2_CUDA(cudaMemcpyAsync(&h_cnt, wcnt + 1, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
    SB2_CUDA(cudaMemcpyAsync(&h_bits, maxnorm, sizeof(h_bits), cudaMemcpyDeviceToHost, st));
    SB2_CUDA(cudaStreamSynchronize(st));
    info->n_uncertified = (int64_t)h_cnt;
    float f;
    memcpy(&f, &h_bits, 4);
    info->max_norm = sqrtf(f);
    info->pass1_ms = evs.total_ms();
    info->pass1_flops = 2.0 * (double)n_query * (double)n_points * (double)d;
    info->pass1_issued_flops = issued_flops;
    info->pass1_tensor = use_tc ? (use_v1 ? 1 : 2) : 0;
    info->n_resweep = n_resweep;
  }
  return SB2_OK;
}

// Debug / test entry: the raw proposals of ONE cold-start tensor sweep (generation 2) in the chosen operand format, with
// everything the rounding-error certificate is computed from, so that a test can MEASURE |s_tensor - s_exact| on the
// hardware against the bound knn_rescore_kernel uses (tests/test_gpu_parity.py::test_knn_tensor_score_error_within_bound).
//   d_score / d_idx [n_points x list_m]: proposal scores in the kernel's scaled units and their point ids (-1: unused)
//   d_dnorm [n_points]: |x - fp16(x)| per point (terms = 1 only, else untouched)
//   h_meta[6] = { inv_s2 (score_true = score * inv_s2), R^2 (largest squared norm), max_p dnorm[p], c_q, c_n, list_m }
extern "C" int32_t sb2_knn_debug_proposals_f32(sb2_ctx* ctx, int64_t n_points, int32_t d, const float* d_x, int32_t terms,
                                               float* d_score, int32_t* d_idx, float* d_dnorm, double* h_meta) {
  SB2_CHECK_ARG(ctx && d_x && d_score && d_idx && h_meta, "null pointer");
  SB2_CHECK_ARG(terms == 1 || terms == 3, "terms must be 1 (fp16) or 3 (split fp16)");
  SB2_CHECK_ARG(n_points >= 1 && d >= 1 && d <= 150, "n_points / d");
  SB2_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  ScratchScope scr(ctx);
  const int64_t n_tiles = ceil_div64(n_points, TILE);
  float* Xt;
  unsigned int* maxnorm;
  float* inv_s2;
  SB2_TRY(scr.alloc(&Xt, (size_t)(n_tiles * (d + 1) * TILE)));
  SB2_TRY(scr.alloc(&maxnorm, 4));
  SB2_TRY(scr.alloc(&inv_s2, 4));
  SB2_CUDA(cudaMemsetAsync(maxnorm, 0, 16, st));
  {
    size_t smem = (size_t)TILE * d * sizeof(float);
    SB2_CUDA(cudaFuncSetAttribute(knn_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_prep_kernel<<<(unsigned)n_tiles, 256, smem, st>>>(d_x, n_points, d, Xt, maxnorm);
    SB2_LAUNCH_CHECK(ctx);
  }
  KnnTc2Shape sh;
  SB2_CHECK_ARG(knn_tc2_shape(ctx, d, terms, false, &sh), "tile does not fit shared memory");
  __half *A, *B;
  SB2_TRY(scr.alloc(&A, knn_tc2_a_halves(sh, n_points)));
  SB2_TRY(scr.alloc(&B, knn_tc2_b_halves(sh, n_points)));
  float* dn = d_dnorm;
  if (!dn) SB2_TRY(scr.alloc(&dn, (size_t)n_points));
  SB2_TRY(knn_tc2_build_images(ctx, sh, d_x, n_points, d, maxnorm, nullptr, 0, A, B, inv_s2, terms == 1 ? dn : nullptr,
                               terms == 1 ? maxnorm + 1 : nullptr));
  const int list_m = 64;
  double fl = 0.0;
  SB2_TRY(knn_tc2_sweep(ctx, sh, A, 0, B, n_points, n_points, list_m, d_score, d_idx, &fl, false));
  float h_inv = 0.0f;
  unsigned int h_bits[2] = {0, 0};
  SB2_CUDA(cudaMemcpyAsync(&h_inv, inv_s2, 4, cudaMemcpyDeviceToHost, st));
  SB2_CUDA(cudaMemcpyAsync(h_bits, maxnorm, 8, cudaMemcpyDeviceToHost, st));
  SB2_CUDA(cudaStreamSynchronize(st));
  float r2, dmax;
  memcpy(&r2, &h_bits[0], 4);
  memcpy(&dmax, &h_bits[1], 4);
  double cq, cn;
  knn_tc2_error_coefs(sh, &cq, &cn);
  h_meta[0] = h_inv; h_meta[1] = r2; h_meta[2] = terms == 1 ? dmax : 0.0; h_meta[3] = cq; h_meta[4] = cn; h_meta[5] = list_m;
  return SB2_OK;
}
