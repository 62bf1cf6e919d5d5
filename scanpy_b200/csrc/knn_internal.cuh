// knn_internal.cuh — interface between knn.cu (host entry, re-score, fallback) and knn_tc.cu (tensor-core sweeps)
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

// tile shape of one tensor-core sweep: `terms` = 1 (fp16 operands, fast first tier) or 3 (split precision, 22 bits)
struct KnnTcShape {
  int qh;       // 128-query tiles per CTA (2 or 1)
  int nsplit;   // slices the K axis of a candidate image is staged in (1, 2, 4)
  int nstage;   // depth of the candidate ring
  int kpad;     // padded K axis (terms*d + 3 rounded up to 16*nsplit)
  int terms;
  size_t smem;
};
bool knn_tc_supported(int d);
bool knn_tc_shape(const sb2_ctx* ctx, int d, int terms, KnnTcShape* out);
size_t knn_tc_image_halves(const KnnTcShape& sh, int64_t n_rows);
// operand images for rows [0, n_rows) of X (or, with d_gather, rows gather_base + d_gather[i]); Aimg or Bimg may be
// null; d_dnorm[i] / *d_dmax_bits (optional) receive |x_i - fp16(x_i)| and its maximum; d_inv_s2 receives 1/s^2 of the power-of-two scale the scores carry (score_true = score * inv_s2)
int32_t knn_tc_build_images(sb2_ctx* ctx, const KnnTcShape& sh, const float* d_x, int64_t n_rows, int d,
                            const unsigned int* d_maxnorm_bits, const int32_t* d_gather, int64_t gather_base,
                            __half* Aimg, __half* Bimg, float* d_inv_s2, float* d_dnorm = nullptr,
                            unsigned int* d_dmax_bits = nullptr);
// proposals (unsorted, list_m per query; unused slots: score -inf, id -1) for the n_query rows whose A images
// start at tile a_tile0; *issued_flops is incremented; estimate = false starts the lists cold (no sampled threshold:
// every row ends with its true top list_m, the certificate can only fail on exact ties)
int32_t knn_tc_sweep(sb2_ctx* ctx, const KnnTcShape& sh, const __half* Aimg, int64_t a_tile0, const __half* Bimg,
                     int64_t n_points, int64_t n_query, int list_m, float* cand_score, int32_t* cand_idx, double* issued_flops,
                     bool estimate = true);
void knn_tc_error_coefs(const KnnTcShape& sh, double* c_q, double* c_n);
