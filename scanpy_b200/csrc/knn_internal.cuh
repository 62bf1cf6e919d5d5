// knn_internal.cuh — interface between knn.cu (host entry, re-score, fallback) and knn_tc.cu (tensor-core pass 1)
#pragma once
#include "common.cuh"

bool knn_tc_supported(int d);
// proposals (unsorted, 32 per query; unused slots: score -inf, id -1) + the scale the scores carry
// (score_true = score * inv_s2) + the coefficient c of the error bound  eps = c * (R^2/2 + |q| R)
int32_t knn_tc_pass1(sb2_ctx* ctx, ScratchScope& scr, const float* d_x, int64_t n_points, int d,
                     const unsigned int* d_maxnorm_bits, int64_t q0, int64_t n_query, int list_m, float* cand_score,
                     int32_t* cand_idx, float* d_inv_s2, double* eps_coef, cudaEvent_t ev_after_prep,
                     double* issued_flops);
