// dense_mma.cuh -- interface of the tcgen05 batched-query dense scan (dense_mma.cu) and the pieces it shares with dense.cu.
#pragma once
#include "common.cuh"

// true when the batched tensor-core scan can serve this index / batch (B >= 16, d_pad % 64 == 0, corpus >= 8192 rows)
bool dense_mma_eligible(const sb_ctx* ctx, const DenseIndex& ix, int B);

// q_pad: [B][d_pad] fp32 device (zero padded).  Enqueues sampling passes + full passes + the exact stage for groups of
// <= 128 queries (cta_group::2 pair kernel) / <= 64 queries (single-CTA kernel).
int dense_mma_topk_enqueue(sb_ctx* ctx, DenseIndex& ix, const float* q_pad, int B, int k, int64_t* out_ids,
                           double* out_scores, int32_t* out_counts, cudaStream_t st);

// dense.cu: normalised fp32 queries / fp16 operand rows / eps / cleared fallback flags for `rows` >= B operand rows
int dense_prep_queries(sb_ctx* ctx, const DenseIndex& ix, const float* q_pad, int B, int rows, bool mma, float** qn_out,
                       __half* q16, float** eps_out, int32_t** fb_out, cudaStream_t st);
// dense.cu: brute-force fp64 answer for every query whose fallback flag is raised (one CTA per query, idle CTAs exit)
int dense_fallback_enqueue(sb_ctx* ctx, const DenseIndex& ix, const float* q_pad, int B, int k, const int32_t* fb,
                           int64_t* out_ids, double* out_scores, int32_t* out_counts, cudaStream_t st);
