// dense_mma.cuh -- interface of the tcgen05 batched-query dense scan (dense_mma.cu).
#pragma once
#include "common.cuh"

// true when the batched tensor-core scan can serve this index / batch (B >= 16, d_pad % 64 == 0, corpus >= 8192 rows)
bool dense_mma_eligible(const sb_ctx* ctx, const DenseIndex& ix, int B);

// q_pad: [B][d_pad] fp32 device (zero padded).  Enqueues sampling pass + full pass + exact stage per block of <= 64 queries.
int dense_mma_topk_enqueue(sb_ctx* ctx, DenseIndex& ix, const float* q_pad, int B, int k, int kprime, int64_t* out_ids,
                           double* out_scores, int32_t* out_counts, cudaStream_t st);
