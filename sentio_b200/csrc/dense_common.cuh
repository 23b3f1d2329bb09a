// dense_common.cuh -- pieces shared by the CUDA-core scan (dense.cu) and the tcgen05 batched scan (dense_mma.cu).
#pragma once
#include "common.cuh"

struct RescoreArgs {
  const __half* rows;
  const float* q;     // this query's fp32 vector [d_pad]
  int32_t d_pad;
  int32_t ch;         // 16-byte chunks per row
  int64_t id_base;
  int32_t k;
  int64_t* out_ids;     // [k] of this query
  double* out_scores;   // [k]
  int32_t* out_count;   // [1]
};

// Exact fp64 re-score of the K approximate survivors sel[0..K) (composite keys, 0 = empty) against the STORED fp16 rows
// and the fp32 query, final order (score desc, row asc), emit k results.  Whole-CTA cooperative; ek/ei are K-entry smem
// scratch arrays, qq_s a shared double.
__device__ __forceinline__ void rescore_and_emit(const unsigned long long* sel, int K, unsigned long long* ek, uint32_t* ei,
                                                 double* qq_s_ptr, const RescoreArgs p) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x, nw = nt >> 5;
  double& qq_s = *qq_s_ptr;
  const float* q = p.q;
  // (4) exact fp64 re-score of the K survivors against the stored fp16 rows
  if (warp == 0) {
    double s = 0.0;
    for (int i = lane; i < p.d_pad; i += 32) {
      const double v = (double)q[i];
      s += v * v;
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) qq_s = s;
  }
  __syncthreads();
  const double qn = sqrt(qq_s);
  for (int c = warp; c < K; c += nw) {
    const unsigned long long key = sel[c];
    unsigned long long okey = 0ull;
    uint32_t idx = 0xffffffffu;
    if (key != 0ull) {
      idx = key32_idx(key);
      const uint4* row = reinterpret_cast<const uint4*>(p.rows + (size_t)idx * p.d_pad);
      double dot = 0.0, xx = 0.0;
      for (int ch = lane; ch < p.ch; ch += 32) {
        const uint4 raw = __ldg(row + ch);
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
        const float4 qa = *reinterpret_cast<const float4*>(q + (size_t)ch * 8);
        const float4 qb = *reinterpret_cast<const float4*>(q + (size_t)ch * 8 + 4);
        const float qv[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 xf = __half22float2(h2[e]);
          const double x0 = (double)xf.x, x1 = (double)xf.y;
          dot += x0 * (double)qv[2 * e];
          dot += x1 * (double)qv[2 * e + 1];
          xx += x0 * x0;
          xx += x1 * x1;
        }
      }
      for (int o = 16; o; o >>= 1) {
        dot += __shfl_xor_sync(0xffffffffu, dot, o);
        xx += __shfl_xor_sync(0xffffffffu, xx, o);
      }
      const double den = qn * sqrt(xx);
      const double score = den > 0.0 ? dot / den : 0.0;
      okey = f64_orderable(score);
      if (okey == 0ull) okey = 1ull;  // keep 0 reserved for "empty"
    }
    if (lane == 0) {
      ek[c] = okey;
      ei[c] = idx;
    }
  }
  __syncthreads();
  // (5) final sort by (exact score desc, row index asc)
  for (int kk = 2; kk <= K; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < K; i += nt) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = ek[i], b = ek[ixj];
          const uint32_t ia = ei[i], ib = ei[ixj];
          const bool a_before_b = (a > b) || (a == b && ia < ib);
          const bool desc = (i & kk) == 0;
          if ((desc ? !a_before_b : a_before_b) && !(a == b && ia == ib)) {
            ek[i] = b; ek[ixj] = a;
            ei[i] = ib; ei[ixj] = ia;
          }
        }
      }
      __syncthreads();
    }
  }
  int64_t* oid = p.out_ids;
  double* osc = p.out_scores;
  for (int i = tid; i < p.k; i += nt) {
    const bool valid = (i < K) && ek[i] != 0ull;
    oid[i] = valid ? p.id_base + (int64_t)ei[i] : -1;
    osc[i] = valid ? orderable_f64(ek[i]) : 0.0;
  }
  if (tid == 0) {
    int lo = 0, hi = min(p.k, K);  // valid entries are a prefix (empty keys sort last)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (ek[mid] != 0ull) lo = mid + 1; else hi = mid;
    }
    p.out_count[0] = lo;
  }
}
