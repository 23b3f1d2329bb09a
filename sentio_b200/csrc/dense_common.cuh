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

// ---- approximate -> exact hand-off (DESIGN.md "K1: exactness") --------------------------------------------------------
// Both scans rank rows by an APPROXIMATE cosine a(x) (fp32 accumulation; the tcgen05 scan additionally rounds the
// normalised query to fp16).  With |a(x) - cos(x)| <= eps for every row, the exact top-k is contained in
//     W = { x : a(x) >= a_k - 2*eps },   a_k = the k-th largest approximate score
// (the k best approximate rows have cos >= a_k - eps, so the k-th best exact cosine is >= a_k - eps, and a row with
// cos >= a_k - eps has a >= a_k - 2*eps).  Every member of W is re-scored in fp64; W has no fixed size.  eps == 0
// (the all-zero query: every product is exactly 0) degenerates to "the k largest composite keys".
// A query whose window does not fit the shared-memory winner buffer -- or, in the CUDA-core scan, reaches the end of a
// full per-CTA list -- raises its fallback flag and is answered by dense_exact_fallback_kernel (brute force in fp64).

// eps of the CUDA-core scan: fp32 FMA chains over d_pad products of |x||q| <= 1 (normalised operands): gamma_d <= d*2^-24;
// doubled, plus 2^-19 for the fp32 normalisation of the query, the fp32 inverse row norm and the final product.
__host__ __device__ __forceinline__ float dense_eps_fp32(int d_pad) { return (float)d_pad * 1.1920929e-7f + 1.9073486e-6f; }
// accumulation part of the tcgen05 scan's eps: the tensor core's fp32 accumulator may truncate (<= 2 ulp per step)
__host__ __device__ __forceinline__ float dense_eps_mma_acc(int d_pad) { return (float)d_pad * 2.3841858e-7f + 1.9073486e-6f; }

// lower edge of the window as a composite key (keys >= it are members)
__device__ __forceinline__ unsigned long long window_lo_key(unsigned long long kth_lb_key, float eps) {
  if (!(eps > 0.f)) return kth_lb_key;
  const float lo = __fsub_rd(key32_score(kth_lb_key), __fmul_ru(2.f, eps));
  return (unsigned long long)f32_orderable(lo) << 32;
}

// Exact fp64 cosine of stored row `idx` with the fp32 query q (norm qn), computed by one full warp; every lane returns it.
__device__ __forceinline__ double exact_cosine_warp(const __half* rows, uint32_t idx, const float* q, int d_pad, int nch,
                                                    double qn, int lane) {
  const uint4* row = reinterpret_cast<const uint4*>(rows + (size_t)idx * d_pad);
  double dot = 0.0, xx = 0.0;
  for (int ch = lane; ch < nch; ch += 32) {
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
  return den > 0.0 ? dot / den : 0.0;
}

// ||q|| in fp64 (whole CTA; result broadcast through *qq_s)
__device__ __forceinline__ double query_norm_cta(const float* q, int d_pad, double* qq_s) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 0) {
    double s = 0.0;
    for (int i = lane; i < d_pad; i += 32) {
      const double v = (double)q[i];
      s += v * v;
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) *qq_s = s;
  }
  __syncthreads();
  return sqrt(*qq_s);
}

// (exact key desc, row asc) bitonic sort of P = 2^m (key, row) pairs in shared memory; (0, *) = empty sorts last
__device__ __forceinline__ void sort_exact_pairs(unsigned long long* ek, uint32_t* ei, int P, int tid, int nt) {
  for (int kk = 2; kk <= P; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P; i += nt) {
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
}

// first min(k, P) sorted pairs -> this query's output rows
__device__ __forceinline__ void emit_exact_pairs(const unsigned long long* ek, const uint32_t* ei, int P,
                                                 const RescoreArgs& p) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < p.k; i += nt) {
    const bool valid = (i < P) && ek[i] != 0ull;
    p.out_ids[i] = valid ? p.id_base + (int64_t)ei[i] : -1;
    p.out_scores[i] = valid ? orderable_f64(ek[i]) : 0.0;
  }
  if (tid == 0) {
    int lo = 0, hi = min(p.k, P);  // valid entries are a prefix (empty keys sort last)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (ek[mid] != 0ull) lo = mid + 1; else hi = mid;
    }
    p.out_count[0] = lo;
  }
}

// Exact fp64 re-score of the window members sel[0..nsel) (composite keys) against the STORED fp16 rows and the fp32
// query, final order (score desc, row asc), emit k results.  Whole-CTA cooperative; ek/ei are P-entry shared-memory
// arrays (P = power of two >= nsel), qq_s a shared double.
__device__ __forceinline__ void rescore_and_emit(const unsigned long long* sel, int nsel, int P, unsigned long long* ek,
                                                 uint32_t* ei, double* qq_s_ptr, const RescoreArgs p) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x, nw = nt >> 5;
  const double qn = query_norm_cta(p.q, p.d_pad, qq_s_ptr);
  for (int c = warp; c < P; c += nw) {
    const unsigned long long key = c < nsel ? sel[c] : 0ull;
    unsigned long long okey = 0ull;
    uint32_t idx = 0xffffffffu;
    if (key != 0ull) {
      idx = key32_idx(key);
      okey = f64_orderable(exact_cosine_warp(p.rows, idx, p.q, p.d_pad, p.ch, qn, lane));
      if (okey == 0ull) okey = 1ull;  // keep 0 reserved for "empty"
    }
    if (lane == 0) {
      ek[c] = okey;
      ei[c] = idx;
    }
  }
  __syncthreads();
  sort_exact_pairs(ek, ei, P, tid, nt);
  emit_exact_pairs(ek, ei, P, p);
}
