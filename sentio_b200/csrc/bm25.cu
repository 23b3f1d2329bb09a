// bm25.cu -- K2: BM25 term-at-a-time scoring over term-major CSR postings + streaming top-k.
//
// Replaces rank_bm25 0.2.2 BM25Okapi/BM25Plus.get_scores followed by np.argsort / `score > 0`
// (reference call sites src/core/retrievers/sparse.py:177-198).
//
// Bit-exactness contract (fp64, no FMA contraction -- every operation below is an explicit __d*_rn intrinsic):
//   dnorm[d]  = k1 * ((1 - b) + (b * dl[d]) / avgdl)                      (load time)
//   ratio[p]  = (tf[p] * (k1 + 1)) / (tf[p] + dnorm[doc[p]])              (load time, query independent)
//   Okapi:  score[d] += idf[t] * ratio[p]            for every posting p of query term t, terms in QUERY ORDER
//   Plus :  score[d] += idf[t] * (delta + ratio or 0.0)   for EVERY doc d (rank_bm25 adds delta to all docs)
// A doc occurs at most once in a term's posting list, so one launch per query-term position needs no atomics and the
// per-doc addition order equals NumPy's `score += ...` loop order.
//
// Algorithmic bytes per query: sum_t df(t) * (4 B doc + 8 B ratio + 16 B accumulator RMW) + N * 8 B zero fill
// + N * 8 B top-k read; the accumulators of a sub-batch are sized to stay L2 resident.
#include <algorithm>
#include <string.h>

#include "common.cuh"

namespace {

constexpr int kScoreThreads = 256;
constexpr int kPostPerThread = 8;
constexpr int kChunk = kScoreThreads * kPostPerThread;  // postings per work chunk
constexpr int kSelThreads = 1024;
constexpr int kSelPerIter = 2 * kSelThreads;  // scores examined per CTA iteration

// ------------------------------------------------------------------------------------------------ load kernels
__global__ void bm25_dnorm_kernel(const int32_t* __restrict__ doc_len, int64_t n, double k1, double b, double omb,
                                  double avgdl, double* __restrict__ dnorm) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double t1 = __dmul_rn(b, (double)doc_len[i]);
  const double t2 = __ddiv_rn(t1, avgdl);
  const double t4 = __dadd_rn(omb, t2);
  dnorm[i] = __dmul_rn(k1, t4);
}

__global__ void bm25_ratio_kernel(const int32_t* __restrict__ post_doc, const uint16_t* __restrict__ post_tf,
                                  int64_t nnz, const double* __restrict__ dnorm, double k1p1,
                                  double* __restrict__ ratio) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  const double tf = (double)post_tf[i];
  const double num = __dmul_rn(tf, k1p1);
  const double den = __dadd_rn(tf, dnorm[post_doc[i]]);
  ratio[i] = __ddiv_rn(num, den);
}

// ------------------------------------------------------------------------------------------------ plan kernel
// For the sub-batch of nq queries: for every term position j < max_len, the exclusive prefix of work chunks per query.
// chunk_prefix[j * (nq + 1) + b]; term id < 0 or idf == 0 -> no work (adds exactly +-0.0 in the reference).
__global__ void bm25_plan_kernel(const int32_t* __restrict__ q_terms, const int32_t* __restrict__ q_off, int nq,
                                 int max_len, const int64_t* __restrict__ indptr, const double* __restrict__ idf,
                                 int64_t n_terms, int32_t* __restrict__ chunk_prefix) {
  const int j = blockIdx.x;
  if (j >= max_len || threadIdx.x != 0) return;
  int32_t run = 0;
  int32_t* out = chunk_prefix + (size_t)j * (nq + 1);
  for (int b = 0; b < nq; ++b) {
    out[b] = run;
    const int len = q_off[b + 1] - q_off[b];
    if (j < len) {
      const int t = q_terms[q_off[b] + j];
      if (t >= 0 && t < n_terms && idf[t] != 0.0) {
        const int64_t df = indptr[t + 1] - indptr[t];
        run += (int32_t)((df + kChunk - 1) / kChunk);
      }
    }
  }
  out[nq] = run;
}

// ------------------------------------------------------------------------------------------------ scoring kernels
struct ScoreParams {
  const int32_t* q_terms;
  const int32_t* q_off;  // offsets of THIS sub-batch (q_off[0] may be > 0)
  int nq;
  int j;                 // term position handled by this launch
  const int32_t* chunk_prefix;  // [nq + 1] for this j
  const int64_t* indptr;
  const int32_t* post_doc;
  const double* ratio;
  const double* idf;
  double* acc;   // [nq][n_docs]   (Okapi: accumulators; Plus: per-term ratio scratch)
  int64_t n_docs;
};

// Okapi: acc[b][doc] += idf * ratio.   Plus (scatter phase): scratch[b][doc] = ratio.
template <bool PLUS>
__global__ void __launch_bounds__(kScoreThreads) bm25_score_kernel(const ScoreParams p) {
  __shared__ int s_pref[64 + 1];
  const int nq = p.nq;
  for (int i = threadIdx.x; i <= nq; i += blockDim.x) s_pref[i] = p.chunk_prefix[i];
  __syncthreads();
  const int total = s_pref[nq];
  for (int c = blockIdx.x; c < total; c += gridDim.x) {
    int b = 0;
    while (b + 1 < nq && s_pref[b + 1] <= c) ++b;  // nq <= 64: linear scan
    const int t = p.q_terms[p.q_off[b] + p.j];
    const int64_t lo = p.indptr[t], hi = p.indptr[t + 1];
    const double idf = p.idf[t];
    const int64_t base = lo + (int64_t)(c - s_pref[b]) * kChunk;
    double* acc = p.acc + (size_t)b * p.n_docs;
#pragma unroll
    for (int it = 0; it < kPostPerThread; ++it) {
      const int64_t pidx = base + (int64_t)it * kScoreThreads + threadIdx.x;
      if (pidx < hi) {
        const int32_t doc = __ldg(p.post_doc + pidx);
        const double r = __ldg(p.ratio + pidx);
        if (PLUS) {
          acc[doc] = r;
        } else {
          acc[doc] = __dadd_rn(acc[doc], __dmul_rn(idf, r));
        }
      }
    }
  }
}

// Plus (dense phase): for every query whose term j is scored: acc[b][d] += idf * (delta + scratch[b][d]); scratch = 0.
__global__ void bm25_plus_dense_kernel(const int32_t* q_terms, const int32_t* q_off, int nq, int j,
                                       const double* __restrict__ idf, int64_t n_terms, double delta,
                                       double* __restrict__ acc, double* __restrict__ scratch, int64_t n_docs) {
  const int b = blockIdx.y;
  const int len = q_off[b + 1] - q_off[b];
  if (j >= len) return;
  const int t = q_terms[q_off[b] + j];
  if (t < 0 || t >= n_terms) return;
  const double w = idf[t];
  if (w == 0.0) return;
  double* a = acc + (size_t)b * n_docs;
  double* s = scratch + (size_t)b * n_docs;
  for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < n_docs; d += (int64_t)gridDim.x * blockDim.x) {
    const double r = s[d];
    a[d] = __dadd_rn(a[d], __dmul_rn(w, __dadd_rn(delta, r)));
    if (r != 0.0) s[d] = 0.0;
  }
}

// ------------------------------------------------------------------------------------------------ top-k select
// Pair ordering: larger key first, ties -> smaller idx first.  key 0 == empty.
__device__ __forceinline__ bool pair_before(unsigned long long ka, uint32_t ia, unsigned long long kb, uint32_t ib) {
  return (ka > kb) || (ka == kb && ia < ib);
}

template <int NT>
__device__ __forceinline__ void block_bitonic_sort_pairs(unsigned long long* key, uint32_t* idx, int len, int tid) {
  for (int k = 2; k <= len; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < len; i += NT) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = key[i], b = key[ixj];
          const uint32_t ia = idx[i], ib = idx[ixj];
          const bool a_first = pair_before(a, ia, b, ib);
          const bool desc = (i & k) == 0;
          if ((desc ? !a_first : a_first) && !(a == b && ia == ib)) {
            key[i] = b; key[ixj] = a;
            idx[i] = ib; idx[ixj] = ia;
          }
        }
      }
      __syncthreads();
    }
  }
}

struct SelectParams {
  const double* acc;  // [nq][n]
  int64_t n;
  int cq;             // slices (CTAs) per query
  int kprime;         // power of two >= k
  int cap;            // smem candidate capacity (power of two, >= kprime + kSelPerIter)
  unsigned long long* list_key;  // [nq][cq][kprime]
  uint32_t* list_idx;            // [nq][cq][kprime]
};

// Each CTA streams one contiguous slice of one query's scores in ascending doc order and keeps the K' best
// (score desc, doc asc) among scores > 0.  Strict '>' against the running K'-th best is exact: a later doc that ties
// with the K'-th best has a larger index and loses the tie-break anyway.
__global__ void __launch_bounds__(kSelThreads, 1) bm25_select_kernel(const SelectParams p) {
  extern __shared__ __align__(16) uint8_t ssm[];
  unsigned long long* bkey = reinterpret_cast<unsigned long long*>(ssm);
  uint32_t* bidx = reinterpret_cast<uint32_t*>(bkey + p.cap);
  __shared__ int s_cnt;
  __shared__ unsigned long long s_thr;  // orderable key of the K'-th best so far (0 = none)
  const int tid = threadIdx.x;
  const int slice = blockIdx.x, qi = blockIdx.y;
  const int64_t per = (p.n + p.cq - 1) / p.cq;
  const int64_t lo = (int64_t)slice * per, hi = min(p.n, lo + per);
  const double* acc = p.acc + (size_t)qi * p.n;
  if (tid == 0) {
    s_cnt = 0;
    s_thr = f64_orderable(0.0);  // only scores strictly greater than +0.0 qualify
  }
  __syncthreads();
  const int trigger = p.cap - kSelPerIter;
  for (int64_t base = lo; base < hi; base += kSelPerIter) {
    const unsigned long long thr = s_thr;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t i = base + (int64_t)u * kSelThreads + tid;
      if (i < hi) {
        const double s = acc[i];
        const unsigned long long ok = f64_orderable(s);
        if (ok > thr && s == s) {
          const int pos = atomicAdd(&s_cnt, 1);
          bkey[pos] = ok;
          bidx[pos] = (uint32_t)i;
        }
      }
    }
    __syncthreads();
    const int c = s_cnt;
    if (c > trigger) {
      for (int z = c + tid; z < p.cap; z += kSelThreads) {
        bkey[z] = 0ull;
        bidx[z] = 0xffffffffu;
      }
      __syncthreads();
      block_bitonic_sort_pairs<kSelThreads>(bkey, bidx, p.cap, tid);
      if (tid == 0) {
        s_cnt = min(c, p.kprime);
        if (c >= p.kprime) s_thr = bkey[p.kprime - 1];
      }
      __syncthreads();
    }
  }
  __syncthreads();
  const int c = s_cnt;
  for (int z = c + tid; z < p.cap; z += kSelThreads) {
    bkey[z] = 0ull;
    bidx[z] = 0xffffffffu;
  }
  __syncthreads();
  block_bitonic_sort_pairs<kSelThreads>(bkey, bidx, p.cap, tid);
  unsigned long long* ok = p.list_key + ((size_t)qi * p.cq + slice) * p.kprime;
  uint32_t* oi = p.list_idx + ((size_t)qi * p.cq + slice) * p.kprime;
  for (int z = tid; z < p.kprime; z += kSelThreads) {
    ok[z] = bkey[z];
    oi[z] = bidx[z];
  }
}

// One CTA per query: sort the cq * K' surviving pairs, emit the first k (ids = id_base + doc, fp64 scores, count).
__global__ void __launch_bounds__(kSelThreads, 1) bm25_final_kernel(const unsigned long long* list_key,
                                                                    const uint32_t* list_idx, int cq, int kprime,
                                                                    int len_pow2, int k, int64_t id_base,
                                                                    int64_t* out_ids, double* out_scores,
                                                                    int32_t* out_counts) {
  extern __shared__ __align__(16) uint8_t fsm[];
  unsigned long long* key = reinterpret_cast<unsigned long long*>(fsm);
  uint32_t* idx = reinterpret_cast<uint32_t*>(key + len_pow2);
  const int tid = threadIdx.x, qi = blockIdx.x;
  const int total = cq * kprime;
  for (int i = tid; i < len_pow2; i += kSelThreads) {
    key[i] = i < total ? list_key[(size_t)qi * total + i] : 0ull;
    idx[i] = i < total ? list_idx[(size_t)qi * total + i] : 0xffffffffu;
  }
  __syncthreads();
  if (cq > 1) block_bitonic_sort_pairs<kSelThreads>(key, idx, len_pow2, tid);
  for (int i = tid; i < k; i += kSelThreads) {
    const bool valid = i < len_pow2 && key[i] != 0ull;
    out_ids[(size_t)qi * k + i] = valid ? id_base + (int64_t)idx[i] : -1;
    out_scores[(size_t)qi * k + i] = valid ? orderable_f64(key[i]) : 0.0;
  }
  if (tid == 0) {
    int lo = 0, hi = min(k, len_pow2);
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (key[mid] != 0ull) lo = mid + 1; else hi = mid;
    }
    out_counts[qi] = lo;
  }
}

int pow2_at_least(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// Score queries [0, nq) of a sub-batch into ctx->acc_dev ([nq][n_docs]); q_off points at the sub-batch's offsets.
int bm25_score_subbatch(sb_ctx* ctx, const int32_t* q_terms_dev, const int32_t* q_off_dev, int nq, int max_len,
                        double* acc, double* scratch, cudaStream_t st) {
  Bm25Index& ix = ctx->bm25;
  const size_t acc_bytes = (size_t)nq * ix.n_docs * sizeof(double);
  SB_CUDA(cudaMemsetAsync(acc, 0, acc_bytes, st));
  if (max_len <= 0) return SB_OK;
  int rc = ctx->misc_dev.reserve((size_t)max_len * (nq + 1) * sizeof(int32_t));
  if (rc) return rc;
  int32_t* chunk_prefix = ctx->misc_dev.as<int32_t>();
  ctx->launches += 1;
  bm25_plan_kernel<<<max_len, 32, 0, st>>>(q_terms_dev, q_off_dev, nq, max_len, ix.indptr, ix.idf, ix.n_terms,
                                           chunk_prefix);
  SB_CUDA(cudaGetLastError());
  const int grid = ctx->num_sms * 8;
  for (int j = 0; j < max_len; ++j) {
    ScoreParams sp;
    sp.q_terms = q_terms_dev;
    sp.q_off = q_off_dev;
    sp.nq = nq;
    sp.j = j;
    sp.chunk_prefix = chunk_prefix + (size_t)j * (nq + 1);
    sp.indptr = ix.indptr;
    sp.post_doc = ix.post_doc;
    sp.ratio = ix.post_ratio;
    sp.idf = ix.idf;
    sp.n_docs = ix.n_docs;
    ProfScope ps(ctx, SB_PROF_BM25_SCORE, st, ix.variant == SB_BM25_PLUS ? 2 : 1);
    if (ix.variant == SB_BM25_PLUS) {
      sp.acc = scratch;
      bm25_score_kernel<true><<<grid, kScoreThreads, 0, st>>>(sp);
      SB_CUDA(cudaGetLastError());
      dim3 g((unsigned)std::min<int64_t>((ix.n_docs + 255) / 256, (int64_t)ctx->num_sms * 4), (unsigned)nq);
      bm25_plus_dense_kernel<<<g, 256, 0, st>>>(q_terms_dev, q_off_dev, nq, j, ix.idf, ix.n_terms, ix.delta, acc,
                                                scratch, ix.n_docs);
      SB_CUDA(cudaGetLastError());
    } else {
      sp.acc = acc;
      bm25_score_kernel<false><<<grid, kScoreThreads, 0, st>>>(sp);
      SB_CUDA(cudaGetLastError());
    }
  }
  return SB_OK;
}

int bm25_subbatch_size(sb_ctx* ctx, int B) {
  const Bm25Index& ix = ctx->bm25;
  // keep the fp64 accumulators of a sub-batch L2 resident (~96 MB of the 126 MB L2)
  const size_t per_q = (size_t)ix.n_docs * sizeof(double) * (ix.variant == SB_BM25_PLUS ? 2 : 1);
  int64_t sbq = (int64_t)((96ull << 20) / (per_q ? per_q : 1));
  if (sbq < 1) sbq = 1;
  if (sbq > 64) sbq = 64;
  if (sbq > B) sbq = B;
  return (int)sbq;
}

int bm25_topk_enqueue(sb_ctx* ctx, const int32_t* q_terms_dev, const int32_t* q_off_dev, int B, int max_len, int k,
                      int64_t* out_ids, double* out_scores, int32_t* out_counts, cudaStream_t st) {
  Bm25Index& ix = ctx->bm25;
  const int sbq = bm25_subbatch_size(ctx, B);
  const size_t per_q = (size_t)ix.n_docs * sizeof(double) * (ix.variant == SB_BM25_PLUS ? 2 : 1);
  int rc = ctx->acc_dev.reserve(per_q * sbq);
  if (rc) return rc;
  if (ix.variant == SB_BM25_PLUS)
    SB_CUDA(cudaMemsetAsync(ctx->acc_dev.p, 0, per_q * sbq, st));  // ratio scratch must start at zero
  const int kprime = std::max(32, pow2_at_least(k));
  SB_REQUIRE(kprime <= 1024, SB_ERR_UNSUPPORTED, "bm25: top_k %d too large (max 1024)", k);
  const int cap = std::max(2 * kSelPerIter, pow2_at_least(kprime + kSelPerIter));
  int cq = (2 * ctx->num_sms + sbq - 1) / sbq;
  cq = std::max(1, std::min(cq, 16));
  while (cq > 1 && (int64_t)cq * kSelPerIter > ix.n_docs) cq >>= 1;
  while (cq * kprime > 8192) cq >>= 1;
  const int len_pow2 = pow2_at_least(cq * kprime);
  if ((rc = ctx->misc2_dev.reserve((size_t)sbq * cq * kprime * 8))) return rc;
  if ((rc = ctx->misc3_dev.reserve((size_t)sbq * cq * kprime * 4))) return rc;
  const size_t sel_smem = (size_t)cap * 12;
  const size_t fin_smem = (size_t)len_pow2 * 12;
  SB_CUDA(cudaFuncSetAttribute(bm25_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
  SB_CUDA(cudaFuncSetAttribute(bm25_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem));
  for (int b0 = 0; b0 < B; b0 += sbq) {
    const int nq = std::min(sbq, B - b0);
    double* acc = ctx->acc_dev.as<double>();
    double* scratch = acc + (size_t)sbq * ix.n_docs;  // Plus only; fixed offset so it is always all-zero on entry
    if ((rc = bm25_score_subbatch(ctx, q_terms_dev, q_off_dev + b0, nq, max_len, acc, scratch, st))) return rc;
    SelectParams sp;
    sp.acc = ctx->acc_dev.as<double>();
    sp.n = ix.n_docs;
    sp.cq = cq;
    sp.kprime = kprime;
    sp.cap = cap;
    sp.list_key = ctx->misc2_dev.as<unsigned long long>();
    sp.list_idx = ctx->misc3_dev.as<uint32_t>();
    ProfScope ps(ctx, SB_PROF_BM25_SELECT, st, 2);
    bm25_select_kernel<<<dim3(cq, nq), kSelThreads, sel_smem, st>>>(sp);
    SB_CUDA(cudaGetLastError());
    bm25_final_kernel<<<nq, kSelThreads, fin_smem, st>>>(sp.list_key, sp.list_idx, cq, kprime, len_pow2, k,
                                                         ix.id_base, out_ids + (size_t)b0 * k,
                                                         out_scores + (size_t)b0 * k, out_counts + b0);
    SB_CUDA(cudaGetLastError());
  }
  return SB_OK;
}

__global__ void bm25_fill_empty_kernel(int64_t* ids, double* sc, int32_t* cnt, int B, int k) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * k) {
    ids[i] = -1;
    sc[i] = 0.0;
  }
  if (i < B) cnt[i] = 0;
}

}  // namespace

extern "C" {

int sb_bm25_load(sb_ctx* ctx, const int64_t* indptr, const int32_t* post_doc, const uint16_t* post_tf,
                 int64_t n_terms, int64_t nnz, const int32_t* doc_len, int64_t n_docs, double avgdl,
                 const double* idf, int32_t variant, double k1, double b, double delta, int64_t id_base) {
  SB_REQUIRE(ctx != nullptr, SB_ERR_ARG, "sb_bm25_load: ctx is NULL");
  SB_REQUIRE(n_terms >= 0 && nnz >= 0 && n_docs >= 0, SB_ERR_ARG, "sb_bm25_load: negative size");
  SB_REQUIRE(n_docs < (1ll << 31), SB_ERR_ARG, "sb_bm25_load: a shard holds at most 2^31-1 docs");
  SB_REQUIRE(variant == SB_BM25_OKAPI || variant == SB_BM25_PLUS, SB_ERR_ARG, "sb_bm25_load: bad variant %d", variant);
  SB_REQUIRE(indptr && (nnz == 0 || (post_doc && post_tf)) && (n_docs == 0 || doc_len) && (n_terms == 0 || idf),
             SB_ERR_ARG, "sb_bm25_load: NULL buffer");
  SB_REQUIRE(indptr[0] == 0 && indptr[n_terms] == nnz, SB_ERR_ARG, "sb_bm25_load: indptr does not span nnz");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  Bm25Index& ix = ctx->bm25;
  SB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ix.indptr) cudaFree(ix.indptr);
  if (ix.post_doc) cudaFree(ix.post_doc);
  if (ix.post_ratio) cudaFree(ix.post_ratio);
  if (ix.dnorm) cudaFree(ix.dnorm);
  if (ix.idf) cudaFree(ix.idf);
  ix = Bm25Index();
  ix.n_docs = n_docs;
  ix.n_terms = n_terms;
  ix.nnz = nnz;
  ix.id_base = id_base;
  ix.variant = variant;
  ix.k1 = k1;
  ix.b = b;
  ix.delta = delta;
  ix.avgdl = avgdl;
  ix.h_indptr.assign(indptr, indptr + n_terms + 1);
  cudaStream_t st = ctx->stream;
  SB_CUDA(cudaMalloc(&ix.indptr, (size_t)(n_terms + 1) * 8));
  SB_CUDA(cudaMemcpyAsync(ix.indptr, indptr, (size_t)(n_terms + 1) * 8, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMalloc(&ix.idf, (size_t)std::max<int64_t>(n_terms, 1) * 8));
  if (n_terms) SB_CUDA(cudaMemcpyAsync(ix.idf, idf, (size_t)n_terms * 8, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMalloc(&ix.dnorm, (size_t)std::max<int64_t>(n_docs, 1) * 8));
  SB_CUDA(cudaMalloc(&ix.post_doc, (size_t)std::max<int64_t>(nnz, 1) * 4));
  // query-independent fp64 ratio tf*(k1+1)/(tf+dnorm[doc]) (8 B per posting) replaces the 2 B tf + 8 B dnorm gather
  SB_CUDA(cudaMalloc(&ix.post_ratio, (size_t)std::max<int64_t>(nnz, 1) * 8));
  if (n_docs) {
    int rc = ctx->misc_dev.reserve((size_t)n_docs * 4);
    if (rc) return rc;
    SB_CUDA(cudaMemcpyAsync(ctx->misc_dev.p, doc_len, (size_t)n_docs * 4, cudaMemcpyHostToDevice, st));
    const double omb = 1.0 - b;  // Python evaluates `1 - self.b` first (left-to-right)
    bm25_dnorm_kernel<<<(unsigned)((n_docs + 255) / 256), 256, 0, st>>>(ctx->misc_dev.as<int32_t>(), n_docs, k1, b,
                                                                        omb, avgdl, ix.dnorm);
    SB_CUDA(cudaGetLastError());
  }
  if (nnz) {
    SB_CUDA(cudaMemcpyAsync(ix.post_doc, post_doc, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
    int rc = ctx->misc2_dev.reserve((size_t)nnz * 2);
    if (rc) return rc;
    SB_CUDA(cudaMemcpyAsync(ctx->misc2_dev.p, post_tf, (size_t)nnz * 2, cudaMemcpyHostToDevice, st));
    const double k1p1 = k1 + 1.0;
    bm25_ratio_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(ix.post_doc, ctx->misc2_dev.as<uint16_t>(), nnz,
                                                                     ix.dnorm, k1p1, ix.post_ratio);
    SB_CUDA(cudaGetLastError());
  }
  SB_CUDA(cudaStreamSynchronize(st));
  return SB_OK;
}

int64_t sb_bm25_count(sb_ctx* ctx) { return ctx ? ctx->bm25.n_docs : -1; }

int sb_bm25_topk_dev(sb_ctx* ctx, const int32_t* q_terms_dev, const int32_t* q_off_dev, int32_t B, int32_t n_q_terms,
                     int32_t max_q_len, int32_t k, int64_t* out_ids_dev, double* out_scores_dev,
                     int32_t* out_counts_dev, void* stream) {
  SB_REQUIRE(ctx != nullptr, SB_ERR_ARG, "sb_bm25_topk_dev: ctx is NULL");
  SB_REQUIRE(B >= 0 && k > 0 && max_q_len >= 0 && n_q_terms >= 0, SB_ERR_ARG, "sb_bm25_topk_dev: bad sizes");
  if (B == 0) return SB_OK;
  SB_REQUIRE(q_off_dev && out_ids_dev && out_scores_dev && out_counts_dev, SB_ERR_ARG, "sb_bm25_topk_dev: NULL buffer");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  cudaStream_t st = pick_stream(ctx, stream);
  if (ctx->bm25.n_docs == 0 || ctx->bm25.indptr == nullptr) {
    bm25_fill_empty_kernel<<<(B * k + 255) / 256, 256, 0, st>>>(out_ids_dev, out_scores_dev, out_counts_dev, B, k);
    SB_CUDA(cudaGetLastError());
    return SB_OK;
  }
  return bm25_topk_enqueue(ctx, q_terms_dev, q_off_dev, B, max_q_len, k, out_ids_dev, out_scores_dev, out_counts_dev,
                           st);
}

int sb_bm25_topk(sb_ctx* ctx, const int32_t* q_terms, const int32_t* q_off, int32_t B, int32_t k, int64_t* out_ids,
                 double* out_scores, int32_t* out_counts) {
  SB_REQUIRE(ctx != nullptr, SB_ERR_ARG, "sb_bm25_topk: ctx is NULL");
  SB_REQUIRE(B >= 0 && k > 0, SB_ERR_ARG, "sb_bm25_topk: bad B=%d k=%d", B, k);
  if (B == 0) return SB_OK;
  SB_REQUIRE(q_off && out_ids && out_scores && out_counts, SB_ERR_ARG, "sb_bm25_topk: NULL buffer");
  const int n_terms_q = q_off[B];
  SB_REQUIRE(n_terms_q >= 0 && (n_terms_q == 0 || q_terms), SB_ERR_ARG, "sb_bm25_topk: bad query term buffers");
  int max_len = 0;
  for (int b = 0; b < B; ++b) {
    SB_REQUIRE(q_off[b + 1] >= q_off[b], SB_ERR_ARG, "sb_bm25_topk: q_off must be non-decreasing");
    max_len = std::max(max_len, q_off[b + 1] - q_off[b]);
  }
  std::unique_lock<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  cudaStream_t st = ctx->stream;
  if (ctx->bm25.n_docs == 0 || ctx->bm25.indptr == nullptr) {
    for (int i = 0; i < B * k; ++i) { out_ids[i] = -1; out_scores[i] = 0.0; }
    for (int i = 0; i < B; ++i) out_counts[i] = 0;
    return SB_OK;
  }
  int rc;
  const size_t tb = (size_t)std::max(n_terms_q, 1) * 4, ob = (size_t)(B + 1) * 4;
  if ((rc = ctx->pin_in.reserve(tb + ob))) return rc;
  if ((rc = ctx->q_dev.reserve(tb + ob))) return rc;
  uint8_t* pi = ctx->pin_in.as<uint8_t>();
  if (n_terms_q) memcpy(pi, q_terms, (size_t)n_terms_q * 4);
  memcpy(pi + tb, q_off, ob);
  SB_CUDA(cudaMemcpyAsync(ctx->q_dev.p, pi, tb + ob, cudaMemcpyHostToDevice, st));
  const int32_t* qt_dev = ctx->q_dev.as<int32_t>();
  const int32_t* qo_dev = reinterpret_cast<const int32_t*>(ctx->q_dev.as<uint8_t>() + tb);
  const size_t nid = (size_t)B * k;
  if ((rc = ctx->out_ids_dev.reserve(nid * 8))) return rc;
  if ((rc = ctx->out_sc_dev.reserve(nid * 8))) return rc;
  if ((rc = ctx->out_cnt_dev.reserve((size_t)B * 4))) return rc;
  if ((rc = bm25_topk_enqueue(ctx, qt_dev, qo_dev, B, max_len, k, ctx->out_ids_dev.as<int64_t>(),
                              ctx->out_sc_dev.as<double>(), ctx->out_cnt_dev.as<int32_t>(), st)))
    return rc;
  if ((rc = ctx->pin_out.reserve(nid * 16 + (size_t)B * 4))) return rc;
  uint8_t* po = ctx->pin_out.as<uint8_t>();
  SB_CUDA(cudaMemcpyAsync(po, ctx->out_ids_dev.p, nid * 8, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(po + nid * 8, ctx->out_sc_dev.p, nid * 8, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(po + nid * 16, ctx->out_cnt_dev.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  memcpy(out_ids, po, nid * 8);
  memcpy(out_scores, po + nid * 8, nid * 8);
  memcpy(out_counts, po + nid * 16, (size_t)B * 4);
  return SB_OK;
}

int sb_bm25_scores(sb_ctx* ctx, const int32_t* q_terms, int32_t n_q, double* out_scores) {
  SB_REQUIRE(ctx != nullptr, SB_ERR_ARG, "sb_bm25_scores: ctx is NULL");
  SB_REQUIRE(n_q >= 0 && (n_q == 0 || q_terms) && out_scores, SB_ERR_ARG, "sb_bm25_scores: bad arguments");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  Bm25Index& ix = ctx->bm25;
  SB_REQUIRE(ix.indptr != nullptr, SB_ERR_STATE, "sb_bm25_scores: no BM25 index loaded");
  if (ix.n_docs == 0) return SB_OK;
  cudaStream_t st = ctx->stream;
  int rc;
  const size_t tb = (size_t)std::max(n_q, 1) * 4;
  if ((rc = ctx->q_dev.reserve(tb + 8))) return rc;
  int32_t off[2] = {0, n_q};
  if (n_q) SB_CUDA(cudaMemcpyAsync(ctx->q_dev.p, q_terms, (size_t)n_q * 4, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(ctx->q_dev.as<uint8_t>() + tb, off, 8, cudaMemcpyHostToDevice, st));
  const size_t per_q = (size_t)ix.n_docs * sizeof(double) * (ix.variant == SB_BM25_PLUS ? 2 : 1);
  if ((rc = ctx->acc_dev.reserve(per_q))) return rc;
  if (ix.variant == SB_BM25_PLUS) SB_CUDA(cudaMemsetAsync(ctx->acc_dev.p, 0, per_q, st));
  if ((rc = bm25_score_subbatch(ctx, ctx->q_dev.as<int32_t>(),
                                reinterpret_cast<const int32_t*>(ctx->q_dev.as<uint8_t>() + tb), 1, n_q,
                                ctx->acc_dev.as<double>(), ctx->acc_dev.as<double>() + ix.n_docs, st)))
    return rc;
  SB_CUDA(cudaMemcpyAsync(out_scores, ctx->acc_dev.p, (size_t)ix.n_docs * 8, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  return SB_OK;
}

}  // extern "C"
