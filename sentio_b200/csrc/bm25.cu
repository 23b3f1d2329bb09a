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

// ------------------------------------------------------------------------------------------------ top-k select v2
// sample threshold -> collect -> radix select.  (1) per query, the k-th best positive score among the first 8192 docs is a
// safe lower bound of the global k-th best; (2) one streaming pass collects every doc with score > 0 and >= bound
// (~ N * k / 8192 docs) into a per-query buffer sized for the worst case; (3) one CTA per query finds the k-th largest
// score by MSB-first radix select, breaks exact ties by ascending doc index with a second radix select, sorts the k
// winners.  Exact for any score distribution (an unrepresentative sample only enlarges step 2's output).
constexpr int kBmSample = 8192;
constexpr int kBmStage = 12288;  // (key, idx) pairs staged in shared memory by the final kernel (144 KB)
constexpr unsigned long long kPosZero = 0x8000000000000000ull;  // f64_orderable(+0.0)

// K-th largest value among keys[0..n) (shared memory, every thread of the CTA participates); n >= K >= 1.
__device__ unsigned long long block_kth_largest(const unsigned long long* keys, int n, int K, int* hist, int* scal) {
  const int tid = threadIdx.x, nt = blockDim.x;
  unsigned long long prefix = 0ull, mask = 0ull;
  int need = K;
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int i = tid; i < 256; i += nt) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
      const unsigned long long key = keys[i];
      if ((key & mask) == prefix) {  // warp-aggregated: equal digits elect one lane (concentrated digits serialise)
        const int dgt = (int)((key >> shift) & 0xffull);
        const unsigned grp = __match_any_sync(__activemask(), dgt);
        if ((int)(threadIdx.x & 31) == __ffs(grp) - 1) atomicAdd(&hist[dgt], __popc(grp));
      }
    }
    __syncthreads();
    if (tid == 0) {
      int cum = 0, sel = 0;
      for (int b = 255; b >= 0; --b) {
        const int c = hist[b];
        if (cum + c >= need) {
          sel = b;
          break;
        }
        cum += c;
      }
      scal[0] = sel;
      scal[1] = need - cum;
    }
    __syncthreads();
    prefix |= (unsigned long long)scal[0] << shift;
    mask |= 0xffull << shift;
    need = scal[1];
    __syncthreads();
  }
  return prefix;
}

__global__ void __launch_bounds__(1024, 1) bm25_sample_thr_kernel(const double* __restrict__ acc, int64_t n, int k,
                                                                  unsigned long long* __restrict__ thr,
                                                                  int32_t* __restrict__ cnt) {
  extern __shared__ __align__(16) uint8_t ssm2[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(ssm2);
  __shared__ int hist[256];
  __shared__ int scal[2];
  __shared__ int npos;
  const int qi = blockIdx.x, tid = threadIdx.x;
  const int ns = (int)min((int64_t)kBmSample, n);
  if (tid == 0) npos = 0;
  __syncthreads();
  int local = 0;
  for (int i = tid; i < ns; i += blockDim.x) {
    const double s = acc[(size_t)qi * n + i];
    const unsigned long long key = (s > 0.0) ? f64_orderable(s) : 0ull;
    keys[i] = key;
    local += key != 0ull;
  }
  atomicAdd(&npos, local);
  __syncthreads();
  unsigned long long t = kPosZero + 1ull;  // "every positive score"
  if (npos >= k) t = block_kth_largest(keys, ns, k, hist, scal);
  if (tid == 0) {
    thr[qi] = t;
    cnt[qi] = 0;
  }
}

__global__ void __launch_bounds__(256) bm25_collect_kernel(const double* __restrict__ acc, int64_t n,
                                                           const unsigned long long* __restrict__ thr,
                                                           int32_t* __restrict__ cnt,
                                                           unsigned long long* __restrict__ ckey,
                                                           uint32_t* __restrict__ cidx) {
  const int qi = blockIdx.y, lane = threadIdx.x & 31;
  const unsigned long long t = thr[qi];
  const double* a = acc + (size_t)qi * n;
  unsigned long long* ok = ckey + (size_t)qi * n;
  uint32_t* oi = cidx + (size_t)qi * n;
  const int64_t base = (int64_t)blockIdx.x * (256 * 8);
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int64_t i = base + (int64_t)u * 256 + threadIdx.x;
    unsigned long long key = 0ull;
    if (i < n) key = f64_orderable(a[i]);
    const bool pass = i < n && key > kPosZero && key >= t && key != 0xffffffffffffffffull;
    const unsigned m = __ballot_sync(0xffffffffu, pass);
    if (m) {
      int pos = 0;
      if (lane == 0) pos = atomicAdd(&cnt[qi], __popc(m));
      pos = __shfl_sync(0xffffffffu, pos, 0);
      if (pass) {
        const int at = pos + __popc(m & ((1u << lane) - 1u));
        ok[at] = key;
        oi[at] = (uint32_t)i;
      }
    }
  }
}

__global__ void __launch_bounds__(1024, 1) bm25_final_select_kernel(const unsigned long long* __restrict__ ckey,
                                                                    const uint32_t* __restrict__ cidx,
                                                                    const int32_t* __restrict__ cnt, int64_t n_docs,
                                                                    int k, int kpow2, int64_t id_base,
                                                                    int64_t* __restrict__ out_ids,
                                                                    double* __restrict__ out_scores,
                                                                    int32_t* __restrict__ out_counts) {
  extern __shared__ __align__(16) uint8_t fsm2[];
  unsigned long long* skey = reinterpret_cast<unsigned long long*>(fsm2);   // [kBmStage]
  uint32_t* sidx = reinterpret_cast<uint32_t*>(skey + kBmStage);            // [kBmStage]
  unsigned long long* wkey = reinterpret_cast<unsigned long long*>(sidx + kBmStage);  // [kpow2]
  uint32_t* widx = reinterpret_cast<uint32_t*>(wkey + kpow2);                           // [kpow2]
  __shared__ int hist[256];
  __shared__ int scal[4];
  const int qi = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int n = cnt[qi];
  const unsigned long long* gk = ckey + (size_t)qi * n_docs;
  const uint32_t* gi = cidx + (size_t)qi * n_docs;
  const bool staged = n <= kBmStage;
  if (staged) {
    for (int i = tid; i < n; i += nt) {
      skey[i] = gk[i];
      sidx[i] = gi[i];
    }
  }
  const unsigned long long* K_ = staged ? skey : gk;
  const uint32_t* I_ = staged ? sidx : gi;
  __syncthreads();
  unsigned long long T = 0ull;        // k-th largest score key; winners: key > T, or key == T with idx <= Icut
  uint32_t Icut = 0xffffffffu;
  if (n > k) {
    // (a) k-th largest key
    unsigned long long prefix = 0ull, mask = 0ull;
    int need = k;
    for (int shift = 56; shift >= 0; shift -= 8) {
      for (int i = tid; i < 256; i += nt) hist[i] = 0;
      __syncthreads();
      for (int i = tid; i < n; i += nt) {
        const unsigned long long key = K_[i];
        if ((key & mask) == prefix) {
          const int dgt = (int)((key >> shift) & 0xffull);
          const unsigned grp = __match_any_sync(__activemask(), dgt);
          if ((int)(threadIdx.x & 31) == __ffs(grp) - 1) atomicAdd(&hist[dgt], __popc(grp));
        }
      }
      __syncthreads();
      if (tid == 0) {
        int cum = 0, sel = 0;
        for (int b = 255; b >= 0; --b) {
          const int c = hist[b];
          if (cum + c >= need) {
            sel = b;
            break;
          }
          cum += c;
        }
        scal[0] = sel;
        scal[1] = need - cum;
      }
      __syncthreads();
      prefix |= (unsigned long long)scal[0] << shift;
      mask |= 0xffull << shift;
      need = scal[1];
      __syncthreads();
    }
    T = prefix;
    const int m = need;  // how many docs with key == T belong to the top k (1 <= m <= #equal)
    // (b) m-th smallest doc index among key == T
    uint32_t ipre = 0u, imask = 0u;
    int ineed = m;
    for (int shift = 24; shift >= 0; shift -= 8) {
      for (int i = tid; i < 256; i += nt) hist[i] = 0;
      __syncthreads();
      for (int i = tid; i < n; i += nt) {
        if (K_[i] == T) {
          const uint32_t ix = I_[i];
          if ((ix & imask) == ipre) atomicAdd(&hist[(int)((ix >> shift) & 0xffu)], 1);
        }
      }
      __syncthreads();
      if (tid == 0) {
        int cum = 0, sel = 255;
        for (int b = 0; b < 256; ++b) {
          const int c = hist[b];
          if (cum + c >= ineed) {
            sel = b;
            break;
          }
          cum += c;
        }
        scal[0] = sel;
        scal[1] = ineed - cum;
      }
      __syncthreads();
      ipre |= (uint32_t)scal[0] << shift;
      imask |= 0xffu << shift;
      ineed = scal[1];
      __syncthreads();
    }
    Icut = ipre;
  }
  // gather the winners, then order them (score desc, doc asc)
  if (tid == 0) scal[2] = 0;
  for (int i = tid; i < kpow2; i += nt) {
    wkey[i] = 0ull;
    widx[i] = 0xffffffffu;
  }
  __syncthreads();
  for (int i = tid; i < n; i += nt) {
    const unsigned long long key = K_[i];
    const uint32_t ix = I_[i];
    if (key > T || (key == T && ix <= Icut)) {
      const int at = atomicAdd(&scal[2], 1);
      if (at < kpow2) {
        wkey[at] = key;
        widx[at] = ix;
      }
    }
  }
  __syncthreads();
  block_bitonic_sort_pairs<1024>(wkey, widx, kpow2, tid);
  const int nw = min(min(scal[2], k), kpow2);
  for (int i = tid; i < k; i += nt) {
    const bool valid = i < nw;
    out_ids[(size_t)qi * k + i] = valid ? id_base + (int64_t)widx[i] : -1;
    out_scores[(size_t)qi * k + i] = valid ? orderable_f64(wkey[i]) : 0.0;
  }
  if (tid == 0) out_counts[qi] = nw;
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
  const int kpow2 = std::max(32, pow2_at_least(k));
  SB_REQUIRE(kpow2 <= 1024, SB_ERR_UNSUPPORTED, "bm25: top_k %d too large (max 1024)", k);
  // select v2 scratch: thresholds + counters + worst-case (key, idx) buffers of the sub-batch
  if ((rc = ctx->misc2_dev.reserve((size_t)sbq * ix.n_docs * 8 + (size_t)sbq * 16 + 64))) return rc;
  if ((rc = ctx->misc3_dev.reserve((size_t)sbq * ix.n_docs * 4 + 64))) return rc;
  unsigned long long* ckey = ctx->misc2_dev.as<unsigned long long>();
  unsigned long long* thr = ckey + (size_t)sbq * ix.n_docs;
  int32_t* cnt = reinterpret_cast<int32_t*>(thr + sbq);
  uint32_t* cidx = ctx->misc3_dev.as<uint32_t>();
  const size_t samp_smem = (size_t)kBmSample * 8;
  const size_t fin_smem = (size_t)kBmStage * 12 + (size_t)kpow2 * 12 + 64;
  SB_CUDA(cudaFuncSetAttribute(bm25_sample_thr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)samp_smem));
  SB_CUDA(cudaFuncSetAttribute(bm25_final_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem));
  for (int b0 = 0; b0 < B; b0 += sbq) {
    const int nq = std::min(sbq, B - b0);
    double* acc = ctx->acc_dev.as<double>();
    double* scratch = acc + (size_t)sbq * ix.n_docs;  // Plus only; fixed offset so it is always all-zero on entry
    if ((rc = bm25_score_subbatch(ctx, q_terms_dev, q_off_dev + b0, nq, max_len, acc, scratch, st))) return rc;
    ProfScope ps(ctx, SB_PROF_BM25_SELECT, st, 3);
    bm25_sample_thr_kernel<<<nq, 1024, samp_smem, st>>>(acc, ix.n_docs, k, thr, cnt);
    SB_CUDA(cudaGetLastError());
    dim3 cg((unsigned)((ix.n_docs + 2047) / 2048), (unsigned)nq);
    bm25_collect_kernel<<<cg, 256, 0, st>>>(acc, ix.n_docs, thr, cnt, ckey, cidx);
    SB_CUDA(cudaGetLastError());
    bm25_final_select_kernel<<<nq, 1024, fin_smem, st>>>(ckey, cidx, cnt, ix.n_docs, k, kpow2, ix.id_base,
                                                         out_ids + (size_t)b0 * k, out_scores + (size_t)b0 * k,
                                                         out_counts + b0);
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
