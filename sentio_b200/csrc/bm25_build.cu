// bm25_build.cu -- GPU construction of the BM25 index.
//
// Replaces the corpus pass of BM25Retriever.index (reference src/core/retrievers/sparse.py:70-100: tokenise every doc,
// hand the token lists to rank_bm25, which builds per-doc frequency dicts, df and doc_len in Python -- about a minute
// and ~10 GB at 1 M docs) with a sort-based build on the device:
//
//   token stream (doc i = flat[off[i] : off[i+1]])
//     -> keys (token << 32 | position), values doc            one pass over the docs
//     -> cub::DeviceRadixSort::SortPairs                       token-major, position-minor => docs ascend inside a token
//     -> head flags + exclusive scans                          token runs (= vocabulary) and (token, doc) runs (= postings)
//     -> term ids in FIRST-OCCURRENCE order (sort of the runs by their first position): rank_bm25 sums idf over its
//        insertion-ordered dict, and the epsilon floor of negative idfs depends on that float sum, so the order is part
//        of the bit-exactness contract (sentio_b200/index.py)
//     -> term-major CSR: indptr (scan of df), post_doc, tf = length of each (token, doc) run
//
// The integer work ends here.  idf needs libm's log bit for bit (math.log in the reference), so df[V] goes back to the
// host (V values), the host computes the idf table exactly like rank_bm25 and sb_bm25_build_finish installs the index
// (dnorm / ratio kernels of bm25.cu).  Nothing of size T or nnz ever exists on the host.
#include <cub/cub.cuh>

#include <algorithm>

#include "common.cuh"

int bm25_install_device_csr(sb_ctx* ctx, int64_t* indptr_dev, int32_t* post_doc_dev, const uint16_t* tf_dev,
                            const int32_t* doc_len_dev, int64_t n_terms, int64_t nnz, int64_t n_docs, double avgdl,
                            const double* idf_host, int32_t variant, double k1, double b, double delta, int64_t id_base,
                            cudaStream_t st);  // bm25.cu

struct Bm25Build {
  int64_t T = 0, n_docs = 0, V = 0, nnz = 0;
  int64_t* indptr = nullptr;    // [V+1]
  int32_t* post_doc = nullptr;  // [nnz]
  uint16_t* tf = nullptr;       // [nnz]
  int32_t* doc_len = nullptr;   // [n_docs]
  int64_t* df = nullptr;        // [V]  (term-id order)
  int32_t* term_token = nullptr;  // [V] raw token of every term id
};

void bm25_build_free(Bm25Build* b) {
  if (!b) return;
  if (b->indptr) cudaFree(b->indptr);
  if (b->post_doc) cudaFree(b->post_doc);
  if (b->tf) cudaFree(b->tf);
  if (b->doc_len) cudaFree(b->doc_len);
  if (b->df) cudaFree(b->df);
  if (b->term_token) cudaFree(b->term_token);
  delete b;
}

namespace {

// one warp per doc: keys / values of its tokens, its length; flags negative tokens
__global__ void build_keys_kernel(const int32_t* __restrict__ flat, const int64_t* __restrict__ off, int64_t n_docs,
                                  unsigned long long* __restrict__ keys, int32_t* __restrict__ vals,
                                  int32_t* __restrict__ doc_len, int32_t* __restrict__ err) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t d = warp0; d < n_docs; d += nwarps) {
    const int64_t lo = off[d], hi = off[d + 1];
    if (lane == 0) doc_len[d] = (int32_t)(hi - lo);
    for (int64_t p = lo + lane; p < hi; p += 32) {
      const int32_t t = flat[p];
      if (t < 0) atomicExch(err, 1);
      keys[p] = ((unsigned long long)(uint32_t)t << 32) | (unsigned long long)(uint32_t)p;
      vals[p] = (int32_t)d;
    }
  }
}

// head flags of token runs and of (token, doc) runs in the sorted stream
__global__ void head_flags_kernel(const unsigned long long* __restrict__ keys, const int32_t* __restrict__ docs, int64_t T,
                                  int32_t* __restrict__ run_head, int32_t* __restrict__ pair_head) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T) return;
  const uint32_t tok = (uint32_t)(keys[i] >> 32);
  const bool rh = i == 0 || tok != (uint32_t)(keys[i - 1] >> 32);
  const bool ph = rh || docs[i] != docs[i - 1];
  run_head[i] = rh;
  pair_head[i] = ph;
}

// per token run: first position (its sort key), raw token, first posting;  per posting: its start in the sorted stream
__global__ void scatter_heads_kernel(const unsigned long long* __restrict__ keys, const int32_t* __restrict__ run_head,
                                     const int32_t* __restrict__ pair_head, const int32_t* __restrict__ run_idx,
                                     const int32_t* __restrict__ pair_idx, int64_t T, uint32_t* __restrict__ run_first_pos,
                                     int32_t* __restrict__ run_token, int32_t* __restrict__ run_pair_start,
                                     int32_t* __restrict__ run_iota, int32_t* __restrict__ pair_pos) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T) return;
  if (pair_head[i]) pair_pos[pair_idx[i]] = (int32_t)i;
  if (run_head[i]) {
    const int u = run_idx[i];
    run_first_pos[u] = (uint32_t)(keys[i] & 0xffffffffull);
    run_token[u] = (int32_t)(keys[i] >> 32);
    run_pair_start[u] = pair_idx[i];
    run_iota[u] = u;
  }
}

// r = term id (rank by first occurrence), u = sorted_runs[r]: df, raw token, inverse permutation
__global__ void rank_runs_kernel(const int32_t* __restrict__ sorted_runs, const int32_t* __restrict__ run_pair_start,
                                 const int32_t* __restrict__ run_token, int64_t V, int64_t nnz,
                                 int32_t* __restrict__ rank_of_run, int64_t* __restrict__ df,
                                 int32_t* __restrict__ term_token) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= V) return;
  const int u = sorted_runs[r];
  rank_of_run[u] = (int32_t)r;
  const int64_t end = u + 1 < V ? run_pair_start[u + 1] : nnz;
  df[r] = end - run_pair_start[u];
  term_token[r] = run_token[u];
}

__global__ void scatter_postings_kernel(const int32_t* __restrict__ pair_pos, const int32_t* __restrict__ run_idx,
                                        const int32_t* __restrict__ run_head, const int32_t* __restrict__ docs, const int32_t* __restrict__ rank_of_run,
                                        const int32_t* __restrict__ run_pair_start, const int64_t* __restrict__ indptr,
                                        int64_t nnz, int64_t T, int32_t* __restrict__ post_doc,
                                        uint16_t* __restrict__ tf, int32_t* __restrict__ err) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  const int64_t i = pair_pos[p];
  const int64_t next = p + 1 < nnz ? (int64_t)pair_pos[p + 1] : T;
  const int u = run_idx[i] + run_head[i] - 1;  // exclusive scan + own flag - 1 = index of the run that contains i
  const int64_t dest = indptr[rank_of_run[u]] + (p - run_pair_start[u]);
  post_doc[dest] = docs[i];
  const int64_t f = next - i;
  if (f > 65535) atomicExch(err, 2);
  tf[dest] = (uint16_t)min(f, (int64_t)65535);
}

struct TempPool {
  std::vector<void*> ptrs;
  ~TempPool() {
    for (void* p : ptrs) cudaFree(p);
  }
  template <typename T>
  int alloc(T** out, size_t n) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) {
      sb_set_error("bm25 build: cudaMalloc(%zu) failed: %s", n * sizeof(T), cudaGetErrorString(e));
      return SB_ERR_CUDA;
    }
    ptrs.push_back(p);
    *out = reinterpret_cast<T*>(p);
    return SB_OK;
  }
};

inline unsigned blocks_for(int64_t n, int per = 256) { return (unsigned)((n + per - 1) / per); }

}  // namespace

extern "C" {

int sb_bm25_build_tokens(sb_ctx* ctx, const int32_t* flat_tokens, int64_t n_tokens, const int64_t* doc_off, int64_t n_docs,
                         int64_t* n_terms_out, int64_t* nnz_out) {
  SB_REQUIRE(ctx != nullptr && n_terms_out && nnz_out, SB_ERR_ARG, "sb_bm25_build_tokens: NULL argument");
  SB_REQUIRE(n_docs > 0 && doc_off != nullptr, SB_ERR_ARG, "sb_bm25_build_tokens: empty corpus");
  SB_REQUIRE(n_docs < (1ll << 31), SB_ERR_ARG, "sb_bm25_build_tokens: a shard holds at most 2^31-1 docs");
  SB_REQUIRE(doc_off[0] == 0 && doc_off[n_docs] == n_tokens, SB_ERR_ARG, "sb_bm25_build_tokens: doc_off does not span the stream");
  SB_REQUIRE(n_tokens > 0 && n_tokens < (1ll << 31) && flat_tokens != nullptr, SB_ERR_ARG,
             "sb_bm25_build_tokens: the token stream must hold 1 .. 2^31-1 tokens (got %lld)", (long long)n_tokens);
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  cudaStream_t st = ctx->stream;
  if (ctx->bm25_build) {
    bm25_build_free(ctx->bm25_build);
    ctx->bm25_build = nullptr;
  }
  const int64_t T = n_tokens;
  TempPool tmp;
  int rc;
  int32_t* flat_d;
  int64_t* off_d;
  unsigned long long *keys, *keys2;
  int32_t *vals, *vals2, *run_head, *pair_head, *run_idx, *pair_idx, *err_d;
  if ((rc = tmp.alloc(&flat_d, (size_t)T))) return rc;
  if ((rc = tmp.alloc(&off_d, (size_t)n_docs + 1))) return rc;
  if ((rc = tmp.alloc(&keys, (size_t)T))) return rc;
  if ((rc = tmp.alloc(&keys2, (size_t)T))) return rc;
  if ((rc = tmp.alloc(&vals, (size_t)T))) return rc;
  if ((rc = tmp.alloc(&vals2, (size_t)T))) return rc;
  if ((rc = tmp.alloc(&err_d, 1))) return rc;
  Bm25Build* B = new Bm25Build();
  struct Guard {
    Bm25Build* b;
    ~Guard() { if (b) bm25_build_free(b); }
  } guard{B};
  B->T = T;
  B->n_docs = n_docs;
  SB_CUDA(cudaMalloc(&B->doc_len, (size_t)n_docs * 4));
  SB_CUDA(cudaMemsetAsync(err_d, 0, 4, st));
  SB_CUDA(cudaMemcpyAsync(flat_d, flat_tokens, (size_t)T * 4, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(off_d, doc_off, (size_t)(n_docs + 1) * 8, cudaMemcpyHostToDevice, st));
  ctx->launches += 1;
  build_keys_kernel<<<ctx->num_sms * 8, 256, 0, st>>>(flat_d, off_d, n_docs, keys, vals, B->doc_len, err_d);
  SB_CUDA(cudaGetLastError());
  // number of significant token bits (sorting fewer bits = fewer radix passes)
  int32_t* max_d;
  if ((rc = tmp.alloc(&max_d, 1))) return rc;
  {
    size_t bytes = 0;
    SB_CUDA(cub::DeviceReduce::Max(nullptr, bytes, flat_d, max_d, (int)T, st));
    uint8_t* ws;
    if ((rc = tmp.alloc(&ws, bytes))) return rc;
    SB_CUDA(cub::DeviceReduce::Max(ws, bytes, flat_d, max_d, (int)T, st));
  }
  int32_t max_tok = 0, err_h = 0;
  SB_CUDA(cudaMemcpyAsync(&max_tok, max_d, 4, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(&err_h, err_d, 4, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  SB_REQUIRE(err_h == 0, SB_ERR_ARG, "sb_bm25_build_tokens: negative token id in the stream");
  int tok_bits = 1;
  while (tok_bits < 31 && (1ll << tok_bits) <= (int64_t)max_tok) ++tok_bits;
  {
    size_t bytes = 0;
    SB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys, keys2, vals, vals2, (int)T, 0, 32 + tok_bits, st));
    uint8_t* ws;
    if ((rc = tmp.alloc(&ws, bytes))) return rc;
    ctx->launches += 1;
    SB_CUDA(cub::DeviceRadixSort::SortPairs(ws, bytes, keys, keys2, vals, vals2, (int)T, 0, 32 + tok_bits, st));
  }
  // keys2 / vals2 = sorted stream; keys / vals are free again: reuse them as the int32 flag / index arrays
  run_head = reinterpret_cast<int32_t*>(keys);
  pair_head = run_head + T;
  run_idx = vals;
  if ((rc = tmp.alloc(&pair_idx, (size_t)T))) return rc;
  ctx->launches += 1;
  head_flags_kernel<<<blocks_for(T), 256, 0, st>>>(keys2, vals2, T, run_head, pair_head);
  SB_CUDA(cudaGetLastError());
  {
    size_t bytes = 0;
    SB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, run_head, run_idx, (int)T, st));
    uint8_t* ws;
    if ((rc = tmp.alloc(&ws, bytes))) return rc;
    SB_CUDA(cub::DeviceScan::ExclusiveSum(ws, bytes, run_head, run_idx, (int)T, st));
    SB_CUDA(cub::DeviceScan::ExclusiveSum(ws, bytes, pair_head, pair_idx, (int)T, st));
  }
  int32_t last[4];
  SB_CUDA(cudaMemcpyAsync(&last[0], run_idx + T - 1, 4, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(&last[1], run_head + T - 1, 4, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(&last[2], pair_idx + T - 1, 4, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(&last[3], pair_head + T - 1, 4, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  const int64_t V = (int64_t)last[0] + last[1], nnz = (int64_t)last[2] + last[3];
  B->V = V;
  B->nnz = nnz;
  uint32_t *run_first_pos, *run_first_sorted;
  int32_t *run_token, *run_pair_start, *run_iota, *sorted_runs, *rank_of_run, *pair_pos;
  if ((rc = tmp.alloc(&run_first_pos, (size_t)V))) return rc;
  if ((rc = tmp.alloc(&run_first_sorted, (size_t)V))) return rc;
  if ((rc = tmp.alloc(&run_token, (size_t)V))) return rc;
  if ((rc = tmp.alloc(&run_pair_start, (size_t)V))) return rc;
  if ((rc = tmp.alloc(&run_iota, (size_t)V))) return rc;
  if ((rc = tmp.alloc(&sorted_runs, (size_t)V))) return rc;
  if ((rc = tmp.alloc(&rank_of_run, (size_t)V))) return rc;
  if ((rc = tmp.alloc(&pair_pos, (size_t)nnz))) return rc;
  ctx->launches += 1;
  scatter_heads_kernel<<<blocks_for(T), 256, 0, st>>>(keys2, run_head, pair_head, run_idx, pair_idx, T, run_first_pos,
                                                     run_token, run_pair_start, run_iota, pair_pos);
  SB_CUDA(cudaGetLastError());
  {
    size_t bytes = 0;
    SB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, run_first_pos, run_first_sorted, run_iota, sorted_runs, (int)V, 0,
                                            32, st));
    uint8_t* ws;
    if ((rc = tmp.alloc(&ws, bytes))) return rc;
    SB_CUDA(cub::DeviceRadixSort::SortPairs(ws, bytes, run_first_pos, run_first_sorted, run_iota, sorted_runs, (int)V, 0, 32,
                                            st));
  }
  SB_CUDA(cudaMalloc(&B->df, (size_t)V * 8));
  SB_CUDA(cudaMalloc(&B->term_token, (size_t)V * 4));
  SB_CUDA(cudaMalloc(&B->indptr, (size_t)(V + 1) * 8));
  SB_CUDA(cudaMalloc(&B->post_doc, (size_t)std::max<int64_t>(nnz, 1) * 4));
  SB_CUDA(cudaMalloc(&B->tf, (size_t)std::max<int64_t>(nnz, 1) * 2));
  ctx->launches += 1;
  rank_runs_kernel<<<blocks_for(V), 256, 0, st>>>(sorted_runs, run_pair_start, run_token, V, nnz, rank_of_run, B->df,
                                                 B->term_token);
  SB_CUDA(cudaGetLastError());
  {
    size_t bytes = 0;
    SB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, B->df, B->indptr, (int)V, st));
    uint8_t* ws;
    if ((rc = tmp.alloc(&ws, bytes))) return rc;
    SB_CUDA(cub::DeviceScan::ExclusiveSum(ws, bytes, B->df, B->indptr, (int)V, st));
  }
  SB_CUDA(cudaMemcpyAsync(B->indptr + V, &nnz, 8, cudaMemcpyHostToDevice, st));
  ctx->launches += 1;
  scatter_postings_kernel<<<blocks_for(nnz), 256, 0, st>>>(pair_pos, run_idx, run_head, vals2, rank_of_run, run_pair_start, B->indptr,
                                                          nnz, T, B->post_doc, B->tf, err_d);
  SB_CUDA(cudaGetLastError());
  SB_CUDA(cudaMemcpyAsync(&err_h, err_d, 4, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  SB_REQUIRE(err_h == 0, SB_ERR_UNSUPPORTED,
             "sb_bm25_build_tokens: a term frequency above 65535 is not representable in the uint16 postings");
  guard.b = nullptr;
  ctx->bm25_build = B;
  *n_terms_out = V;
  *nnz_out = nnz;
  return SB_OK;
}

int sb_bm25_build_read(sb_ctx* ctx, int64_t* df_out, int32_t* term_token_out) {
  SB_REQUIRE(ctx != nullptr, SB_ERR_ARG, "sb_bm25_build_read: ctx is NULL");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  Bm25Build* B = ctx->bm25_build;
  SB_REQUIRE(B != nullptr, SB_ERR_STATE, "sb_bm25_build_read: no build in progress (sb_bm25_build_tokens)");
  if (df_out) SB_CUDA(cudaMemcpy(df_out, B->df, (size_t)B->V * 8, cudaMemcpyDeviceToHost));
  if (term_token_out) SB_CUDA(cudaMemcpy(term_token_out, B->term_token, (size_t)B->V * 4, cudaMemcpyDeviceToHost));
  return SB_OK;
}

int sb_bm25_build_export(sb_ctx* ctx, int64_t* indptr_out, int32_t* post_doc_out, uint16_t* post_tf_out,
                         int32_t* doc_len_out) {
  SB_REQUIRE(ctx != nullptr, SB_ERR_ARG, "sb_bm25_build_export: ctx is NULL");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  Bm25Build* B = ctx->bm25_build;
  SB_REQUIRE(B != nullptr, SB_ERR_STATE, "sb_bm25_build_export: no build in progress (sb_bm25_build_tokens)");
  if (indptr_out) SB_CUDA(cudaMemcpy(indptr_out, B->indptr, (size_t)(B->V + 1) * 8, cudaMemcpyDeviceToHost));
  if (post_doc_out) SB_CUDA(cudaMemcpy(post_doc_out, B->post_doc, (size_t)B->nnz * 4, cudaMemcpyDeviceToHost));
  if (post_tf_out) SB_CUDA(cudaMemcpy(post_tf_out, B->tf, (size_t)B->nnz * 2, cudaMemcpyDeviceToHost));
  if (doc_len_out) SB_CUDA(cudaMemcpy(doc_len_out, B->doc_len, (size_t)B->n_docs * 4, cudaMemcpyDeviceToHost));
  return SB_OK;
}

int sb_bm25_build_finish(sb_ctx* ctx, const double* idf, double avgdl, int32_t variant, double k1, double b, double delta,
                         int64_t id_base) {
  SB_REQUIRE(ctx != nullptr && idf != nullptr, SB_ERR_ARG, "sb_bm25_build_finish: NULL argument");
  SB_REQUIRE(variant == SB_BM25_OKAPI || variant == SB_BM25_PLUS, SB_ERR_ARG, "sb_bm25_build_finish: bad variant %d", variant);
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  Bm25Build* B = ctx->bm25_build;
  SB_REQUIRE(B != nullptr, SB_ERR_STATE, "sb_bm25_build_finish: no build in progress (sb_bm25_build_tokens)");
  int64_t* indptr = B->indptr;
  int32_t* post_doc = B->post_doc;
  B->indptr = nullptr;    // ownership of the CSR passes to bm25_install_device_csr unconditionally (it frees on failure)
  B->post_doc = nullptr;
  const int rc = bm25_install_device_csr(ctx, indptr, post_doc, B->tf, B->doc_len, B->V, B->nnz, B->n_docs, avgdl, idf,
                                         variant, k1, b, delta, id_base, ctx->stream);
  bm25_build_free(B);     // the build is consumed either way
  ctx->bm25_build = nullptr;
  return rc;
}

}  // extern "C"
