// dense_mma.cu -- K1b: batched-query dense scan on the 5th-gen tensor cores (tcgen05 / TMEM / TMA).
//
// One HBM pass over the fp16 corpus serves QBN = 16 / 32 / 64 queries at once: the scores of a 128-row corpus tile against
// all QBN queries are one UMMA accumulator  D[128, QBN] = A[128, D] * Q[QBN, D]^T  (A = corpus tile, K-major fp16,
// streamed by TMA in 64-column SWIZZLE_128B boxes; B = the query block, K-major fp16, resident in shared memory for the
// whole kernel; D in tensor memory, double buffered).  The pass stays HBM bound: per 16 KB of corpus the tensor pipe
// needs 4 MMAs of 128 x QBN x 16, a few percent of its capacity.
//
// Exactness is kept by construction, not by luck:
//   * every query gets a SAFE initial threshold from a sampling pass (the K'-th best approximate score over a sample of
//     corpus tiles is <= the global K'-th best), so only ~1 % of the rows survive the epilogue compare;
//   * survivors are appended to a per-(CTA, query) global buffer that is sized for the worst case (every row of the
//     CTA's share) -- nothing is ever dropped, an adversarial corpus order only costs time in the select kernel;
//   * dense_select_kernel reduces each query's survivors to the K' best approximate keys (chunked bitonic sort) and the
//     shared exact stage (dense_common.cuh) re-scores them in fp64 against the stored rows and the fp32 query.
//
// Warp roles (192 threads, 1 CTA / SM, persistent): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld 32x32b, one corpus row per thread, QBN scores in registers).
#include <cuda.h>

#include <algorithm>
#include <vector>

#include "dense_common.cuh"
#include "dense_mma.cuh"

namespace {

constexpr int kTileRows = 128;
constexpr int kBK = 64;                       // fp16 elements per 128-byte swizzle row
constexpr uint32_t kATileBytes = kTileRows * kBK * 2;   // 16 KB
constexpr int kMmaThreads = 192;
constexpr int kSelectThreads = 1024;
constexpr int kSelStage = 20480;              // survivors staged in shared memory by dense_select_kernel (160 KB)
constexpr int kSelTop = 2048;                 // winner buffer: < K' keys above the selected bucket + the bucket

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;             // LBO (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;   // SBO: 8 rows x 128 B
  d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;             // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

template <int N>
struct TmemLd;
template <>
struct TmemLd<16> {
  static __device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
  }
};

struct MmaScanParams {
  const float* inv_norm;
  const float* thr_init;        // [QBN] safe initial thresholds (NULL = -inf: sampling pass)
  unsigned long long* cand;     // [QBN][grid][capg]
  int32_t* counts;              // [QBN][grid]
  int64_t n;                    // valid rows
  int32_t kb_count;             // d_pad / 64
  int32_t num_tiles;            // tiles visited by this launch
  int32_t tile_first, tile_step; // global tile index = tile_first + t * tile_step,  t in [0, num_tiles)
  int32_t capg;
  int32_t stages;
};

template <int QBN>
__global__ void __launch_bounds__(kMmaThreads, 1)
dense_scan_mma_kernel(const __grid_constant__ CUtensorMap tm_rows, const __grid_constant__ CUtensorMap tm_q,
                      const MmaScanParams p) {
  extern __shared__ uint8_t msm_raw[];
  const uint32_t raw = smem_u32(msm_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B operands need 1024-byte alignment
  uint8_t* sm = msm_raw + (base - raw);
  constexpr uint32_t kQBlockBytes = QBN * kBK * 2;      // one 64-column block of the query operand
  const uint32_t q_bytes = (uint32_t)p.kb_count * kQBlockBytes;
  const uint32_t a0 = base + q_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + q_bytes + (size_t)p.stages * kATileBytes);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + p.stages);
  const uint32_t bar_q = smem_u32(bars + 2 * p.stages);
  const uint32_t bar_acc_full = smem_u32(bars + 2 * p.stages + 1), bar_acc_empty = smem_u32(bars + 2 * p.stages + 3);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 5);
  volatile float* thr = reinterpret_cast<volatile float*>(tmem_slot + 2);   // [QBN]
  int* cnt = reinterpret_cast<int*>(const_cast<float*>(thr) + QBN);        // [QBN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grid = gridDim.x, cta = blockIdx.x;
  const int my_tiles = cta < p.num_tiles ? (p.num_tiles - 1 - cta) / grid + 1 : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_q, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_acc_full + 8 * s, 1);
      mbar_init(bar_acc_empty + 8 * s, 4);  // one arrival per epilogue warp
    }
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_rows) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_q) : "memory");
  }
  for (int i = threadIdx.x; i < QBN; i += blockDim.x) {
    thr[i] = p.thr_init ? p.thr_init[i] : -INFINITY;
    cnt[i] = 0;
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(2 * QBN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      mbar_expect_tx(bar_q, q_bytes);
      for (int kb = 0; kb < p.kb_count; ++kb) tma_load_2d(base + (uint32_t)kb * kQBlockBytes, &tm_q, kb * kBK, 0, bar_q);
      int it = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int tile = p.tile_first + (cta + t * grid) * p.tile_step;
        for (int kb = 0; kb < p.kb_count; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t use = (uint32_t)(it / p.stages);
          if (it >= p.stages) mbar_wait(bar_empty + 8 * s, (use & 1u) ^ 1u);
          mbar_expect_tx(bar_full + 8 * s, kATileBytes);
          tma_load_2d(a0 + (uint32_t)s * kATileBytes, &tm_rows, kb * kBK, tile * kTileRows, bar_full + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      // kind::f16: D = f32, A = B = f16 K-major, N >> 3 at [17,23), M >> 4 at [24,29)
      const uint32_t idesc = (1u << 4) | ((uint32_t)(QBN >> 3) << 17) | ((uint32_t)(kTileRows >> 4) << 24);
      mbar_wait(bar_q, 0);
      int it = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int as = t & 1;
        if (t >= 2) mbar_wait(bar_acc_empty + 8 * as, (((uint32_t)t >> 1) & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * QBN);
        for (int kb = 0; kb < p.kb_count; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t use = (uint32_t)(it / p.stages);
          mbar_wait(bar_full + 8 * s, use & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = make_smem_desc_sw128(a0 + (uint32_t)s * kATileBytes);
          const uint64_t db = make_smem_desc_sw128(base + (uint32_t)kb * kQBlockBytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          umma_commit(bar_empty + 8 * s);
        }
        umma_commit(bar_acc_full + 8 * as);
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps 2..5: one corpus row per thread
    const int quad = warp & 3;
    unsigned long long* my_cand = p.cand + (size_t)cta * p.capg;
    const size_t q_stride = (size_t)grid * p.capg;
    for (int t = 0; t < my_tiles; ++t) {
      const int as = t & 1;
      const int tile = p.tile_first + (cta + t * grid) * p.tile_step;
      const int64_t row = (int64_t)tile * kTileRows + quad * 32 + lane;
      const float invn = row < p.n ? __ldg(p.inv_norm + row) : 0.f;
      mbar_wait(bar_acc_full + 8 * as, ((uint32_t)t >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int c0 = 0; c0 < QBN; c0 += 16) {
        uint32_t v[16];
        TmemLd<16>::ld(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * QBN + c0), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < p.n) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float score = __uint_as_float(v[j]) * invn;
            if (score >= thr[c0 + j]) {
              const int pos = atomicAdd(&cnt[c0 + j], 1);
              if (pos < p.capg) my_cand[(size_t)(c0 + j) * q_stride + pos] = make_key32(score, (uint32_t)row);
            }
          }
        }
      }
      // the accumulator stage has been read into registers by this warp -> give it back to the MMA issuer
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + 8 * as);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  for (int i = threadIdx.x; i < QBN; i += blockDim.x) p.counts[(size_t)i * grid + cta] = min(cnt[i], p.capg);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * QBN) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ select kernel
struct SelectParams {
  const unsigned long long* cand;   // [nq][grid][capg]
  const int32_t* counts;            // [nq][grid]
  int32_t grid, capg, kprime;
  int32_t nq;                       // real queries in this block (padded operand rows get a +inf threshold)
  int32_t mode;                     // 0 = write the K'-th best approximate score (threshold pass), 1 = exact stage + emit
  float* thr_out;                   // [nq]  (mode 0)
  const __half* rows;               // mode 1
  const float* q;                   // [nq][d_pad]
  int32_t d_pad, ch;
  int64_t id_base;
  int32_t k;
  int64_t* out_ids;
  double* out_scores;
  int32_t* out_counts;
};

__device__ __forceinline__ void select_sort_desc(unsigned long long* a, int len, int tid, int nt) {
  for (int k = 2; k <= len; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < len; i += nt) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long x = a[i], y = a[ixj];
          const bool desc = (i & k) == 0;
          if (desc ? (x < y) : (x > y)) {
            a[i] = y;
            a[ixj] = x;
          }
        }
      }
      __syncthreads();
    }
  }
}

// One CTA per query: the K' best approximate keys among the survivors of all CTAs by an MSB-first RADIX SELECT
// (8-bit digits, shared-memory histogram; the pass loop stops as soon as the bucket holding the K'-th key is small),
// then a small sort of {keys above the bucket} U {bucket}.  Survivors are staged in shared memory when they fit
// (the normal case: ~1.5 % of the corpus); otherwise every pass streams them from HBM/L2 -- slower, still exact.
__global__ void __launch_bounds__(kSelectThreads, 1) dense_select_kernel(const SelectParams p) {
  extern __shared__ __align__(16) uint8_t ssm[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(ssm);  // [kSelStage] staged survivors
  unsigned long long* top = keys + kSelStage;                             // [kSelTop]   gathered winners
  unsigned long long* ek = top + kSelTop;                                 // [K]
  uint32_t* ei = reinterpret_cast<uint32_t*>(ek + p.kprime);              // [K]
  __shared__ double qq_s;
  __shared__ int s_prefix[1024 + 1];
  __shared__ int s_hist[256];
  __shared__ int s_scal[4];
  __shared__ int s_ntop;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5, qi = blockIdx.x;
  const int K = p.kprime, G = p.grid;
  const int32_t* counts = p.counts + (size_t)qi * G;
  const unsigned long long* cand = p.cand + (size_t)qi * G * p.capg;
  // exclusive prefix of the per-CTA survivor counts (G <= 1024): one element per thread, warp scans + a scan of the sums
  {
    const int v = tid < G ? counts[tid] : 0;
    int x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_hist[warp] = x;
    if (tid == 0) s_ntop = 0;
    __syncthreads();
    if (warp == 0) {
      int w = lane < nw ? s_hist[lane] : 0;
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      s_hist[32 + lane] = w;  // inclusive warp totals
    }
    __syncthreads();
    const int incl = x + (warp ? s_hist[32 + warp - 1] : 0);
    if (tid < G) s_prefix[tid] = incl - v;
    if (tid == G - 1) s_prefix[G] = incl;
    __syncthreads();
  }
  const int total = s_prefix[G];
  const bool staged = total <= kSelStage;
  if (staged) {
    for (int g = warp; g < G; g += nw) {
      const int c = s_prefix[g + 1] - s_prefix[g];
      const unsigned long long* src = cand + (size_t)g * p.capg;
      unsigned long long* dst = keys + s_prefix[g];
      for (int i = lane; i < c; i += 32) dst[i] = src[i];
    }
    __syncthreads();
  }
  // visit every survivor once: `body(key)`
#define SB_FOR_EACH_KEY(BODY)                                                       \
  if (staged) {                                                                     \
    for (int i_ = tid; i_ < total; i_ += nt) {                                      \
      const unsigned long long key = keys[i_];                                      \
      BODY                                                                          \
    }                                                                               \
  } else {                                                                          \
    for (int g_ = warp; g_ < G; g_ += nw) {                                         \
      const int c_ = s_prefix[g_ + 1] - s_prefix[g_];                               \
      const unsigned long long* src_ = cand + (size_t)g_ * p.capg;                  \
      for (int i_ = lane; i_ < c_; i_ += 32) {                                      \
        const unsigned long long key = src_[i_];                                    \
        BODY                                                                        \
      }                                                                             \
    }                                                                               \
  }

  unsigned long long prefix = 0ull, mask = 0ull;
  if (total > K) {
    int need = K;  // rank (from the top) of the key we are looking for inside the current bucket
    for (int shift = 56; shift >= 0; shift -= 8) {
      for (int i = tid; i < 256; i += nt) s_hist[i] = 0;
      __syncthreads();
      // warp-aggregated histogram update: in the leading passes nearly all keys share a digit, and per-key shared
      // atomics on one bin serialise; lanes with equal digits elect one lane to add their count
      SB_FOR_EACH_KEY(if ((key & mask) == prefix) {
        const int dgt = (int)((key >> shift) & 0xffull);
        const unsigned grp = __match_any_sync(__activemask(), dgt);
        if ((int)(threadIdx.x & 31) == __ffs(grp) - 1) atomicAdd(&s_hist[dgt], __popc(grp));
      })
      __syncthreads();
      if (warp == 0) warp_select_bin<true>(s_hist, need, lane, s_scal);
      __syncthreads();
      prefix |= (unsigned long long)s_scal[0] << shift;
      mask |= 0xffull << shift;
      need = s_scal[1];
      const int bucket = s_scal[2];
      __syncthreads();
      // mode 1: {above} (< K keys) + a small bucket: finish by sorting.  mode 0 only needs a LOWER BOUND of the K-th
      // best key, so it narrows the bucket a little further and takes the bucket's lower edge -- no gather, no sort.
      if (p.mode == 0 ? bucket <= 16 : (bucket <= 256 && bucket <= kSelTop - K)) break;
    }
  }
  if (p.mode == 0) {
    // undecided low bits of `prefix` are zero: a key <= the K-th best; its score field (possibly with cleared low bits)
    // is a safe threshold.  Fewer than K survivors: no threshold.
    if (tid == 0) p.thr_out[qi] = qi >= p.nq ? INFINITY : (total > K ? key32_score(prefix) : -INFINITY);
    return;
  }
  // winners: every key whose decided digits are >= the selected bucket's (at most K - 1 + bucket <= kSelTop keys)
  SB_FOR_EACH_KEY(if ((key & mask) >= prefix) {
    const int at = atomicAdd(&s_ntop, 1);
    if (at < kSelTop) top[at] = key;
  })
#undef SB_FOR_EACH_KEY
  __syncthreads();
  const int ntop = min(s_ntop, kSelTop);
  int P = 32;
  while (P < ntop || P < K) P <<= 1;
  for (int i = ntop + tid; i < P; i += nt) top[i] = 0ull;
  __syncthreads();
  select_sort_desc(top, P, tid, nt);
  RescoreArgs ra;
  ra.rows = p.rows;
  ra.q = p.q + (size_t)qi * p.d_pad;
  ra.d_pad = p.d_pad;
  ra.ch = p.ch;
  ra.id_base = p.id_base;
  ra.k = p.k;
  ra.out_ids = p.out_ids + (size_t)qi * p.k;
  ra.out_scores = p.out_scores + (size_t)qi * p.k;
  ra.out_count = p.out_counts + qi;
  rescore_and_emit(top, K, ek, ei, &qq_s, ra);
}

// fp32 padded queries -> fp16 operand block [QBN][d_pad] (rows beyond nq are zero)
__global__ void queries_to_f16_kernel(const float* __restrict__ q_pad, int nq, int qbn, int d_pad, __half* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)qbn * d_pad) return;
  const int r = (int)(i / d_pad);
  out[i] = r < nq ? __float2half_rn(q_pad[i]) : __float2half_rn(0.f);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_map(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int box_rows) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* pfn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &pfn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(pfn);
  }
  SB_REQUIRE(fn != nullptr, SB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SB_REQUIRE(r == CUDA_SUCCESS, SB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return SB_OK;
}

template <int QBN>
int launch_mma(const CUtensorMap& tm_rows, const CUtensorMap& tm_q, const MmaScanParams& mp, int grid, size_t smem,
               cudaStream_t st) {
  auto kern = dense_scan_mma_kernel<QBN>;
  SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kMmaThreads, smem, st>>>(tm_rows, tm_q, mp);
  SB_CUDA(cudaGetLastError());
  return SB_OK;
}

int dispatch_mma(int qbn, const CUtensorMap& tm_rows, const CUtensorMap& tm_q, const MmaScanParams& mp, int grid,
                 size_t smem, cudaStream_t st) {
  switch (qbn) {
    case 16: return launch_mma<16>(tm_rows, tm_q, mp, grid, smem, st);
    case 32: return launch_mma<32>(tm_rows, tm_q, mp, grid, smem, st);
    case 64: return launch_mma<64>(tm_rows, tm_q, mp, grid, smem, st);
  }
  sb_set_error("dense_mma: unsupported query block %d", qbn);
  return SB_ERR_UNSUPPORTED;
}

}  // namespace

bool dense_mma_eligible(const sb_ctx* ctx, const DenseIndex& ix, int B) {
  if (B < 16) return false;
  if (ix.d_pad % kBK != 0 || ix.n_pad % kTileRows != 0) return false;
  if (ix.n < 64 * kTileRows) return false;  // tiny corpora: the CUDA-core scan is already launch bound
  const size_t need = (size_t)ix.d_pad * 16 * 2 + 3 * (size_t)kATileBytes + 4096;
  return need <= ctx->smem_optin;
}

int dense_mma_topk_enqueue(sb_ctx* ctx, DenseIndex& ix, const float* q_pad, int B, int k, int kprime, int64_t* out_ids,
                           double* out_scores, int32_t* out_counts, cudaStream_t st) {
  const int kb_count = ix.d_pad / kBK;
  const int total_tiles = (int)(ix.n_pad / kTileRows);
  // largest query block whose resident operand leaves >= 3 pipeline stages
  int qbn_max = 64;
  while (qbn_max > 16 && (size_t)qbn_max * ix.d_pad * 2 + 3 * (size_t)kATileBytes + 4096 > ctx->smem_optin) qbn_max >>= 1;
  const int grid = std::min(ctx->num_sms, total_tiles);
  const int capg = ((total_tiles + grid - 1) / grid) * kTileRows;  // worst case: every row of the CTA's share survives
  int rc;
  // Query groups of qbn_max (the last one may use a smaller operand block).  Several groups are kept in flight so that
  // the two select launches (one CTA per query) cover ALL their queries at once instead of qbn at a time: with small
  // shards the selects, not the scans, would otherwise dominate.  In-flight groups are bounded by the survivor scratch.
  const int n_groups = (B + qbn_max - 1) / qbn_max;
  const size_t cand_per_group = (size_t)qbn_max * grid * capg * 8;
  int gmax = (int)std::max<size_t>(1, (size_t)(2048ull << 20) / cand_per_group);
  gmax = std::min(gmax, n_groups);
  // scratch: fp16 query blocks | thresholds | counts | survivors
  const size_t q16_group = (size_t)qbn_max * ix.d_pad * 2;
  if ((rc = ctx->misc2_dev.reserve(q16_group * gmax + 256))) return rc;
  if ((rc = ctx->misc3_dev.reserve(((size_t)qbn_max * 4 + (size_t)qbn_max * grid * 4) * gmax + 256))) return rc;
  if ((rc = ctx->cand_dev.reserve(cand_per_group * gmax))) return rc;
  __half* q16 = ctx->misc2_dev.as<__half>();
  float* thr = ctx->misc3_dev.as<float>();
  int32_t* counts = reinterpret_cast<int32_t*>(thr + (size_t)qbn_max * gmax);
  unsigned long long* cand = ctx->cand_dev.as<unsigned long long>();
  if (ix.tm_rows_ptr != ix.rows) {  // (re)build the corpus tensor map once per loaded index
    if ((rc = encode_map(reinterpret_cast<CUtensorMap*>(ix.tm_rows), ix.rows, ix.n_pad, ix.d_pad, kTileRows))) return rc;
    ix.tm_rows_ptr = ix.rows;
  }
  const CUtensorMap& tm_rows = *reinterpret_cast<const CUtensorMap*>(ix.tm_rows);
  const size_t sel_smem = (size_t)(kSelStage + kSelTop) * 8 + (size_t)kprime * 12 + 64;
  SB_CUDA(cudaFuncSetAttribute(dense_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
  // sampling pass geometry: ~64 tiles spread evenly over the corpus
  const int sample_tiles = std::min(64, total_tiles);
  const int sample_step = total_tiles / sample_tiles;
  const int sgrid = std::min(grid, sample_tiles);
  for (int c0 = 0; c0 < B; c0 += gmax * qbn_max) {
    const int nq_chunk = std::min(B - c0, gmax * qbn_max);   // real queries of this chunk of groups
    const int ng = (nq_chunk + qbn_max - 1) / qbn_max;
    struct Group { int qbn, nq; CUtensorMap tm_q; size_t smem; int stages; };
    std::vector<Group> gs((size_t)ng);
    int rows_total = 0;  // operand rows of the chunk (padding only at the very end)
    for (int g = 0; g < ng; ++g) {
      Group& G = gs[(size_t)g];
      const int left = nq_chunk - g * qbn_max;
      G.qbn = qbn_max;
      while (G.qbn > 16 && G.qbn / 2 >= left) G.qbn >>= 1;
      G.nq = std::min(G.qbn, left);
      __half* q16g = q16 + (size_t)g * qbn_max * ix.d_pad;
      if ((rc = encode_map(&G.tm_q, q16g, G.qbn, ix.d_pad, G.qbn))) return rc;
      const int64_t nconv = (int64_t)G.qbn * ix.d_pad;
      ctx->launches += 1;
      queries_to_f16_kernel<<<(unsigned)((nconv + 255) / 256), 256, 0, st>>>(
          q_pad + (size_t)(c0 + g * qbn_max) * ix.d_pad, G.nq, G.qbn, ix.d_pad, q16g);
      SB_CUDA(cudaGetLastError());
      const size_t q_bytes = (size_t)G.qbn * ix.d_pad * 2;
      G.stages = std::max(3, std::min((int)((ctx->smem_optin - q_bytes - 4096) / kATileBytes), 8));
      G.smem = q_bytes + (size_t)G.stages * kATileBytes + 2048 + 1024;
      rows_total = g * qbn_max + G.qbn;
    }
    MmaScanParams mp;
    mp.inv_norm = ix.inv_norm;
    mp.n = ix.n;
    mp.kb_count = kb_count;
    mp.capg = capg;
    SelectParams sp;
    sp.cand = cand;
    sp.counts = counts;
    sp.capg = capg;
    sp.kprime = kprime;
    sp.nq = nq_chunk;
    sp.thr_out = thr;
    sp.rows = ix.rows;
    sp.q = q_pad + (size_t)c0 * ix.d_pad;
    sp.d_pad = ix.d_pad;
    sp.ch = ix.d_pad / 8;
    sp.id_base = ix.id_base;
    sp.k = k;
    sp.out_ids = out_ids + (size_t)c0 * k;
    sp.out_scores = out_scores + (size_t)c0 * k;
    sp.out_counts = out_counts + c0;
    // (1) sampling passes -> safe thresholds for every query of the chunk (one select launch)
    mp.thr_init = nullptr;
    mp.num_tiles = sample_tiles;
    mp.tile_first = 0;
    mp.tile_step = sample_step;
    for (int g = 0; g < ng; ++g) {
      const Group& G = gs[(size_t)g];
      mp.cand = cand + (size_t)g * qbn_max * sgrid * capg;
      mp.counts = counts + (size_t)g * qbn_max * sgrid;
      mp.stages = G.stages;
      ctx->launches += 1;
      if ((rc = dispatch_mma(G.qbn, tm_rows, G.tm_q, mp, sgrid, G.smem, st))) return rc;
    }
    sp.grid = sgrid;
    sp.mode = 0;
    ctx->launches += 1;
    dense_select_kernel<<<rows_total, kSelectThreads, sel_smem, st>>>(sp);
    SB_CUDA(cudaGetLastError());
    // (2) the full passes, then one select + exact re-score launch for the chunk
    mp.num_tiles = total_tiles;
    mp.tile_first = 0;
    mp.tile_step = 1;
    for (int g = 0; g < ng; ++g) {
      const Group& G = gs[(size_t)g];
      mp.thr_init = thr + (size_t)g * qbn_max;
      mp.cand = cand + (size_t)g * qbn_max * grid * capg;
      mp.counts = counts + (size_t)g * qbn_max * grid;
      mp.stages = G.stages;
      ProfScope ps(ctx, SB_PROF_DENSE_SCAN, st);
      if ((rc = dispatch_mma(G.qbn, tm_rows, G.tm_q, mp, grid, G.smem, st))) return rc;
    }
    sp.grid = grid;
    sp.mode = 1;
    {
      ProfScope ps(ctx, SB_PROF_DENSE_MERGE, st);
      dense_select_kernel<<<nq_chunk, kSelectThreads, sel_smem, st>>>(sp);
    }
    SB_CUDA(cudaGetLastError());
  }
  return SB_OK;
}
