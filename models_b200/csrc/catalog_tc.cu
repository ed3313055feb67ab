// Query x catalog scoring without materialising the (B, N_I) logits: tcgen05 GEMM whose epilogue
// streams every 128x128 logits tile straight into per-row running statistics —
//   * soft-max cross-entropy inputs: row max, log-sum-exp, logit of the target item
//     (CategoricalCrossEntropy(from_logits=True), merlin/models/tf/losses/listwise.py:38-50), and/or
//   * top-k scores + item ids (tf.math.top_k: outputs/topk.py:221-223, core/index.py:236-237).
// Logits: x @ E^T (+ bias) as in EmbeddingTablePrediction.call (outputs/classification.py:347-357)
// and ItemRetrievalScorer._get_logits_for_sampled_softmax (blocks/retrieval/base.py:431-438).
//
// Work item = (128-query block, split of the item range).  The query tile (A, split-bf16) is
// loaded once per work item and stays in shared memory; item tiles (B operand, 128 items x Kp,
// split-bf16, precomputed once per catalog with mm_split_rows) stream through a TMA pipeline; the
// fp32 accumulator is double-buffered in TMEM so the reduction of tile j overlaps the MMAs of tile
// j+1.  Every epilogue thread owns one query row and keeps (max, sum-exp, target logit, top-k list)
// for the columns it sees; partial results go to a small workspace and a second kernel merges the
// 2*S partials per row.  3-pass split-bf16 = fp32-grade logits.
#include <cstring>

#include "tc_common.cuh"

namespace mm {
namespace cat {

using namespace mm::tc;

constexpr int BM = 128, BN = 128, BLOCK_K = 64, UMMA_K = 16;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int MAX_K = 32;  // top-k list length held per thread
constexpr uint32_t TILE_BYTES = 128 * BLOCK_K * 2;  // one 128-row x 64-col bf16 tile = 16 KB

struct Params {
  long long M, I;
  int Kp, KB, stages, S, tiles_per_split, n_tiles;
  const float* bias;
  const void* targets;
  int id_is64;
  int do_lse, topk;
  float* ws_lse;         // (M, 2S, 3)
  float* ws_vals;        // (M, 2S, topk)
  long long* ws_ids;     // (M, 2S, topk)
};

__global__ void __launch_bounds__(kThreads, 1)
catalog_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // pointer arithmetic keeps the shared state space (LDS/STS, not generic LD/ST)
  const uint32_t A_BYTES = 2u * p.KB * TILE_BYTES;      // [hi kb0..][lo kb0..]
  const uint32_t STAGE_BYTES = 2u * p.KB * TILE_BYTES;  // one item tile, all k-blocks, hi + lo
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + (size_t)p.stages * STAGE_BYTES);
  uint64_t* full_bar = bars;                        // [stages]
  uint64_t* empty_bar = bars + p.stages;            // [stages]
  uint64_t* tmem_full = bars + 2 * p.stages;        // [2]
  uint64_t* tmem_empty = bars + 2 * p.stages + 2;   // [2]
  uint64_t* a_full = bars + 2 * p.stages + 4;       // [1]
  uint64_t* a_empty = bars + 2 * p.stages + 5;      // [1]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 6);
  float* bias_s = reinterpret_cast<float*>(bars + 2 * p.stages + 8);  // [2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long m_blocks = (p.M + BM - 1) / BM;
  const long long items = m_blocks * p.S;
  const uint32_t tmem_cols = 256;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(full_bar + s), 1);
      mbar_init(smem_u32(empty_bar + s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(tmem_full + a), 1);
      mbar_init(smem_u32(tmem_empty + a), kEpiWarps);
    }
    mbar_init(smem_u32(a_full), 1);
    mbar_init(smem_u32(a_empty), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, a_phase = 0;
      for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const int m0 = (int)(item / p.S) * BM;
        const int split = (int)(item % p.S);
        const int t0 = split * p.tiles_per_split, t1 = min(p.n_tiles, t0 + p.tiles_per_split);
        // query tile: reused by every item tile of this work item
        mbar_wait(smem_u32(a_empty), a_phase ^ 1);
        mbar_expect_tx(smem_u32(a_full), A_BYTES);
        for (int kb = 0; kb < p.KB; ++kb) {
          tma_load_2d(smem_u32(smem_a + kb * TILE_BYTES), &tmA, smem_u32(a_full), kb * BLOCK_K, m0);
          tma_load_2d(smem_u32(smem_a + (p.KB + kb) * TILE_BYTES), &tmA, smem_u32(a_full), p.Kp + kb * BLOCK_K, m0);
        }
        a_phase ^= 1;
        for (int t = t0; t < t1; ++t) {
          mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
          const uint32_t fb = smem_u32(full_bar + stage);
          uint8_t* st = smem_b + (size_t)stage * STAGE_BYTES;
          mbar_expect_tx(fb, STAGE_BYTES);
          for (int kb = 0; kb < p.KB; ++kb) {
            tma_load_2d(smem_u32(st + kb * TILE_BYTES), &tmB, fb, kb * BLOCK_K, t * BN);
            tma_load_2d(smem_u32(st + (p.KB + kb) * TILE_BYTES), &tmB, fb, p.Kp + kb * BLOCK_K, t * BN);
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, BN);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0, a_phase = 0;
      for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const int split = (int)(item % p.S);
        const int t0 = split * p.tiles_per_split, t1 = min(p.n_tiles, t0 + p.tiles_per_split);
        mbar_wait(smem_u32(a_full), a_phase);
        a_phase ^= 1;
        tcgen05_fence_after();
        const uint32_t a_base = smem_u32(smem_a);
        for (int t = t0; t < t1; ++t) {
          mbar_wait(smem_u32(tmem_empty + acc), acc_phase ^ 1);
          mbar_wait(smem_u32(full_bar + stage), phase);
          tcgen05_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
          const uint32_t b_base = smem_u32(smem_b + (size_t)stage * STAGE_BYTES);
          uint32_t accumulate = 0;
          for (int kb = 0; kb < p.KB; ++kb) {
            const uint32_t a_hi = a_base + kb * TILE_BYTES, a_lo = a_base + (p.KB + kb) * TILE_BYTES;
            const uint32_t b_hi = b_base + kb * TILE_BYTES, b_lo = b_base + (p.KB + kb) * TILE_BYTES;
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              umma_bf16(d_tmem, make_desc_sw128(a_hi + k * 32), make_desc_sw128(b_lo + k * 32), idesc, accumulate);
              accumulate = 1;
            }
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_bf16(d_tmem, make_desc_sw128(a_lo + k * 32), make_desc_sw128(b_hi + k * 32), idesc, 1);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_bf16(d_tmem, make_desc_sw128(a_hi + k * 32), make_desc_sw128(b_hi + k * 32), idesc, 1);
          }
          tcgen05_commit(smem_u32(empty_bar + stage));
          tcgen05_commit(smem_u32(tmem_full + acc));
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
        tcgen05_commit(smem_u32(a_empty));  // all MMAs reading this query tile have retired
      }
    }
  } else {
    // ===================== epilogue: streaming reduction =====================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    const float LOG2E = 1.4426950408889634f;
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
      const long long m0 = (item / p.S) * BM;
      const int split = (int)(item % p.S);
      const int t0 = split * p.tiles_per_split, t1 = min(p.n_tiles, t0 + p.tiles_per_split);
      const long long row = m0 + q * 32 + lane;
      long long target = -1;
      if (p.do_lse && p.targets && row < p.M)
        target = p.id_is64 ? reinterpret_cast<const long long*>(p.targets)[row]
                           : (long long)reinterpret_cast<const int*>(p.targets)[row];
      float run_m = -INFINITY, run_s = 0.0f, tlogit = __int_as_float(0x7fc00000);  // NaN = "not seen"
      float tv[MAX_K];
      long long ti[MAX_K];
#pragma unroll
      for (int r = 0; r < MAX_K; ++r) {
        tv[r] = -INFINITY;
        ti[r] = -1;
      }
      for (int t = t0; t < t1; ++t) {
        const long long n0 = (long long)t * BN;
        float* bs = bias_s + acc * BN;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
        for (int i = threadIdx.x - 64; i < BN; i += 32 * kEpiWarps) bs[i] = (p.bias && n0 + i < p.I) ? p.bias[n0 + i] : 0.0f;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
        mbar_wait(smem_u32(tmem_full + acc), acc_phase);
        tcgen05_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
        for (int ch = half; ch < BN / 32; ch += 2) {
          const int c0 = ch << 5;
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_row + c0, r);
          tmem_ld_wait();
          if (ch + 2 >= BN / 32) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(tmem_empty + acc));
          }
          float v[32];
          float cmax = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = (n0 + c0 + j < p.I) ? __uint_as_float(r[j]) + bs[c0 + j] : -INFINITY;
            v[j] = x;
            cmax = fmaxf(cmax, x);
          }
          if (p.do_lse && cmax > -INFINITY) {
            const float m_new = fmaxf(run_m, cmax);
            float s = run_s * exp2f((run_m - m_new) * LOG2E);  // run_m = -inf -> 0
            const float mb = m_new * LOG2E;
#pragma unroll
            for (int j = 0; j < 32; ++j) s += exp2f(fmaf(v[j], LOG2E, -mb));
            run_m = m_new;
            run_s = s;
            const long long dt = target - (n0 + c0);
            if (dt >= 0 && dt < 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j == (int)dt) tlogit = v[j];
            }
          }
          if (p.topk > 0 && cmax > tv[p.topk - 1]) {
            // ascending item order inside a thread; strict '>' keeps the lower id on ties
#pragma unroll 1
            for (int j = 0; j < 32; ++j) {
              const float x = v[j];
              if (!(x > tv[p.topk - 1])) continue;
              int pos = p.topk - 1;
              while (pos > 0 && tv[pos - 1] < x) {
                tv[pos] = tv[pos - 1];
                ti[pos] = ti[pos - 1];
                --pos;
              }
              tv[pos] = x;
              ti[pos] = n0 + c0 + j;
            }
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      if (row < p.M) {
        const long long pidx = row * (2ll * p.S) + split * 2 + half;
        if (p.do_lse) {
          p.ws_lse[pidx * 3 + 0] = run_m;
          p.ws_lse[pidx * 3 + 1] = run_s;
          p.ws_lse[pidx * 3 + 2] = tlogit;
        }
        for (int r = 0; r < p.topk; ++r) {
          p.ws_vals[pidx * p.topk + r] = tv[r];
          p.ws_ids[pidx * p.topk + r] = ti[r];
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// merge the 2S partials of every row: one warp per row
__global__ void catalog_merge_kernel(long long M, int P, int do_lse, int topk, const float* __restrict__ ws_lse,
                                     const float* __restrict__ ws_vals, const long long* __restrict__ ws_ids,
                                     float* __restrict__ out_stats, float* __restrict__ out_scores,
                                     long long* __restrict__ out_ids) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = warp0; row < M; row += n_warps) {
    if (do_lse) {
      float m = -INFINITY, tl = __int_as_float(0x7fc00000);
      for (int i = lane; i < P; i += 32) {
        m = fmaxf(m, ws_lse[(row * P + i) * 3]);
        const float t = ws_lse[(row * P + i) * 3 + 2];
        if (t == t) tl = t;
      }
      m = warp_max(m);
      float s = 0.0f;
      for (int i = lane; i < P; i += 32) {
        const float mi = ws_lse[(row * P + i) * 3];
        if (mi > -INFINITY) s += ws_lse[(row * P + i) * 3 + 1] * expf(mi - m);
      }
      s = warp_sum(s);
      // the target logit lives in exactly one partial
      unsigned has = __ballot_sync(0xffffffffu, tl == tl);
      if (has) tl = __shfl_sync(0xffffffffu, tl, __ffs(has) - 1);
      if (lane == 0) {
        out_stats[row * 3 + 0] = m;
        out_stats[row * 3 + 1] = m + logf(s);
        out_stats[row * 3 + 2] = tl;
      }
    }
    if (topk > 0) {
      // k rounds of warp arg-max over the P*topk candidates, (value desc, id asc); picked ones are skipped
      const int C = P * topk;
      float last_v = INFINITY;
      long long last_i = -1;
      for (int r = 0; r < topk; ++r) {
        float bv = -INFINITY;
        long long bi = 0x7fffffffffffffffll;
        for (int c = lane; c < C; c += 32) {
          const float v = ws_vals[row * C + c];
          const long long id = ws_ids[row * C + c];
          if (id < 0) continue;
          // strictly after the previously emitted (value, id) in the output order
          const bool after = (v < last_v) || (v == last_v && id > last_i);
          if (!after) continue;
          if (v > bv || (v == bv && id < bi)) {
            bv = v;
            bi = id;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > bv || (ov == bv && oi < bi)) {
            bv = ov;
            bi = oi;
          }
        }
        if (lane == 0) {
          out_scores[row * topk + r] = bv;
          out_ids[row * topk + r] = (bi == 0x7fffffffffffffffll) ? -1 : bi;
        }
        last_v = bv;
        last_i = bi;
      }
    }
  }
}

static int splits_for(long long M, long long I) {
  const long long m_blocks = (M + BM - 1) / BM;
  const long long n_tiles = (I + BN - 1) / BN;
  long long S = (4ll * sm_count() + m_blocks - 1) / m_blocks;
  if (S > n_tiles) S = n_tiles;
  if (S > 64) S = 64;
  if (S < 1) S = 1;
  return (int)S;
}

}  // namespace cat
}  // namespace mm

extern "C" {

int mm_tc_padded_k(int K);

int64_t mm_catalog_workspace_bytes(int64_t B, int64_t I, int k) {
  if (B <= 0 || I <= 0 || k < 0) return 0;
  const int64_t P = 2ll * mm::cat::splits_for(B, I);
  return B * P * (3 * (int64_t)sizeof(float) + (int64_t)k * (sizeof(float) + sizeof(long long))) + 256;
}

int mm_catalog_score(const void* q_split, int64_t B, int D, const void* e_split, int64_t I, const float* bias,
                     const void* targets, int id_dtype, float* out_stats, int k, float* topk_scores,
                     int64_t* topk_ids, void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace mm::cat;
  MM_REQUIRE(q_split && e_split && workspace && B >= 0 && I > 0 && D > 0, MM_ERR_ARG, "mm_catalog_score: null pointer or bad size");
  MM_REQUIRE(out_stats || k > 0, MM_ERR_ARG, "mm_catalog_score: nothing requested (out_stats null and k == 0)");
  MM_REQUIRE(k >= 0 && k <= MAX_K && (k == 0 || (topk_scores && topk_ids)), MM_ERR_ARG,
             "mm_catalog_score: k must be 0..%d with score/id outputs", MAX_K);
  MM_REQUIRE(k <= I, MM_ERR_ARG, "mm_catalog_score: k exceeds the catalog size");
  MM_REQUIRE(id_dtype == MM_I32 || id_dtype == MM_I64, MM_ERR_ARG, "mm_catalog_score: bad id dtype");
  const int Kp = mm_tc_padded_k(D);
  MM_REQUIRE(Kp <= 128, MM_ERR_UNSUPPORTED, "mm_catalog_score: D up to 128 (query tile is kept resident in shared memory)");
  MM_REQUIRE(B < (1ll << 31) && I < (1ll << 31), MM_ERR_UNSUPPORTED, "mm_catalog_score: sizes exceed 32-bit TMA coordinates");
  MM_REQUIRE(workspace_bytes >= mm_catalog_workspace_bytes(B, I, k), MM_ERR_ARG, "mm_catalog_score: workspace too small (%lld < %lld)",
             (long long)workspace_bytes, (long long)mm_catalog_workspace_bytes(B, I, k));
  MM_REQUIRE(((uintptr_t)q_split % 16) == 0 && ((uintptr_t)e_split % 16) == 0 && ((uintptr_t)workspace % 16) == 0, MM_ERR_ALIGN,
             "mm_catalog_score: operands / workspace must be 16-B aligned");
  if (B == 0) return MM_OK;
  Params p;
  memset(&p, 0, sizeof(p));
  p.M = B;
  p.I = I;
  p.Kp = Kp;
  p.KB = Kp / BLOCK_K;
  p.S = splits_for(B, I);
  p.n_tiles = (int)((I + BN - 1) / BN);
  p.tiles_per_split = (p.n_tiles + p.S - 1) / p.S;
  p.bias = bias;
  p.targets = targets;
  p.id_is64 = id_dtype == MM_I64;
  p.do_lse = out_stats != nullptr;
  p.topk = k;
  const int64_t P = 2ll * p.S;
  uint8_t* ws = (uint8_t*)workspace;
  p.ws_lse = (float*)ws;
  p.ws_ids = (long long*)(ws + ((B * P * 3 * (int64_t)sizeof(float) + 15) / 16) * 16);
  p.ws_vals = (float*)((uint8_t*)p.ws_ids + B * P * (int64_t)k * sizeof(long long));
  const size_t a_bytes = 2ull * p.KB * TILE_BYTES, stage_bytes = 2ull * p.KB * TILE_BYTES;
  int stages = (int)((220 * 1024 - 2048 - a_bytes) / stage_bytes);
  if (stages > 4) stages = 4;
  MM_REQUIRE(stages >= 2, MM_ERR_UNSUPPORTED, "mm_catalog_score: tiles do not fit two pipeline stages");
  p.stages = stages;
  const size_t smem = 1024 + a_bytes + stages * stage_bytes + (2 * stages + 8) * sizeof(uint64_t) + 2 * BN * sizeof(float);
  CUtensorMap tmA, tmB;
  int rc = make_map(&tmA, q_split, (uint64_t)B, (uint64_t)2 * Kp, BM);
  if (rc) return rc;
  rc = make_map(&tmB, e_split, (uint64_t)I, (uint64_t)2 * Kp, BN);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(catalog_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      mm::set_error("mm_catalog_score: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr = true;
  }
  const long long items = ((B + BM - 1) / BM) * p.S;
  const int sms = mm::sm_count();
  const unsigned grid = (unsigned)(items < sms ? items : sms);
  cudaStream_t st = (cudaStream_t)stream;
  catalog_kernel<<<grid, kThreads, smem, st>>>(tmA, tmB, p);
  rc = mm::check_launch("mm_catalog_score");
  if (rc) return rc;
  long long blocks = (B * 32 + 255) / 256;
  const long long cap = (long long)sms * 16;
  if (blocks > cap) blocks = cap;
  catalog_merge_kernel<<<(unsigned)blocks, 256, 0, st>>>(B, (int)P, p.do_lse, k, p.ws_lse, p.ws_vals, p.ws_ids, out_stats,
                                                        topk_scores, (long long*)topk_ids);
  return mm::check_launch("mm_catalog_score(merge)");
}

}  // extern "C"
