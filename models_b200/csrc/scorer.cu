// Two-tower scoring: row-wise positive score and in-batch / sampled negatives logits with the
// false-negative mask, logQ correction, [pos | neg] concat and temperature folded into the GEMM
// epilogue, so the (B, 1+N) logits tensor is written exactly once.
// Replaces ItemRetrievalScorer.call/call_outputs (merlin/models/tf/blocks/retrieval/base.py:
// 271-281, 339-396), rescore_false_negatives (utils/tf_utils.py:126-154),
// ContrastiveOutput.outputs (outputs/contrastive.py:303-326) and LogitsTemperatureScaler.
#include "mm_common.cuh"

namespace mm {

// one warp per row; lanes stride over D
__global__ void rowwise_dot_kernel(const float* __restrict__ q, const float* __restrict__ it,
                                   long long B, int D, long long q_stride, long long i_stride,
                                   const float* __restrict__ pos_prob, float temperature,
                                   float* __restrict__ out, long long out_stride) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long b = warp0; b < B; b += n_warps) {
    float acc = 0.0f;
    for (int d = lane; d < D; d += 32) acc = fmaf(q[b * q_stride + d], it[b * i_stride + d], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      if (pos_prob) acc -= logf(pos_prob[b] + 1e-16f);
      out[b * out_stride] = acc / temperature;
    }
  }
}

constexpr int SBM = 64, SBN = 64, SBK = 16;

// C[b, n] = q[b,:] . neg[n,:]   (both K-major), epilogue = mask / logQ / temperature,
// written at column 1+n of `out`.
template <typename IdT>
__global__ void __launch_bounds__(256)
inbatch_scores_kernel(const float* __restrict__ q, const float* __restrict__ neg, long long B,
                      long long N, int D, const IdT* __restrict__ pos_ids,
                      const IdT* __restrict__ neg_ids, int downscore, float false_neg_score,
                      const float* __restrict__ neg_prob, float temperature,
                      float* __restrict__ out, long long out_stride) {
  __shared__ float As[SBK][SBM + 4];
  __shared__ float Bs[SBK][SBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.y * SBM;
  const long long n0 = (long long)blockIdx.x * SBN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  const int l_r = tid >> 2, l_k = (tid & 3) * 4;
  for (int k0 = 0; k0 < D; k0 += SBK) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = k0 + l_k + u;
      const long long m = m0 + l_r, n = n0 + l_r;
      As[l_k + u][l_r] = (m < B && k < D) ? q[m * D + k] : 0.0f;
      Bs[l_k + u][l_r] = (n < N && k < D) ? neg[n * D + k] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SBK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= B) continue;
    const IdT pid = (downscore && pos_ids) ? pos_ids[m] : (IdT)0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (neg_prob) v -= logf(neg_prob[n] + 1e-16f);
      if (downscore && pos_ids && neg_ids[n] == pid) v = false_neg_score;
      out[m * out_stride + 1 + n] = v / temperature;
    }
  }
}

}  // namespace mm

extern "C" {

int mm_rowwise_dot(const float* q, const float* items, int64_t B, int D, int64_t q_stride,
                   int64_t i_stride, float* out, void* stream) {
  MM_REQUIRE(q && items && out && B >= 0 && D > 0 && q_stride >= D && i_stride >= D, MM_ERR_ARG,
             "mm_rowwise_dot: null pointer, D<=0 or stride < D");
  if (B == 0) return MM_OK;
  const int threads = 256;
  long long blocks = (B * 32 + threads - 1) / threads;
  const long long cap = (long long)mm::sm_count() * 32;
  if (blocks > cap) blocks = cap;
  mm::rowwise_dot_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
      q, items, B, D, q_stride, i_stride, nullptr, 1.0f, out, 1);
  return mm::check_launch("mm_rowwise_dot");
}

int mm_positive_scores(const float* q, const float* pos, int64_t B, int D, const float* pos_prob,
                       float temperature, float* out, int64_t out_stride, void* stream) {
  MM_REQUIRE(q && pos && out && B >= 0 && D > 0 && out_stride >= 1, MM_ERR_ARG, "mm_positive_scores: null pointer or D<=0");
  MM_REQUIRE(temperature != 0.0f, MM_ERR_ARG, "mm_positive_scores: temperature must be non-zero");
  if (B == 0) return MM_OK;
  const int threads = 256;
  long long blocks = (B * 32 + threads - 1) / threads;
  const long long cap = (long long)mm::sm_count() * 32;
  if (blocks > cap) blocks = cap;
  mm::rowwise_dot_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(q, pos, B, D, D, D, pos_prob, temperature,
                                                                                out, out_stride);
  return mm::check_launch("mm_positive_scores");
}

int mm_inbatch_scores(const float* q, const float* pos, const float* neg, int64_t B, int64_t N,
                      int D, const void* pos_ids, const void* neg_ids, int id_dtype, int downscore,
                      float false_neg_score, const float* pos_prob, const float* neg_prob,
                      float temperature, float* out, int64_t out_stride, void* stream) {
  MM_REQUIRE(q && pos && neg && out && B >= 0 && N >= 0 && D > 0, MM_ERR_ARG,
             "mm_inbatch_scores: null pointer or D<=0");
  MM_REQUIRE(out_stride >= N + 1, MM_ERR_ARG, "mm_inbatch_scores: out_stride < 1+N");
  MM_REQUIRE(!downscore || (pos_ids && neg_ids), MM_ERR_ARG,
             "mm_inbatch_scores: downscore_false_negatives needs positive and negative ids");
  MM_REQUIRE(id_dtype == MM_I32 || id_dtype == MM_I64, MM_ERR_ARG, "mm_inbatch_scores: bad id dtype");
  MM_REQUIRE(temperature != 0.0f, MM_ERR_ARG, "mm_inbatch_scores: temperature must be non-zero");
  if (B == 0) return MM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  {
    const int threads = 256;
    long long blocks = (B * 32 + threads - 1) / threads;
    const long long cap = (long long)mm::sm_count() * 32;
    if (blocks > cap) blocks = cap;
    mm::rowwise_dot_kernel<<<(unsigned)blocks, threads, 0, st>>>(q, pos, B, D, D, D, pos_prob,
                                                                 temperature, out, out_stride);
    int rc = mm::check_launch("mm_inbatch_scores(positive)");
    if (rc) return rc;
  }
  if (N == 0) return MM_OK;
  dim3 grid((unsigned)((N + mm::SBN - 1) / mm::SBN), (unsigned)((B + mm::SBM - 1) / mm::SBM));
  MM_REQUIRE(grid.y <= 65535, MM_ERR_UNSUPPORTED, "mm_inbatch_scores: B too large for one launch");
  if (id_dtype == MM_I32)
    mm::inbatch_scores_kernel<int32_t><<<grid, 256, 0, st>>>(
        q, neg, B, N, D, (const int32_t*)pos_ids, (const int32_t*)neg_ids, downscore,
        false_neg_score, neg_prob, temperature, out, out_stride);
  else
    mm::inbatch_scores_kernel<int64_t><<<grid, 256, 0, st>>>(
        q, neg, B, N, D, (const int64_t*)pos_ids, (const int64_t*)neg_ids, downscore,
        false_neg_score, neg_prob, temperature, out, out_stride);
  return mm::check_launch("mm_inbatch_scores");
}

}  // extern "C"
