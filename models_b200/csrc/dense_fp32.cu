// Exact-fp32 dense layer on CUDA cores: out = act(x @ W + bias), optional DCN-v2 cross
// epilogue out = x0 * (x @ W + bias) + x.  This is the parity anchor for the tensor-core path
// (dense_tc.cu) and the path taken for shapes the tensor-core kernel does not cover.
// Replaces tf.keras.layers.Dense (merlin/models/tf/blocks/mlp.py:275-280) and Cross.call
// (blocks/cross.py:188-202).
#include "mm_common.cuh"

namespace mm {

constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256)
dense_fp32_kernel(const float* __restrict__ x, long long B, int K, long long x_stride,
                  const float* __restrict__ W, const float* __restrict__ bias, int N, int act,
                  const float* __restrict__ x0, long long x0_stride, float* __restrict__ out,
                  long long out_stride) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  const int a_r = tid >> 2, a_k = (tid & 3) * 4;   // A tile: 64 rows x 16 k, 4 k per thread
  const int b_k = tid >> 4, b_n = (tid & 15) * 4;  // B tile: 16 k x 64 n, 4 n per thread
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long m = m0 + a_r;
      const int k = k0 + a_k + u;
      As[a_k + u][a_r] = (m < B && k < K) ? x[m * x_stride + k] : 0.0f;
      const int kk = k0 + b_k, n = n0 + b_n + u;
      Bs[b_k][b_n + u] = (kk < K && n < N) ? __ldg(W + (long long)kk * N + n) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
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
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.0f);
      if (x0) v = __fadd_rn(__fmul_rn(x0[m * x0_stride + n], v), x[m * x_stride + n]);
      else v = apply_act(v, act);
      out[m * out_stride + n] = v;
    }
  }
}

}  // namespace mm

extern "C" int mm_dense_fp32(const float* x, int64_t B, int K, int64_t x_stride, const float* W,
                             const float* bias, int N, int act, const float* x0,
                             int64_t x0_stride, float* out, int64_t out_stride, void* stream) {
  MM_REQUIRE(x && W && out && B >= 0 && K > 0 && N > 0, MM_ERR_ARG,
             "mm_dense_fp32: null pointer or non-positive K/N");
  MM_REQUIRE(x_stride >= K && out_stride >= N, MM_ERR_ARG, "mm_dense_fp32: stride smaller than row");
  MM_REQUIRE(act >= MM_ACT_LINEAR && act <= MM_ACT_GELU, MM_ERR_ARG, "mm_dense_fp32: unknown activation %d", act);
  MM_REQUIRE(!x0 || (N == K && x0_stride >= N), MM_ERR_ARG,
             "mm_dense_fp32: the cross epilogue needs a square kernel (N == K)");
  if (B == 0) return MM_OK;
  dim3 grid((unsigned)((B + mm::BM - 1) / mm::BM), (unsigned)((N + mm::BN - 1) / mm::BN));
  mm::dense_fp32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, B, K, x_stride, W, bias, N, act, x0,
                                                              x0_stride, out, out_stride);
  return mm::check_launch("mm_dense_fp32");
}
