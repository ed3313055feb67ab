// Backward of the Dense layers of the DLRM path (training step): what tf.GradientTape
// (merlin/models/tf/models/base.py:1121-1231, `train_step`) computes for blocks/mlp.py:275-280 and the
// BinaryOutput head (outputs/classification.py:114), written for narrow layers (K, N <= a few hundred) over a huge
// batch: the whole backward is 19 GFLOP against ~1 GB of activations per 65 536 samples, so every kernel
// here is organised around streaming the activations once.
//
//   mm_bce_head_fwd_bwd  z = x.w + b, loss += sum BCE(z, y), dz = (sigmoid(z) - y) * w_i / M,
//                        dx = dz * w [x > 0], dw += x^T dz, db += sum dz        (one warp per row, no GEMM)
//   mm_dense_wgrad       dW += X^T dZ, db += column sums of dZ                  (reduction over the batch)
//   mm_dense_dgrad       dX = (dZ W^T) [mask > 0]                                (mask: the layer input = the
//                        previous layer's relu output, so the result is that layer's pre-activation gradient)
//
// Arithmetic: mma.sync.m16n8k16 bf16 with the library's 3-pass split (hi*lo + lo*hi + hi*hi, fp32 accumulate):
// fp32-grade (|err| ~ 2^-16 relative) like the forward layers.  The batch reduction of wgrad runs over CTAs and ends in
// fp32 atomics: the summation ORDER is not fixed, results agree with autograd to fp32 rounding, not bit for bit.
#include <cuda_bf16.h>

#include <cstring>

#include "mm_common.cuh"

namespace mm {
namespace trn {

__device__ __forceinline__ void split_pair(float x, float y, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  const float xh = __uint_as_float(hi << 16), yh = __uint_as_float(hi & 0xffff0000u);
  __nv_bfloat162 l = __floats2bfloat162_rn(x - xh, y - yh);
  lo = *reinterpret_cast<uint32_t*>(&l);
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}

// ---------------------------------------------------------------------------------------------------------------
// Head: Dense(K -> 1) + sigmoid + binary cross-entropy, forward and backward in one pass over x.
// Keras evaluates BCE on the logits cached by the sigmoid activation (`_keras_logits`):
//   l_i = max(z,0) - z*y + log(1 + exp(-|z|)),   loss = sum_i w_i l_i / sum... (mean over the batch: / M)
// ---------------------------------------------------------------------------------------------------------------
struct HeadParams {
  const float* x;
  long long ldx;
  long long M;
  int K;
  const float* w;
  const float* bias;  // device scalar or null
  const void* y;
  int y_dtype;
  const float* sample_w;  // (M,) or null
  float inv_m;            // 1 / M  (Keras "sum_over_batch_size")
  float* logits;          // (M,) or null
  float* loss;            // device scalar, accumulated
  float* dx;
  long long lddx;
  int mask_relu;
  float* dw;  // (K,) accumulated
  float* db;  // scalar accumulated
};

__device__ __forceinline__ float load_target(const void* y, long long i, int dt) {
  switch (dt) {
    case MM_I32: return (float)reinterpret_cast<const int32_t*>(y)[i];
    case MM_I64: return (float)reinterpret_cast<const long long*>(y)[i];
    case MM_F64: return (float)reinterpret_cast<const double*>(y)[i];
    default: return reinterpret_cast<const float*>(y)[i];
  }
}

constexpr int HEAD_KMAX = 256;  // 8 columns per lane

__global__ void __launch_bounds__(256) head_kernel(const HeadParams p) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  constexpr int C = HEAD_KMAX / 32;
  float w[C], dw[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int k = lane + 32 * c;
    w[c] = k < p.K ? p.w[k] : 0.0f;
    dw[c] = 0.0f;
  }
  const float b = p.bias ? p.bias[0] : 0.0f;
  float loss = 0.0f, db = 0.0f;
  for (long long m = warp; m < p.M; m += n_warps) {
    float x[C];
    float dot = 0.0f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int k = lane + 32 * c;
      x[c] = k < p.K ? p.x[m * p.ldx + k] : 0.0f;
      dot = fmaf(x[c], w[c], dot);
    }
    const float z = warp_sum(dot) + b;
    const float y = load_target(p.y, m, p.y_dtype);
    const float sw = p.sample_w ? p.sample_w[m] : 1.0f;
    const float e = expf(-fabsf(z));
    const float l = fmaxf(z, 0.0f) - z * y + log1pf(e);
    const float sig = z >= 0.0f ? 1.0f / (1.0f + e) : e / (1.0f + e);
    const float dz = (sig - y) * sw * p.inv_m;
    if (lane == 0) {
      loss += l * sw * p.inv_m;
      db += dz;
      if (p.logits) p.logits[m] = z;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int k = lane + 32 * c;
      if (k < p.K) {
        dw[c] = fmaf(x[c], dz, dw[c]);
        if (p.dx) p.dx[m * p.lddx + k] = (!p.mask_relu || x[c] > 0.0f) ? dz * w[c] : 0.0f;
      }
    }
  }
  // block-level reduction of dw / db / loss before the atomics
  __shared__ float red[8][HEAD_KMAX + 2];
  const int wid = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < C; ++c) red[wid][lane + 32 * c] = dw[c];
  if (lane == 0) {
    red[wid][HEAD_KMAX] = db;
    red[wid][HEAD_KMAX + 1] = loss;
  }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  for (int k = threadIdx.x; k < HEAD_KMAX + 2; k += blockDim.x) {
    float s = 0.0f;
    for (int i = 0; i < nw; ++i) s += red[i][k];
    if (k < p.K) {
      if (p.dw) atomicAdd(p.dw + k, s);
    } else if (k == HEAD_KMAX) {
      if (p.db) atomicAdd(p.db, s);
    } else if (k == HEAD_KMAX + 1) {
      if (p.loss) atomicAdd(p.loss, s);
    }
  }
}

// K a multiple of 4 with 16-byte aligned rows: G = K/4 (rounded up to a power of two, <= 32) lanes own one row as
// float4 pieces, a warp processes 32/G rows at a time (K = 32: 4 rows per warp, one 128-byte row per 8 lanes).
template <int G>
__global__ void __launch_bounds__(256) head_kernel_v4(const HeadParams p) {
  const int lane = threadIdx.x & 31;
  const int c = lane % G, sub = lane / G;
  constexpr int RPW = 32 / G;  // rows per warp and iteration
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const bool on = 4 * c < p.K;
  const float4 w = on ? *reinterpret_cast<const float4*>(p.w + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 dw = make_float4(0.f, 0.f, 0.f, 0.f);
  const float b = p.bias ? p.bias[0] : 0.0f;
  float loss = 0.0f, db = 0.0f;
  const long long groups = (p.M + RPW - 1) / RPW;
  for (long long gi = warp; gi < groups; gi += n_warps) {
    const long long m = gi * RPW + sub;
    const bool live = m < p.M;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live && on) x = __ldg(reinterpret_cast<const float4*>(p.x + m * p.ldx + 4 * c));
    float dot = x.x * w.x + x.y * w.y + x.z * w.z + x.w * w.w;
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    const float z = dot + b;
    float dz = 0.0f;
    if (live) {
      const float y = load_target(p.y, m, p.y_dtype);
      const float sw = p.sample_w ? p.sample_w[m] : 1.0f;
      const float e = expf(-fabsf(z));
      const float sig = z >= 0.0f ? 1.0f / (1.0f + e) : e / (1.0f + e);
      dz = (sig - y) * sw * p.inv_m;
      if (c == 0) {
        loss += (fmaxf(z, 0.0f) - z * y + log1pf(e)) * sw * p.inv_m;
        db += dz;
        if (p.logits) p.logits[m] = z;
      }
    }
    dw.x = fmaf(x.x, dz, dw.x);
    dw.y = fmaf(x.y, dz, dw.y);
    dw.z = fmaf(x.z, dz, dw.z);
    dw.w = fmaf(x.w, dz, dw.w);
    if (live && on && p.dx) {
      float4 d;
      d.x = (!p.mask_relu || x.x > 0.0f) ? dz * w.x : 0.0f;
      d.y = (!p.mask_relu || x.y > 0.0f) ? dz * w.y : 0.0f;
      d.z = (!p.mask_relu || x.z > 0.0f) ? dz * w.z : 0.0f;
      d.w = (!p.mask_relu || x.w > 0.0f) ? dz * w.w : 0.0f;
      *reinterpret_cast<float4*>(p.dx + m * p.lddx + 4 * c) = d;
    }
  }
  // lanes with the same c hold partial sums of the same columns: fold the row groups of the warp, then the block
#pragma unroll
  for (int o = G; o < 32; o <<= 1) {
    dw.x += __shfl_xor_sync(0xffffffffu, dw.x, o);
    dw.y += __shfl_xor_sync(0xffffffffu, dw.y, o);
    dw.z += __shfl_xor_sync(0xffffffffu, dw.z, o);
    dw.w += __shfl_xor_sync(0xffffffffu, dw.w, o);
    loss += __shfl_xor_sync(0xffffffffu, loss, o);
    db += __shfl_xor_sync(0xffffffffu, db, o);
  }
  __shared__ float red[8][4 * G + 2];
  const int wid = threadIdx.x >> 5;
  if (lane < G) {
    red[wid][4 * lane] = dw.x;
    red[wid][4 * lane + 1] = dw.y;
    red[wid][4 * lane + 2] = dw.z;
    red[wid][4 * lane + 3] = dw.w;
  }
  if (lane == 0) {
    red[wid][4 * G] = db;
    red[wid][4 * G + 1] = loss;
  }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  for (int k = threadIdx.x; k < 4 * G + 2; k += blockDim.x) {
    float s = 0.0f;
    for (int i = 0; i < nw; ++i) s += red[i][k];
    if (k < 4 * G) {
      if (k < p.K && p.dw) atomicAdd(p.dw + k, s);
    } else if (k == 4 * G) {
      if (p.db) atomicAdd(p.db, s);
    } else if (p.loss) {
      atomicAdd(p.loss, s);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// wgrad:  dW[K, N] += sum_m X[m, :]^T dZ[m, :],  db[N] += sum_m dZ[m, :]
// MMA view: M' = K (rows of dW), N' = N, K' = batch rows.  A CTA owns a (KS x NS) tile of dW and a contiguous slice of
// the batch; 32-row chunks of X and dZ are split into bf16 (hi, lo) while they are stored to shared memory ([m][k]
// row-major, +16 B row padding) and both operands are fetched with ldmatrix.trans (the reduction index m is the slow
// index of both).  Global loads of chunk i+1 are in flight while chunk i is multiplied.
// ---------------------------------------------------------------------------------------------------------------
struct WgradParams {
  const float* x;
  long long ldx;
  const float* dz;
  long long ldz;
  long long M;
  int K, N;
  float* dw;  // (K, N) contiguous, accumulated
  float* db;  // (N,) accumulated or null
  long long rows_per_cta;  // multiple of 32
  int x_vec, z_vec;        // 16-byte loads are legal
  const __nv_bfloat16* x_split;  // XSPLIT: X as split-bf16 rows (M, 2*Kp) = [hi | lo] (mm_split_rows layout) instead of fp32
  int Kp;
};

__device__ __forceinline__ float4 load4(const float* row, int c, int C, bool vec) {
  if (vec && c + 3 < C) return __ldg(reinterpret_cast<const float4*>(row + c));
  float4 v;
  v.x = c < C ? __ldg(row + c) : 0.0f;
  v.y = c + 1 < C ? __ldg(row + c + 1) : 0.0f;
  v.z = c + 2 < C ? __ldg(row + c + 2) : 0.0f;
  v.w = c + 3 < C ? __ldg(row + c + 3) : 0.0f;
  return v;
}

template <int WM, int MT, int WN, int NT, bool XSPLIT>
__global__ void __launch_bounds__(32 * WM * WN) wgrad_kernel(const WgradParams p) {
  constexpr int T = 32 * WM * WN;
  constexpr int KS = WM * MT * 16, NS = WN * NT * 8;
  constexpr int SX = KS * 2 + 16, SZ = NS * 2 + 16;  // bytes per staged row
  constexpr int XQ = KS / 4, ZQ = NS / 4;            // float4 per staged row
  constexpr int XV = (32 * XQ + T - 1) / T, ZV = (32 * ZQ + T - 1) / T;
  static_assert(T % ZQ == 0, "a thread stages a fixed column group of dZ (bias-gradient partial sums)");
  __shared__ __align__(16) uint8_t smem[2 * 32 * SX + 2 * 32 * SZ];
  uint8_t* xh = smem;
  uint8_t* xl = smem + 32 * SX;
  uint8_t* zh = smem + 64 * SX;
  uint8_t* zl = zh + 32 * SZ;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp / WN, wn = warp % WN;
  const int k0 = blockIdx.y * KS, n0 = blockIdx.z * NS;
  const long long m_begin = (long long)blockIdx.x * p.rows_per_cta;
  const long long m_end = min(p.M, m_begin + p.rows_per_cta);
  if (m_begin >= m_end) return;

  float acc[MT][NT][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.0f;
  float4 dbs = make_float4(0.f, 0.f, 0.f, 0.f);

  float4 xr[XV], zr[ZV];
  auto prefetch = [&](long long mb) {
#pragma unroll
    for (int i = 0; i < XV; ++i) {
      const int e = tid + i * T;
      const int r = e / XQ, c = (e % XQ) * 4;
      const long long m = mb + r;
      if (XSPLIT) {
        // four bf16 hi values and four lo values, already split: bit patterns travel through the float4
        uint2 h = make_uint2(0u, 0u), l = h;
        if (e < 32 * XQ && m < m_end && k0 + c < p.Kp) {
          const __nv_bfloat16* row = p.x_split + m * (2ll * p.Kp) + k0 + c;
          h = __ldg(reinterpret_cast<const uint2*>(row));
          l = __ldg(reinterpret_cast<const uint2*>(row + p.Kp));
        }
        xr[i] = make_float4(__uint_as_float(h.x), __uint_as_float(h.y), __uint_as_float(l.x), __uint_as_float(l.y));
      } else {
        xr[i] = (e < 32 * XQ && m < m_end) ? load4(p.x + m * p.ldx, k0 + c, p.K, p.x_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int i = 0; i < ZV; ++i) {
      const int e = tid + i * T;
      const int r = e / ZQ, c = (e % ZQ) * 4;
      const long long m = mb + r;
      zr[i] = (e < 32 * ZQ && m < m_end) ? load4(p.dz + m * p.ldz, n0 + c, p.N, p.z_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int i = 0; i < XV; ++i) {
      const int e = tid + i * T;
      if (e < 32 * XQ) {
        const int r = e / XQ, c = (e % XQ) * 4;
        uint32_t h0, l0, h1, l1;
        if (XSPLIT) {
          h0 = __float_as_uint(xr[i].x);
          h1 = __float_as_uint(xr[i].y);
          l0 = __float_as_uint(xr[i].z);
          l1 = __float_as_uint(xr[i].w);
        } else {
          split_pair(xr[i].x, xr[i].y, h0, l0);
          split_pair(xr[i].z, xr[i].w, h1, l1);
        }
        *reinterpret_cast<uint2*>(xh + r * SX + c * 2) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(xl + r * SX + c * 2) = make_uint2(l0, l1);
      }
    }
#pragma unroll
    for (int i = 0; i < ZV; ++i) {
      const int e = tid + i * T;
      if (e < 32 * ZQ) {
        const int r = e / ZQ, c = (e % ZQ) * 4;
        uint32_t h0, l0, h1, l1;
        split_pair(zr[i].x, zr[i].y, h0, l0);
        split_pair(zr[i].z, zr[i].w, h1, l1);
        *reinterpret_cast<uint2*>(zh + r * SZ + c * 2) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(zl + r * SZ + c * 2) = make_uint2(l0, l1);
        dbs.x += zr[i].x;
        dbs.y += zr[i].y;
        dbs.z += zr[i].z;
        dbs.w += zr[i].w;
      }
    }
  };

  // ldmatrix row addresses of this lane (see the fragment maps in the file header of tower_small.cu):
  //   A (= X^T) quad of m-tile mt, k-step ks: matrices (kk 0-7, i 0-7) (kk 0-7, i 8-15) (kk 8-15, i 0-7) (kk 8-15, i 8-15)
  //   B (= dZ) pairs of n-tiles 2u, 2u+1:     matrices (kk 0-7, 2u) (kk 8-15, 2u) (kk 0-7, 2u+1) (kk 8-15, 2u+1)
  const int q = lane >> 3, r8 = lane & 7;
  const uint32_t xs_base = (uint32_t)__cvta_generic_to_shared(xh), zs_base = (uint32_t)__cvta_generic_to_shared(zh);
  const uint32_t a_lane = (uint32_t)(((q >> 1) * 8 + r8) * SX + (wm * MT * 16 + (q & 1) * 8) * 2);
  const uint32_t b_lane = (uint32_t)(((q & 1) * 8 + r8) * SZ + (wn * NT * 8 + (NT > 1 ? (q >> 1) * 8 : 0)) * 2);

  prefetch(m_begin);
  for (long long mb = m_begin; mb < m_end; mb += 32) {
    __syncthreads();  // the previous chunk has been consumed
    stage();
    __syncthreads();
    if (mb + 32 < m_end) prefetch(mb + 32);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t ah[MT][4], al[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const uint32_t a = xs_base + a_lane + (uint32_t)(ks * 16 * SX + mt * 32);
        ldsm_x4_t(a, ah[mt][0], ah[mt][1], ah[mt][2], ah[mt][3]);
        ldsm_x4_t(a + 32 * SX, al[mt][0], al[mt][1], al[mt][2], al[mt][3]);
      }
      constexpr int NP = (NT + 1) / 2;
#pragma unroll
      for (int u = 0; u < NP; ++u) {
        uint32_t bh[4], bl[4];
        const uint32_t b = zs_base + b_lane + (uint32_t)(ks * 16 * SZ + u * 32);
        ldsm_x4_t(b, bh[0], bh[1], bh[2], bh[3]);
        ldsm_x4_t(b + 32 * SZ, bl[0], bl[1], bl[2], bl[3]);
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const int nt = 2 * u + v;
          if (nt < NT) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) mma16816(acc[mt][nt], ah[mt], bl[2 * v], bl[2 * v + 1]);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) mma16816(acc[mt][nt], al[mt], bh[2 * v], bh[2 * v + 1]);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) mma16816(acc[mt][nt], ah[mt], bh[2 * v], bh[2 * v + 1]);
          }
        }
      }
    }
  }
  // ---- this CTA's partial tile -> global (fp32 atomics)
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int k = k0 + (wm * MT + mt) * 16 + g + 8 * (c >> 1);
        const int n = n0 + (wn * NT + nt) * 8 + 2 * t + (c & 1);
        if (k < p.K && n < p.N) atomicAdd(p.dw + (long long)k * p.N + n, acc[mt][nt][c]);
      }
  if (p.db && blockIdx.y == 0) {
    // thread tid stages column group (tid % ZQ): reduce the T / ZQ threads of a group through shared memory
    __syncthreads();
    float* red = reinterpret_cast<float*>(smem);
    for (int i = tid; i < NS; i += T) red[i] = 0.0f;
    __syncthreads();
    const int c = (tid % ZQ) * 4;
    if (tid < 32 * ZQ || ZV > 1) {
      atomicAdd(red + c, dbs.x);
      atomicAdd(red + c + 1, dbs.y);
      atomicAdd(red + c + 2, dbs.z);
      atomicAdd(red + c + 3, dbs.w);
    }
    __syncthreads();
    for (int i = tid; i < NS; i += T)
      if (n0 + i < p.N) atomicAdd(p.db + n0 + i, red[i]);
  }
}

template <int WM, int MT, int WN, int NT>
static int launch_wgrad(WgradParams p, cudaStream_t st) {
  constexpr int KS = WM * MT * 16, NS = WN * NT * 8;
  const int ky = (p.K + KS - 1) / KS, nz = (p.N + NS - 1) / NS;
  long long ctas = std::max(1LL, 2LL * sm_count() / ((long long)ky * nz));
  long long rows = (p.M + ctas - 1) / ctas;
  rows = (rows + 31) / 32 * 32;
  if (rows < 256) rows = 256;
  p.rows_per_cta = rows;
  const long long gx = (p.M + rows - 1) / rows;
  if (p.x_split) wgrad_kernel<WM, MT, WN, NT, true><<<dim3((unsigned)gx, (unsigned)ky, (unsigned)nz), 32 * WM * WN, 0, st>>>(p);
  else wgrad_kernel<WM, MT, WN, NT, false><<<dim3((unsigned)gx, (unsigned)ky, (unsigned)nz), 32 * WM * WN, 0, st>>>(p);
  return check_launch("mm_dense_wgrad");
}

// ---------------------------------------------------------------------------------------------------------------
// dgrad:  dX[M, K] = dZ[M, N] W^T  (optionally zeroed where mask <= 0).   N <= 128.
// MMA view: M' = batch rows (one warp per 16 rows), N' = K, K' = N.  B[k' = n][n' = k] = W[k][n]: the Keras kernel is
// already "n'-major with k' contiguous", so a 128-row slab of W sits in shared memory as split-bf16 rows [hi | lo] and B
// fragments come from plain ldmatrix.  The A fragments (dZ rows, split once) stay in registers for the whole slab.
// ---------------------------------------------------------------------------------------------------------------
struct DgradParams {
  const float* dz;
  long long ldz;
  const float* w;  // (K, N) contiguous
  const float* mask;
  long long ldmask;
  float* dx;
  long long lddx;
  long long M;
  int K, N;
  int z_vec2, x_vec2;  // 8-byte loads of dz / 8-byte accesses of mask and dx are legal
  int x_vec4;          // 16-byte accesses of mask and dx are legal: the output leaves through a shared-memory stage
};

constexpr int DG_SLAB = 128;  // output columns (rows of W) per CTA
constexpr int DG_WARPS = 8;
constexpr int DG_STG = 72;  // floats per row of a warp's (16 x 64) output stage (+8: conflict-free 64-bit writes per half warp)

template <int KSTEPS>
__global__ void __launch_bounds__(32 * DG_WARPS, 2) dgrad_kernel(const DgradParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int NP = 16 * KSTEPS;        // padded N
  constexpr int SW = NP * 4 + 16;        // bytes per staged W row: [hi 0..NP | lo 0..NP] + pad
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int k0 = blockIdx.y * DG_SLAB;
  // ---- W slab -> shared memory (split on the fly)
  for (int e = tid; e < DG_SLAB * (NP / 2); e += blockDim.x) {
    const int r = e / (NP / 2), c = (e % (NP / 2)) * 2;
    const int k = k0 + r;
    float a = 0.0f, b = 0.0f;
    if (k < p.K) {
      if (c < p.N) a = __ldg(p.w + (long long)k * p.N + c);
      if (c + 1 < p.N) b = __ldg(p.w + (long long)k * p.N + c + 1);
    }
    uint32_t h, l;
    split_pair(a, b, h, l);
    *reinterpret_cast<uint32_t*>(smem + r * SW + c * 2) = h;
    *reinterpret_cast<uint32_t*>(smem + r * SW + NP * 2 + c * 2) = l;
  }
  __syncthreads();
  // matrices (hi, k' 0-7) (hi, k' 8-15) (lo, k' 0-7) (lo, k' 8-15) of W rows 8*nt + (lane & 7)
  const uint32_t w_lane = (uint32_t)__cvta_generic_to_shared(smem) + (uint32_t)(lane & 7) * SW + (uint32_t)((lane >> 3) & 1) * 16u +
                          (uint32_t)(lane >> 4) * (NP * 2);
  const long long tiles = (p.M + 15) >> 4;
  for (long long tile = (long long)blockIdx.x * DG_WARPS + warp; tile < tiles; tile += (long long)gridDim.x * DG_WARPS) {
    const long long r0 = tile * 16 + g, r1 = r0 + 8;
    const bool v0 = r0 < p.M, v1 = r1 < p.M;
    uint32_t ah[KSTEPS][4], al[KSTEPS][4];
    const float* z0 = p.dz + r0 * p.ldz;
    const float* z1 = p.dz + r1 * p.ldz;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
      float2 x00 = make_float2(0.f, 0.f), x01 = x00, x10 = x00, x11 = x00;  // (row, k-half)
      const int c0 = 16 * ks + 2 * t, c1 = c0 + 8;
      if (p.z_vec2 && c1 + 1 < p.N) {
        if (v0) {
          x00 = __ldg(reinterpret_cast<const float2*>(z0 + c0));
          x01 = __ldg(reinterpret_cast<const float2*>(z0 + c1));
        }
        if (v1) {
          x10 = __ldg(reinterpret_cast<const float2*>(z1 + c0));
          x11 = __ldg(reinterpret_cast<const float2*>(z1 + c1));
        }
      } else {
        if (v0) {
          if (c0 < p.N) x00.x = __ldg(z0 + c0);
          if (c0 + 1 < p.N) x00.y = __ldg(z0 + c0 + 1);
          if (c1 < p.N) x01.x = __ldg(z0 + c1);
          if (c1 + 1 < p.N) x01.y = __ldg(z0 + c1 + 1);
        }
        if (v1) {
          if (c0 < p.N) x10.x = __ldg(z1 + c0);
          if (c0 + 1 < p.N) x10.y = __ldg(z1 + c0 + 1);
          if (c1 < p.N) x11.x = __ldg(z1 + c1);
          if (c1 + 1 < p.N) x11.y = __ldg(z1 + c1 + 1);
        }
      }
      split_pair(x00.x, x00.y, ah[ks][0], al[ks][0]);
      split_pair(x10.x, x10.y, ah[ks][1], al[ks][1]);
      split_pair(x01.x, x01.y, ah[ks][2], al[ks][2]);
      split_pair(x11.x, x11.y, ah[ks][3], al[ks][3]);
    }
#pragma unroll 1
    for (int ch = 0; ch < DG_SLAB / 64; ++ch) {  // 64 output columns at a time
      if (k0 + ch * 64 >= p.K) break;
      float acc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f;
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
        for (int n4 = 0; n4 < 8; n4 += 4) {
          uint32_t bh[4][2], bl[4][2];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            ldsm_x4(w_lane + (uint32_t)((ch * 8 + n4 + u) * 8 * SW + ks * 32), bh[u][0], bh[u][1], bl[u][0], bl[u][1]);
#pragma unroll
          for (int u = 0; u < 4; ++u) mma16816(acc[n4 + u], ah[ks], bl[u][0], bl[u][1]);
#pragma unroll
          for (int u = 0; u < 4; ++u) mma16816(acc[n4 + u], al[ks], bh[u][0], bh[u][1]);
#pragma unroll
          for (int u = 0; u < 4; ++u) mma16816(acc[n4 + u], ah[ks], bh[u][0], bh[u][1]);
        }
      }
      if (p.x_vec4) {
        // stage the (16 x 64) tile in shared memory and leave as full rows: 16 lanes x 16 bytes = one 256-byte row segment
        // (the fragment layout gives every lane 8-byte pieces of 8 different rows: 1.1 us per MB of output, measured)
        float* stg = reinterpret_cast<float*>(smem + DG_SLAB * SW) + warp * (16 * DG_STG);
        __syncwarp();
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          *reinterpret_cast<float2*>(stg + g * DG_STG + nt * 8 + 2 * t) = make_float2(acc[nt][0], acc[nt][1]);
          *reinterpret_cast<float2*>(stg + (g + 8) * DG_STG + nt * 8 + 2 * t) = make_float2(acc[nt][2], acc[nt][3]);
        }
        __syncwarp();
        const int c4 = (lane & 15) * 4;
        const int kc = k0 + ch * 64 + c4;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = it * 2 + (lane >> 4);
          const long long r = tile * 16 + rr;
          if (r < p.M && kc < p.K) {
            float4 v = *reinterpret_cast<const float4*>(stg + rr * DG_STG + c4);
            if (kc + 3 < p.K) {
              if (p.mask) {
                const float4 mv = __ldg(reinterpret_cast<const float4*>(p.mask + r * p.ldmask + kc));
                v.x = mv.x > 0.0f ? v.x : 0.0f;
                v.y = mv.y > 0.0f ? v.y : 0.0f;
                v.z = mv.z > 0.0f ? v.z : 0.0f;
                v.w = mv.w > 0.0f ? v.w : 0.0f;
              }
              *reinterpret_cast<float4*>(p.dx + r * p.lddx + kc) = v;
            } else {  // the last, partial group of a row (K not a multiple of 4)
              const float vv[4] = {v.x, v.y, v.z, v.w};
              for (int e = 0; e < 4 && kc + e < p.K; ++e) {
                float a = vv[e];
                if (p.mask) a = __ldg(p.mask + r * p.ldmask + kc + e) > 0.0f ? a : 0.0f;
                p.dx[r * p.lddx + kc + e] = a;
              }
            }
          }
        }
        continue;
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int k = k0 + ch * 64 + nt * 8 + 2 * t;
        if (k >= p.K) continue;
        const bool two = k + 1 < p.K;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const long long r = h ? r1 : r0;
          if (!(h ? v1 : v0)) continue;
          float a = acc[nt][2 * h], b = acc[nt][2 * h + 1];
          if (p.mask) {
            const float* mrow = p.mask + r * p.ldmask + k;
            if (p.x_vec2 && two) {
              const float2 mv = __ldg(reinterpret_cast<const float2*>(mrow));
              a = mv.x > 0.0f ? a : 0.0f;
              b = mv.y > 0.0f ? b : 0.0f;
            } else {
              a = __ldg(mrow) > 0.0f ? a : 0.0f;
              if (two) b = __ldg(mrow + 1) > 0.0f ? b : 0.0f;
            }
          }
          float* drow = p.dx + r * p.lddx + k;
          if (p.x_vec2 && two) {
            *reinterpret_cast<float2*>(drow) = make_float2(a, b);
          } else {
            drow[0] = a;
            if (two) drow[1] = b;
          }
        }
      }
    }
  }
}

template <int KSTEPS>
static int launch_dgrad(const DgradParams& p, cudaStream_t st) {
  constexpr int NP = 16 * KSTEPS;
  const size_t smem = (size_t)DG_SLAB * (NP * 4 + 16) + (size_t)DG_WARPS * 16 * DG_STG * sizeof(float);
  auto kern = dgrad_kernel<KSTEPS>;
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (smem > 48 * 1024 && (dev < 0 || dev >= 64 || !attr_set[dev])) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("mm_dense_dgrad: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int ky = (p.K + DG_SLAB - 1) / DG_SLAB;
  const long long tiles = (p.M + 15) / 16;
  long long gx = (tiles + DG_WARPS - 1) / DG_WARPS;
  const long long cap = std::max(1LL, 2LL * sm_count() / ky);
  if (gx > cap) gx = cap;
  kern<<<dim3((unsigned)gx, (unsigned)ky), 32 * DG_WARPS, smem, st>>>(p);
  return check_launch("mm_dense_dgrad");
}

}  // namespace trn
}  // namespace mm

extern "C" {

int mm_bce_head_fwd_bwd(const float* x, int64_t M, int K, int64_t x_stride, const float* w, const float* bias, const void* targets,
                        int target_dtype, const float* sample_weight, float* logits, float* loss_sum, float* dx,
                        int64_t dx_stride, int mask_relu, float* dw, float* db, void* stream) {
  using namespace mm::trn;
  MM_REQUIRE(x && w && targets && M >= 0 && K >= 1 && x_stride >= K, MM_ERR_ARG, "mm_bce_head_fwd_bwd: null pointer or bad K / stride");
  MM_REQUIRE(K <= HEAD_KMAX, MM_ERR_UNSUPPORTED, "mm_bce_head_fwd_bwd: K=%d > %d", K, HEAD_KMAX);
  MM_REQUIRE(target_dtype >= MM_I32 && target_dtype <= MM_F64, MM_ERR_ARG, "mm_bce_head_fwd_bwd: bad target dtype");
  MM_REQUIRE(!dx || dx_stride >= K, MM_ERR_ARG, "mm_bce_head_fwd_bwd: dx_stride < K");
  if (M == 0) return MM_OK;
  HeadParams p;
  memset(&p, 0, sizeof(p));
  p.x = x;
  p.ldx = x_stride;
  p.M = M;
  p.K = K;
  p.w = w;
  p.bias = bias;
  p.y = targets;
  p.y_dtype = target_dtype;
  p.sample_w = sample_weight;
  p.inv_m = 1.0f / (float)M;
  p.logits = logits;
  p.loss = loss_sum;
  p.dx = dx;
  p.lddx = dx_stride;
  p.mask_relu = mask_relu;
  p.dw = dw;
  p.db = db;
  long long blocks = (M + 7) / 8;
  const long long cap = 4LL * mm::sm_count();
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (K & 3) == 0 && K <= 128 && (x_stride & 3) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0 &&
                   (!dx || ((dx_stride & 3) == 0 && ((uintptr_t)dx & 15) == 0));
  if (vec) {
    const int g = K / 4;
    if (g <= 2) head_kernel_v4<2><<<(unsigned)blocks, 256, 0, st>>>(p);
    else if (g <= 4) head_kernel_v4<4><<<(unsigned)blocks, 256, 0, st>>>(p);
    else if (g <= 8) head_kernel_v4<8><<<(unsigned)blocks, 256, 0, st>>>(p);
    else if (g <= 16) head_kernel_v4<16><<<(unsigned)blocks, 256, 0, st>>>(p);
    else head_kernel_v4<32><<<(unsigned)blocks, 256, 0, st>>>(p);
  } else {
    head_kernel<<<(unsigned)blocks, 256, 0, st>>>(p);
  }
  return mm::check_launch("mm_bce_head_fwd_bwd");
}

static int wgrad_dispatch(mm::trn::WgradParams p, void* stream) {
  using namespace mm::trn;
  cudaStream_t st = (cudaStream_t)stream;
  const int K = p.K, N = p.N;
  const int ks = K > 64 ? 128 : K > 16 ? 64 : 16;
  const int ns = N > 64 ? 128 : N > 32 ? 64 : 32;
#define MM_WG(KS_, NS_, WM, MT, WN, NT) \
  if (ks == KS_ && ns == NS_) return launch_wgrad<WM, MT, WN, NT>(p, st);
  MM_WG(128, 128, 4, 2, 4, 4) MM_WG(128, 64, 4, 2, 2, 4) MM_WG(128, 32, 4, 2, 2, 2)
  MM_WG(64, 128, 2, 2, 4, 4) MM_WG(64, 64, 2, 2, 4, 2) MM_WG(64, 32, 4, 1, 2, 2)
  MM_WG(16, 128, 1, 1, 8, 2) MM_WG(16, 64, 1, 1, 8, 1) MM_WG(16, 32, 1, 1, 4, 1)
#undef MM_WG
  return MM_ERR_UNSUPPORTED;
}

int mm_dense_wgrad(const float* x, int64_t M, int K, int64_t x_stride, const float* dz, int N, int64_t dz_stride, float* dw,
                   float* db, void* stream) {
  using namespace mm::trn;
  MM_REQUIRE(x && dz && dw && M >= 0 && K >= 1 && N >= 1 && x_stride >= K && dz_stride >= N, MM_ERR_ARG,
             "mm_dense_wgrad: null pointer or bad shape / stride");
  if (M == 0) return MM_OK;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.x = x;
  p.ldx = x_stride;
  p.dz = dz;
  p.ldz = dz_stride;
  p.M = M;
  p.K = K;
  p.N = N;
  p.dw = dw;
  p.db = db;
  p.x_vec = ((x_stride & 3) == 0 && ((uintptr_t)x & 15) == 0) ? 1 : 0;
  p.z_vec = ((dz_stride & 3) == 0 && ((uintptr_t)dz & 15) == 0) ? 1 : 0;
  return wgrad_dispatch(p, stream);
}

int mm_dense_wgrad_split(const void* x_split, int64_t M, int K, int Kp, const float* dz, int N, int64_t dz_stride, float* dw,
                         float* db, void* stream) {
  using namespace mm::trn;
  MM_REQUIRE(x_split && dz && dw && M >= 0 && K >= 1 && N >= 1 && dz_stride >= N, MM_ERR_ARG, "mm_dense_wgrad_split: null pointer or bad shape");
  MM_REQUIRE(Kp == mm_tc_padded_k(K) && ((uintptr_t)x_split & 15) == 0, MM_ERR_ARG,
             "mm_dense_wgrad_split: Kp must be mm_tc_padded_k(K)=%d and x_split 16-byte aligned", mm_tc_padded_k(K));
  if (M == 0) return MM_OK;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.x_split = (const __nv_bfloat16*)x_split;
  p.Kp = Kp;
  p.dz = dz;
  p.ldz = dz_stride;
  p.M = M;
  p.K = K;
  p.N = N;
  p.dw = dw;
  p.db = db;
  p.z_vec = ((dz_stride & 3) == 0 && ((uintptr_t)dz & 15) == 0) ? 1 : 0;
  return wgrad_dispatch(p, stream);
}

int mm_dense_dgrad(const float* dz, int64_t M, int N, int64_t dz_stride, const float* w, int K, const float* mask,
                   int64_t mask_stride, float* dx, int64_t dx_stride, void* stream) {
  using namespace mm::trn;
  MM_REQUIRE(dz && w && dx && M >= 0 && K >= 1 && N >= 1 && dz_stride >= N && dx_stride >= K, MM_ERR_ARG,
             "mm_dense_dgrad: null pointer or bad shape / stride");
  MM_REQUIRE(!mask || mask_stride >= K, MM_ERR_ARG, "mm_dense_dgrad: mask_stride < K");
  MM_REQUIRE(N <= 128, MM_ERR_UNSUPPORTED, "mm_dense_dgrad: N=%d > 128 (run the forward GEMM on the transposed kernel instead)", N);
  if (M == 0) return MM_OK;
  DgradParams p;
  memset(&p, 0, sizeof(p));
  p.dz = dz;
  p.ldz = dz_stride;
  p.w = w;
  p.mask = mask;
  p.ldmask = mask_stride;
  p.dx = dx;
  p.lddx = dx_stride;
  p.M = M;
  p.K = K;
  p.N = N;
  p.z_vec2 = ((dz_stride & 1) == 0 && ((uintptr_t)dz & 7) == 0) ? 1 : 0;
  p.x_vec2 = ((dx_stride & 1) == 0 && ((uintptr_t)dx & 7) == 0 && (!mask || ((mask_stride & 1) == 0 && ((uintptr_t)mask & 7) == 0))) ? 1 : 0;
  p.x_vec4 = ((dx_stride & 3) == 0 && ((uintptr_t)dx & 15) == 0 && (!mask || ((mask_stride & 3) == 0 && ((uintptr_t)mask & 15) == 0))) ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (N <= 32) return launch_dgrad<2>(p, st);
  if (N <= 64) return launch_dgrad<4>(p, st);
  return launch_dgrad<8>(p, st);
}

}  // extern "C"
