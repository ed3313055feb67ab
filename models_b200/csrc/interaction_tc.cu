// DLRM pairwise interaction on tcgen05: FOUR samples per 128x128 UMMA tile.
//
// The per-sample Gram matrix X X^T (X = F x 64, F <= 32) occupies a 32x32 diagonal block of a
// 128x128 accumulator when four samples are stacked into one 128-row operand tile; A and B are the
// SAME shared-memory tile (B = X^T), so one bf16 tile pair (hi, lo) per quad feeds the three split
// passes.  3/4 of the tensor work is wasted on cross-sample blocks, but the tensor pipe has ~10x
// headroom here (12 UMMAs of 64 cycles per quad = 192 cycles per sample, ~45 us per launch), and in
// exchange the per-sample CUDA-core work drops ~2.7x versus the warp-level mma.sync version
// (interaction_mma.cu: fragment loads, per-fragment splits and 72 HMMAs per sample made that kernel
// instruction-issue bound at ~1 250 warp-instructions per sample).
//
// Warp roles (13 warps, one CTA per SM, persistent over quads):
//   warps 0-7  load + convert: two groups of four; warp (g, w) owns sample w of every quad with
//              parity g: 16-byte cp.async of the 27 rows (embedding tables or stacked input) into
//              its private fp32 staging (2 buffers, index prefetch one quad ahead), then fp32 ->
//              split-bf16 conversion straight into rows 32w.. of the quad's SWIZZLE_128B operand tile
//   warp  8    MMA issuer: 12 x tcgen05.mma (hi*lo, lo*hi, hi*hi) per quad into a double-buffered
//              128-column TMEM accumulator; tcgen05.commit releases the operand tile / publishes the
//              accumulator
//   warps 9-12 epilogue: warp w reads ITS sample's diagonal block (TMEM lanes 32w.., columns 32w..)
//              with one tcgen05.ld, writes the upper triangle (contiguous per row) + the bottom-MLP
//              prefix into an fp32 staging row, and stores it coalesced — as fp32 or directly as the
//              split-bf16 operand of the top MLP's first tensor-core layer.
//
// Replaces: StackFeatures + DotProductInteraction + shortcut concat
// (merlin/models/tf/core/aggregation.py:101-108, blocks/interaction.py:86-116, blocks/dlrm.py:126-130).
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace mm {
namespace itc {

using namespace mm::tc;

constexpr int D = 64;                     // embedding dim handled by this path
constexpr int V = D / 4;                  // float4 per row
constexpr int kLoadWarps = 8, kEpiWarps = 8;  // epilogue: two warps per TMEM lane quarter, alternating quads
constexpr int kThreads = 32 * (kLoadWarps + 1 + kEpiWarps);
constexpr uint32_t TILE_BYTES = 128 * 128;  // 128 rows x 64 bf16
constexpr uint32_t OPERAND_BYTES = 2 * TILE_BYTES;  // hi + lo

struct Params {
  const float* x;  // MODE 0: stacked (B, F, D)
  long long x_stride;
  const float* prefix;  // bottom vector (P == 0 or D)
  long long prefix_stride;
  int P, bottom_slot;
  long long B;
  int F, T, rows;
  float* out_f32;
  long long out_stride;
  __nv_bfloat16* out_split;
  int out_Kp;
  int* oob_count;
  unsigned in_bytes;     // fp32 staging bytes per sample buffer
  unsigned stage_bytes;  // output staging bytes per epilogue warp
};

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void split_pair(float x, float y, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  const float xh = __uint_as_float(hi << 16), yh = __uint_as_float(hi & 0xffff0000u);
  __nv_bfloat162 l = __floats2bfloat162_rn(x - xh, y - yh);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

template <int MODE, typename IdxT>
__global__ void __launch_bounds__(kThreads, 1)
interact_tc_kernel(const __grid_constant__ GatherParams gp, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // pointer arithmetic keeps the shared state space (LDS/STS, not generic LD/ST)
  uint8_t* operand = smem;                                       // [2][hi 16 KB | lo 16 KB]
  uint8_t* in_stage = operand + 2 * OPERAND_BYTES;               // [8 warps][2][in_bytes]
  uint8_t* out_stage = in_stage + (size_t)kLoadWarps * 2 * p.in_bytes;  // [4 warps][stage_bytes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + (size_t)kEpiWarps * p.stage_bytes);
  uint64_t* operand_full = bars;       // [2] count 4 (one arrive per load warp of the group)
  uint64_t* operand_empty = bars + 2;  // [2] count 1 (tcgen05.commit)
  uint64_t* tmem_full = bars + 4;      // [2] count 1 (tcgen05.commit)
  uint64_t* tmem_empty = bars + 6;     // [2] count 4 (one arrive per epilogue warp)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int F = p.F;
  const long long n_quads_total = (p.B + 3) >> 2;
  // this CTA's quads: blockIdx.x, blockIdx.x + gridDim.x, ...
  const long long nq = blockIdx.x < n_quads_total ? (n_quads_total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == kLoadWarps && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(operand_full + i), 4);
      mbar_init(smem_u32(operand_empty + i), 1);
      mbar_init(smem_u32(tmem_full + i), 1);
      mbar_init(smem_u32(tmem_empty + i), 4);  // the four epilogue warps of that quad parity
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kLoadWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp > kLoadWarps) {  // zero the output staging once (padding columns stay zero)
    uint32_t* os = reinterpret_cast<uint32_t*>(out_stage + (size_t)(warp - kLoadWarps - 1) * p.stage_bytes);
    for (int i = lane; i < (int)(p.stage_bytes >> 2); i += 32) os[i] = 0u;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp < kLoadWarps) {
    // ===================== load + convert =====================
    const int g = warp >> 2, w = warp & 3;  // quad parity group, sample slot inside the quad
    uint8_t* my_stage = in_stage + (size_t)warp * 2 * p.in_bytes;
    const long long my_nq = nq > g ? (nq - g + 1) >> 1 : 0;  // quads k = g, g+2, ...

    // per-lane row constants (lane r owns staged row r)
    int my_slot = lane;
    const float* my_base = nullptr;
    long long my_rows = 0;
    const IdxT* my_idx_ptr = nullptr;
    if (MODE == 1 && lane < p.T) {
      my_slot = gp.t[lane].out_col >> 6;
      my_base = gp.t[lane].weights;
      my_rows = gp.t[lane].rows;
      my_idx_ptr = reinterpret_cast<const IdxT*>(gp.t[lane].indices);
    } else if (MODE == 1 && lane == p.T) {
      my_slot = p.bottom_slot;
    }
    const uint32_t my_dst = (uint32_t)(my_slot * D * 4);

    auto sample_of = [&](long long j) -> long long {  // sample index of this warp's j-th quad
      return ((long long)blockIdx.x + (2 * j + g) * (long long)gridDim.x) * 4 + w;
    };
    auto load_index = [&](long long j) -> IdxT {
      if (MODE == 1 && lane < p.T && j < my_nq) {
        const long long s = sample_of(j);
        if (s < p.B) return my_idx_ptr[s];
      }
      return (IdxT)0;
    };
    auto issue = [&](long long j, IdxT idx_raw) {
      if (j < my_nq) {
        const long long s = sample_of(j);
        if (s < p.B) {
          const uint32_t xs_u32 = smem_u32(my_stage + (size_t)(j & 1) * p.in_bytes);
          const float* my_src = my_base;
          uint32_t my_bytes = 16;
          if (MODE == 1) {
            if (lane < p.T) {
              const long long idx = (long long)idx_raw;
              if (idx >= 0 && idx < my_rows) my_src = my_base + (idx << 6);
              else {
                my_bytes = 0;
                if (p.oob_count) atomicAdd(p.oob_count, 1);
              }
            } else {
              my_src = p.prefix + s * p.prefix_stride;
            }
          } else {
            my_src = p.x + s * p.x_stride + ((long long)lane << 6);
          }
          const uint32_t src_lo = (uint32_t)(uintptr_t)my_src, src_hi = (uint32_t)((uintptr_t)my_src >> 32);
          const uint32_t my_info = my_dst | (my_bytes << 24);
          const int v = lane & (V - 1), rsub = lane >> 4;
#pragma unroll 4
          for (int r0 = 0; r0 < p.rows; r0 += 2) {
            const int r = r0 + rsub;
            const bool act = r < p.rows;
            const int rr = act ? r : 0;
            const uint32_t lo = __shfl_sync(0xffffffffu, src_lo, rr);
            const uint32_t hi = __shfl_sync(0xffffffffu, src_hi, rr);
            const uint32_t info = __shfl_sync(0xffffffffu, my_info, rr);
            const float* src = reinterpret_cast<const float*>(((uintptr_t)hi << 32) | lo) + v * 4;
            if (act) cp_async16_zfill(xs_u32 + (info & 0x00ffffffu) + (uint32_t)(v << 4), src, info >> 24);
          }
        }
      }
      cp_async_commit();
    };

    IdxT idx_pref = load_index(0);
    {
      const IdxT cur = idx_pref;
      idx_pref = load_index(1);
      issue(0, cur);
    }
    for (long long j = 0; j < my_nq; ++j) {
      {
        const IdxT cur = idx_pref;
        idx_pref = load_index(j + 2);
        issue(j + 1, cur);  // the other staging buffer was converted in the previous iteration
      }
      cp_async_wait<1>();  // quad j's rows have landed
      __syncwarp();
      mbar_wait(smem_u32(operand_empty + g), (uint32_t)((j & 1) ^ 1));  // MMAs of the previous use retired
      // fp32 rows -> split-bf16 rows 32w.. of the SWIZZLE_128B operand tiles (hi, lo)
      const float* xs = reinterpret_cast<const float*>(my_stage + (size_t)(j & 1) * p.in_bytes);
      uint8_t* op_hi = operand + (size_t)g * OPERAND_BYTES;
      uint8_t* op_lo = op_hi + TILE_BYTES;
      const int total = F * V;
      // four float4 per lane per trip: the shared-memory loads of a trip are all in flight before the
      // first conversion (a single warp has no other way to hide the ~30-cycle LDS latency)
      for (int e0 = 0; e0 < total; e0 += 128) {
        float4 f[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = e0 + u * 32 + lane;
          f[u] = e < total ? *reinterpret_cast<const float4*>(xs + e * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = e0 + u * 32 + lane;
          if (e < total) {
            const int r = e >> 4, c4 = e & 15;  // row of the stack, float4 within the row
            uint32_t h0, l0, h1, l1;
            split_pair(f[u].x, f[u].y, h0, l0);
            split_pair(f[u].z, f[u].w, h1, l1);
            const int R = 32 * w + r;
            const uint32_t off = (uint32_t)(R * 128 + (((c4 >> 1) ^ (R & 7)) << 4) + ((c4 & 1) << 3));
            *reinterpret_cast<uint2*>(op_hi + off) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(op_lo + off) = make_uint2(l0, l1);
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> visible to the MMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(operand_full + g));
    }
    cp_async_wait<0>();
  } else if (warp == kLoadWarps) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, 128);
      for (long long k = 0; k < nq; ++k) {
        const int g = (int)(k & 1);
        const uint32_t ph = (uint32_t)((k >> 1) & 1);
        mbar_wait(smem_u32(operand_full + g), ph);
        mbar_wait(smem_u32(tmem_empty + g), ph ^ 1);
        tcgen05_fence_after();
        const uint32_t hi = smem_u32(operand + (size_t)g * OPERAND_BYTES), lo = hi + TILE_BYTES;
        const uint32_t d_tmem = tmem_base + (uint32_t)(g * 128);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_bf16(d_tmem, make_desc_sw128(hi + ks * 32), make_desc_sw128(lo + ks * 32), idesc, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_bf16(d_tmem, make_desc_sw128(lo + ks * 32), make_desc_sw128(hi + ks * 32), idesc, 1);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_bf16(d_tmem, make_desc_sw128(hi + ks * 32), make_desc_sw128(hi + ks * 32), idesc, 1);
        tcgen05_commit(smem_u32(operand_empty + g));
        tcgen05_commit(smem_u32(tmem_full + g));
      }
    }
  } else {
    // ===================== epilogue =====================
    const int w = warp - kLoadWarps - 1;  // epilogue warp index 0..7
    const int eg = w >> 2;                // handles the quads with k % 2 == eg (TMEM buffer eg)
    const int q = warp & 3;               // TMEM lane quarter = warp % 4: rows of sample q of the quad
    float* os = reinterpret_cast<float*>(out_stage + (size_t)w * p.stage_bytes);
    const int OW = p.P + F * (F - 1) / 2;
    const int row_off = p.P + lane * (2 * F - lane - 1) / 2 - lane - 1;  // os[row_off + j] = dot(i = lane, j), j > i
    for (long long k = eg; k < nq; k += 2) {
      const int g = eg;
      const uint32_t ph = (uint32_t)((k >> 1) & 1);
      const long long s = ((long long)blockIdx.x + k * (long long)gridDim.x) * 4 + q;
      // bottom-MLP prefix: issue the global loads before waiting for the accumulator
      float pre0 = 0.0f, pre1 = 0.0f;
      if (p.P > 0 && s < p.B) {
        const float* pr = p.prefix + s * p.prefix_stride;
        pre0 = pr[lane];
        pre1 = pr[lane + 32];
      }
      mbar_wait(smem_u32(tmem_full + g), ph);
      tcgen05_fence_after();
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 128 + q * 32), r);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(tmem_empty + g));
      if (s < p.B) {  // warp-uniform
        if (p.P > 0) {
          os[lane] = pre0;
          os[lane + 32] = pre1;
        }
        if (lane < F) {
#pragma unroll
          for (int j = 1; j < 32; ++j)
            if (j > lane && j < F) os[row_off + j] = __uint_as_float(r[j]);
        }
        __syncwarp();
        if (p.out_split) {
          const float4* src = reinterpret_cast<const float4*>(os);
          __nv_bfloat16* drow = p.out_split + s * (2ll * p.out_Kp);
          const int groups = p.out_Kp >> 3;
          for (int gidx = lane; gidx < groups; gidx += 32) {
            const float4 a = src[2 * gidx], b = src[2 * gidx + 1];
            uint32_t h[4], l[4];
            split_pair(a.x, a.y, h[0], l[0]);
            split_pair(a.z, a.w, h[1], l[1]);
            split_pair(b.x, b.y, h[2], l[2]);
            split_pair(b.z, b.w, h[3], l[3]);
            *reinterpret_cast<uint4*>(drow + 8 * gidx) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(drow + p.out_Kp + 8 * gidx) = make_uint4(l[0], l[1], l[2], l[3]);
          }
        } else {
          float* dst = p.out_f32 + s * p.out_stride;
          if (((p.out_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out_f32) & 15) == 0)) {
            const int n4 = OW >> 2;
            for (int e = lane; e < n4; e += 32) reinterpret_cast<float4*>(dst)[e] = reinterpret_cast<const float4*>(os)[e];
            for (int e = (n4 << 2) + lane; e < OW; e += 32) dst[e] = os[e];
          } else {
            for (int e = lane; e < OW; e += 32) dst[e] = os[e];
          }
        }
        __syncwarp();
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == kLoadWarps) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// Returns MM_ERR_UNSUPPORTED when this path does not apply (the caller then tries the mma.sync kernel).
template <int MODE, typename IdxT>
int launch(const float* x, int64_t x_stride, const GatherParams& gp, const float* prefix, int64_t prefix_stride, int P,
           int bottom_slot, int64_t B, int F, int Dim, float* out_f32, int64_t out_stride, void* out_split, int out_Kp,
           int32_t* oob, cudaStream_t st, const char* who) {
  static int enabled = -1;
  if (enabled < 0) {
    // off by default: correct, but at 0.148 ms it does not yet beat the mma.sync kernel (0.124 ms) —
    // the convert -> MMA -> release chain on two operand buffers is latency-bound (profiles/r01_notes.md)
    const char* e = getenv("MM_INTERACT_TC");
    enabled = (e && e[0] == '1') ? 1 : 0;
  }
  if (!enabled) return MM_ERR_UNSUPPORTED;
  if (Dim != D || F < 2 || F > 32) return MM_ERR_UNSUPPORTED;
  if (P != 0 && P != D) return MM_ERR_UNSUPPORTED;
  if (out_f32 && out_split) return MM_ERR_UNSUPPORTED;
  if (MODE == 1 && gp.n_tables + (bottom_slot >= 0 ? 1 : 0) != F) return MM_ERR_UNSUPPORTED;
  if (MODE == 1 && (P > 0) != (bottom_slot >= 0)) return MM_ERR_UNSUPPORTED;
  if (MODE == 0 && (((uintptr_t)x & 15) || (x_stride & 3))) return MM_ERR_UNSUPPORTED;
  if (P > 0 && (((uintptr_t)prefix & 15) || (prefix_stride & 3))) return MM_ERR_UNSUPPORTED;
  const int OW = P + F * (F - 1) / 2;
  Params p;
  memset(&p, 0, sizeof(p));
  p.x = x;
  p.x_stride = x_stride;
  p.prefix = prefix;
  p.prefix_stride = prefix_stride;
  p.P = P;
  p.bottom_slot = bottom_slot;
  p.B = B;
  p.F = F;
  p.T = MODE == 1 ? gp.n_tables : 0;
  p.rows = F;  // MODE 0: the prefix is read from global memory by the epilogue, not staged
  p.out_f32 = out_f32;
  p.out_stride = out_stride;
  p.out_split = (__nv_bfloat16*)out_split;
  p.out_Kp = out_Kp;
  p.oob_count = oob;
  p.in_bytes = (unsigned)(F * D * 4);
  p.stage_bytes = out_split ? (unsigned)(out_Kp * 4) : (unsigned)(((OW + 3) & ~3) * 4);
  p.stage_bytes = (p.stage_bytes + 15u) & ~15u;
  const size_t smem = 1024 + 2 * (size_t)OPERAND_BYTES + (size_t)kLoadWarps * 2 * p.in_bytes + (size_t)kEpiWarps * p.stage_bytes + 128;
  if (smem > 227 * 1024) return MM_ERR_UNSUPPORTED;
  auto kern = interact_tc_kernel<MODE, IdxT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute failed: %s", who, cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  const long long quads = (B + 3) / 4;
  const long long sms = sm_count();
  const unsigned grid = (unsigned)(quads < sms ? quads : sms);
  kern<<<grid, kThreads, smem, st>>>(gp, p);
  return check_launch(who);
}

template int launch<0, int32_t>(const float*, int64_t, const GatherParams&, const float*, int64_t, int, int, int64_t, int, int,
                                float*, int64_t, void*, int, int32_t*, cudaStream_t, const char*);
template int launch<1, int32_t>(const float*, int64_t, const GatherParams&, const float*, int64_t, int, int, int64_t, int, int,
                                float*, int64_t, void*, int, int32_t*, cudaStream_t, const char*);
template int launch<1, int64_t>(const float*, int64_t, const GatherParams&, const float*, int64_t, int, int, int64_t, int, int,
                                float*, int64_t, void*, int, int32_t*, cudaStream_t, const char*);

}  // namespace itc
}  // namespace mm
