// Diagnostic kernels (NOT part of the product ABI): how fast can one GPU pull / push 256-byte rows from / to a
// peer's HBM over NVLink, by access pattern?  Built by tools/probe/peer_probe.py into tools/probe/_peer_probe.so.
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode 0: sequential 16-byte loads (every thread streams), sum into a sink
__global__ void seq_read(const float4* __restrict__ src, long long n4, float* sink) {
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(src + i));
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 1234.5f) *sink = acc;
}
// mode 3: sequential 16-byte stores
__global__ void seq_write(float4* __restrict__ dst, long long n4) {
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) dst[i] = v;
}

// mode 1: random 256-B rows by LDGSTS: `lpr` lanes per row (8 -> two 128-B halves per lane pair of instructions,
// 16 -> one instruction covers the whole row), `depth` row-groups in flight per warp
template <int LPR>
__global__ void rows_ldgsts(const uint8_t* __restrict__ src, const int* __restrict__ idx, long long n_rows, int depth,
                            int rows_per_group) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  constexpr int RPI = 32 / LPR;        // rows per instruction
  constexpr int J = 16 / LPR;          // 16-B copies per lane and row
  const uint32_t wbase = smem_u32(smem) + warp * depth * rows_per_group * 256;
  const long long gw = (long long)blockIdx.x * nw + warp, stride = (long long)gridDim.x * nw;
  const long long groups = n_rows / rows_per_group;
  long long issued = 0;
  int pending = 0;
  for (long long gi = gw; gi < groups; gi += stride, ++issued) {
    const int buf = (int)(issued % depth);
    for (int r0 = 0; r0 < rows_per_group; r0 += RPI) {
      const int r = r0 + lane / LPR;
      if (r < rows_per_group) {
        const long long row = idx[gi * rows_per_group + r];
        const uint8_t* s = src + row * 256 + (lane % LPR) * 16;
        const uint32_t d = wbase + (buf * rows_per_group + r) * 256 + (lane % LPR) * 16;
#pragma unroll
        for (int j = 0; j < J; ++j)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + j * LPR * 16), "l"(s + j * LPR * 16) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (++pending == depth) {
      asm volatile("cp.async.wait_group %0;" ::"n"(0) : "memory");  // simple: drain (depth groups were in flight)
      pending = 0;
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// modes 7/8: half of the rows local, half on the peer (by the low bit of the row id), like a 2-GPU row-sharded
// lookup.  SEPARATE = false: a warp instruction mixes local and remote rows; true: the local rows of a group are
// issued first, the remote ones in a second pass (no instruction touches both).
template <bool SEPARATE>
__global__ void rows_mixed(const uint8_t* __restrict__ local, const uint8_t* __restrict__ peer, const int* __restrict__ idx,
                           long long n_rows, int depth, int rows_per_group) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const uint32_t wbase = smem_u32(smem) + warp * depth * rows_per_group * 256;
  const long long gw = (long long)blockIdx.x * nw + warp, stride = (long long)gridDim.x * nw;
  const long long groups = n_rows / rows_per_group;
  long long issued = 0;
  int pending = 0;
  for (long long gi = gw; gi < groups; gi += stride, ++issued) {
    const int buf = (int)(issued % depth);
    for (int pass = 0; pass < (SEPARATE ? 2 : 1); ++pass) {
      for (int r0 = 0; r0 < rows_per_group; r0 += 4) {
        const int r = r0 + lane / 8;
        if (r < rows_per_group) {
          const long long row = idx[gi * rows_per_group + r];
          const bool remote = row & 1;
          const uint8_t* s = (remote ? peer : local) + (row >> 1) * 256 + (lane % 8) * 16;
          const uint32_t d = wbase + (buf * rows_per_group + r) * 256 + (lane % 8) * 16;
          if (!SEPARATE || (int)remote == pass) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(s) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 128), "l"(s + 128) : "memory");
          }
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (++pending == depth) {
      asm volatile("cp.async.wait_group %0;" ::"n"(0) : "memory");
      pending = 0;
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// mode 2: random 256-B rows by cp.async.bulk (TMA engine), one elected lane issues the rows of a group
__global__ void rows_bulk(const uint8_t* __restrict__ src, const int* __restrict__ idx, long long n_rows, int depth,
                          int rows_per_group) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);  // nw * depth mbarriers
  uint8_t* data = smem + 1024;
  const uint32_t wbase = smem_u32(data) + warp * depth * rows_per_group * 256;
  if (lane == 0)
    for (int d = 0; d < depth; ++d) {
      const uint32_t b = smem_u32(&bars[warp * depth + d]);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const long long gw = (long long)blockIdx.x * nw + warp, stride = (long long)gridDim.x * nw;
  const long long groups = n_rows / rows_per_group;
  long long issued = 0;
  for (long long gi = gw; gi < groups; gi += stride, ++issued) {
    const int buf = (int)(issued % depth);
    const uint32_t bar = smem_u32(&bars[warp * depth + buf]);
    if (issued >= depth) {  // wait for the previous use of this buffer
      const uint32_t parity = (uint32_t)(((issued / depth) - 1) & 1);
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
    const int my_row = lane < rows_per_group ? idx[gi * rows_per_group + lane] : 0;
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(rows_per_group * 256) : "memory");
    __syncwarp();
    if (lane < rows_per_group) {
      const uint8_t* s = src + (long long)my_row * 256;
      const uint32_t d = wbase + (buf * rows_per_group + lane) * 256;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 256, [%2];" ::"r"(d), "l"(s), "r"(bar) : "memory");
    }
  }
  // drain
  for (int d = 0; d < depth && d < issued; ++d) {
    const long long use = issued - 1 - d;
    const int buf = (int)(use % depth);
    const uint32_t bar = smem_u32(&bars[warp * depth + buf]);
    const uint32_t parity = (uint32_t)((use / depth) & 1);
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}

// mode 4: random 256-B rows by plain 16-byte loads, 16 lanes per row, `depth` rows per lane in registers
__global__ void rows_ldg(const uint8_t* __restrict__ src, const int* __restrict__ idx, long long n_rows, float* sink) {
  const int lane = threadIdx.x & 31;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, stride = ((long long)gridDim.x * blockDim.x) >> 5;
  float acc = 0.f;
  for (long long r0 = gw * 16; r0 + 16 <= n_rows; r0 += stride * 16) {
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long row = idx[r0 + 2 * k + (lane >> 4)];
      const float4* p = reinterpret_cast<const float4*>(src + row * 256) + (lane & 15);
      asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[k].x), "=f"(v[k].y), "=f"(v[k].z), "=f"(v[k].w) : "l"(p));
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += v[k].x + v[k].w;
  }
  if (acc == 1234.5f) *sink = acc;
}

// mode 5: random 256-B row PUSH: 16 lanes per row read a local row and store it to the peer
__global__ void rows_push(const uint8_t* __restrict__ local, uint8_t* __restrict__ peer, const int* __restrict__ idx, long long n_rows) {
  const int lane = threadIdx.x & 31;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, stride = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r0 = gw * 16; r0 + 16 <= n_rows; r0 += stride * 16) {
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long row = idx[r0 + 2 * k + (lane >> 4)];
      v[k] = reinterpret_cast<const float4*>(local + row * 256)[lane & 15];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) reinterpret_cast<float4*>(peer + (r0 + 2 * k + (lane >> 4)) * 256)[lane & 15] = v[k];
  }
}

extern "C" int probe_run(int mode, const void* src, void* dst, const int* idx, long long n_rows, int depth, int rows_per_group,
                         int blocks, int threads, float* sink, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)(threads / 32) * depth * rows_per_group * 256 + 1024;
  switch (mode) {
    case 0: seq_read<<<blocks, threads, 0, st>>>((const float4*)src, n_rows * 16, sink); break;
    case 3: seq_write<<<blocks, threads, 0, st>>>((float4*)dst, n_rows * 16); break;
    case 1:
      cudaFuncSetAttribute(rows_ldgsts<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      rows_ldgsts<8><<<blocks, threads, smem, st>>>((const uint8_t*)src, idx, n_rows, depth, rows_per_group);
      break;
    case 6:
      cudaFuncSetAttribute(rows_ldgsts<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      rows_ldgsts<16><<<blocks, threads, smem, st>>>((const uint8_t*)src, idx, n_rows, depth, rows_per_group);
      break;
    case 2:
      cudaFuncSetAttribute(rows_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      rows_bulk<<<blocks, threads, smem, st>>>((const uint8_t*)src, idx, n_rows, depth, rows_per_group);
      break;
    case 7:
      cudaFuncSetAttribute(rows_mixed<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      rows_mixed<false><<<blocks, threads, smem, st>>>((const uint8_t*)src, (const uint8_t*)dst, idx, n_rows, depth, rows_per_group);
      break;
    case 8:
      cudaFuncSetAttribute(rows_mixed<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      rows_mixed<true><<<blocks, threads, smem, st>>>((const uint8_t*)src, (const uint8_t*)dst, idx, n_rows, depth, rows_per_group);
      break;
    case 4: rows_ldg<<<blocks, threads, 0, st>>>((const uint8_t*)src, idx, n_rows, sink); break;
    case 5: rows_push<<<blocks, threads, 0, st>>>((const uint8_t*)src, (uint8_t*)dst, idx, n_rows); break;
    default: return -1;
  }
  return (int)cudaGetLastError();
}
