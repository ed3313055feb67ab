// Embedding lookups: fused multi-table one-hot gather, ragged bag lookup, dense-sequence
// lookup.  HBM-bound byte movers: coalesced index reads, 128-bit row copies, many independent
// row loads in flight per thread (random 256-B rows need >= ~35 KB in flight per SM).
// Reference call sites replaced: merlin/models/tf/inputs/embedding.py:424-471, :1126-1156,
// core/aggregation.py:54-66,101-108 (the stack/concat is folded into the output addressing).
#include <cstring>

#include "mm_common.cuh"

namespace mm {

// One warp = one (table, 32-sample chunk) task.  Lane l loads idx[b0+l] (one coalesced 128-B
// read), the indices are then broadcast with shuffles and the warp copies 32/VPR rows per step,
// UNROLL steps batched so every thread has UNROLL independent 16-B loads in flight.
template <typename IdxT, int VPR, int UNROLL>
__global__ void __launch_bounds__(256)
gather_rows_vec_kernel(const __grid_constant__ GatherParams p, long long B,
                       float* __restrict__ out, long long out_stride,
                       int* __restrict__ oob_count) {
  constexpr int RPS = 32 / VPR;  // rows per warp step
  const int lane = threadIdx.x & 31;
  const int sub = lane / VPR;
  const int v = lane % VPR;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long chunks = (B + 31) >> 5;
  const long long tasks = chunks * p.n_tables;
  float4* __restrict__ out4 = reinterpret_cast<float4*>(out);
  const long long out_stride4 = out_stride >> 2;

  for (long long task = warp0; task < tasks; task += n_warps) {
    const int t = (int)(task % p.n_tables);
    const long long b0 = (task / p.n_tables) << 5;
    const float4* __restrict__ w4 = reinterpret_cast<const float4*>(p.t[t].weights);
    const long long rows = p.t[t].rows;
    const int col4 = p.t[t].out_col >> 2;

    long long my_idx = -1;
    if (b0 + lane < B) {
      my_idx = (long long)reinterpret_cast<const IdxT*>(p.t[t].indices)[b0 + lane];
      if (my_idx < 0 || my_idx >= rows) {
        my_idx = -1;
        if (oob_count) atomicAdd(oob_count, 1);
      }
    }
#pragma unroll 1
    for (int s = 0; s < 32; s += RPS * UNROLL) {
      float4 val[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int r = s + u * RPS + sub;
        const long long ridx = __shfl_sync(0xffffffffu, my_idx, r);
        val[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ridx >= 0) val[u] = ldg_stream(w4 + ridx * VPR + v);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int r = s + u * RPS + sub;
        if (b0 + r < B) stg_stream(out4 + (b0 + r) * out_stride4 + col4 + v, val[u]);
      }
    }
  }
}

// Generic fallback: any dim / alignment.  One warp per (table, sample) row, scalar accesses.
template <typename IdxT>
__global__ void gather_rows_generic_kernel(const __grid_constant__ GatherParams p, long long B,
                                           float* __restrict__ out, long long out_stride,
                                           int* __restrict__ oob_count) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long tasks = B * p.n_tables;
  for (long long task = warp0; task < tasks; task += n_warps) {
    const int t = (int)(task % p.n_tables);
    const long long b = task / p.n_tables;
    const int dim = p.t[t].dim;
    long long idx = (long long)reinterpret_cast<const IdxT*>(p.t[t].indices)[b];
    const bool ok = idx >= 0 && idx < p.t[t].rows;
    if (!ok && lane == 0 && oob_count) atomicAdd(oob_count, 1);
    const float* __restrict__ src = p.t[t].weights + idx * dim;
    float* __restrict__ dst = out + b * out_stride + p.t[t].out_col;
    for (int d = lane; d < dim; d += 32) dst[d] = ok ? __ldg(src + d) : 0.0f;
  }
}

// Ragged bag lookup: one warp per bag; lanes span the embedding dim, ids are visited left to
// right (sequential fp32 adds = the oracle's order), 4 row loads batched ahead of the adds.
template <typename IdxT, typename OffT>
__global__ void gather_bag_kernel(const float* __restrict__ w, long long rows, int dim,
                                  const IdxT* __restrict__ values, const OffT* __restrict__ offsets,
                                  long long B, int combiner, float* __restrict__ out,
                                  long long out_stride, int out_col, int* __restrict__ oob_count) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long b = warp0; b < B; b += n_warps) {
    const long long beg = (long long)offsets[b], end = (long long)offsets[b + 1];
    for (int d0 = 0; d0 < dim; d0 += 32) {
      const int d = d0 + lane;
      float acc = 0.0f;
      int cnt = 0;
      for (long long i = beg; i < end; i += 4) {
        float x[4];
        bool use[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          use[u] = false;
          x[u] = 0.0f;
          if (i + u < end) {
            const long long id = (long long)values[i + u];
            if (id >= 0) {  // safe_embedding_lookup_sparse prunes ids < 0
              if (id < rows) {
                use[u] = true;
                if (d < dim) x[u] = __ldg(w + id * dim + d);
              } else if (lane == 0 && d0 == 0 && oob_count) {
                atomicAdd(oob_count, 1);
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (use[u]) {
            acc = __fadd_rn(acc, x[u]);
            ++cnt;
          }
      }
      if (cnt > 0) {
        if (combiner == MM_COMBINER_MEAN) acc = __fdiv_rn(acc, (float)cnt);
        else if (combiner == MM_COMBINER_SQRTN) acc = __fdiv_rn(acc, sqrtf((float)cnt));
      }
      if (d < dim) out[b * out_stride + out_col + d] = acc;  // empty bag -> zeros
    }
  }
}

// Dense (B, L) sequence lookup + mean/sum/max over L, padding not masked.
template <typename IdxT>
__global__ void gather_seq_kernel(const float* __restrict__ w, long long rows, int dim,
                                  const IdxT* __restrict__ ids, long long B, int L, int combiner,
                                  float* __restrict__ out, long long out_stride, int out_col,
                                  int* __restrict__ oob_count) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long b = warp0; b < B; b += n_warps) {
    for (int d0 = 0; d0 < dim; d0 += 32) {
      const int d = d0 + lane;
      float acc = (combiner == MM_COMBINER_MAX) ? -INFINITY : 0.0f;
      for (int l = 0; l < L; ++l) {
        const long long id = (long long)ids[b * L + l];
        float x = 0.0f;
        if (id >= 0 && id < rows) {
          if (d < dim) x = __ldg(w + id * dim + d);
        } else if (lane == 0 && d0 == 0 && oob_count) {
          atomicAdd(oob_count, 1);
        }
        acc = (combiner == MM_COMBINER_MAX) ? fmaxf(acc, x) : __fadd_rn(acc, x);
      }
      if (combiner == MM_COMBINER_MEAN) acc = __fdiv_rn(acc, (float)L);
      if (d < dim) out[b * out_stride + out_col + d] = acc;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Row-sharded tables over `world` GPUs of one NVLink domain (row r lives on rank r % world at local
// row r / world).  Owner-computes push: every rank scans the GLOBAL index list (replicated by an
// all-gather, 104 B/sample), and for the rows it owns copies the 256-byte row from its local shard
// straight into the destination rank's (B_local, F, D) stack through peer-mapped memory — the
// gather and the all-to-all are one kernel, no pack/unpack buffers, no size exchange.
// Reference counterpart: sok.lookup_sparse on a distributed sok.Variable
// (merlin/models/tf/distributed/embedding.py:75-84,144-148).
// ---------------------------------------------------------------------------------------------
struct ShardDst {
  float* ptr[16];
};

template <typename IdxT, int VPR>
__global__ void __launch_bounds__(256)
shard_gather_push_kernel(const __grid_constant__ GatherParams p, const __grid_constant__ ShardDst dst, long long B_global,
                         long long B_local, int rank, int world, long long out_stride, int* __restrict__ oob_count) {
  constexpr int RPS = 32 / VPR;  // rows copied per warp step
  const int lane = threadIdx.x & 31;
  const int sub = lane / VPR, v = lane % VPR;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long chunks = (B_global + 31) >> 5;
  const long long tasks = chunks * p.n_tables;
  for (long long task = warp0; task < tasks; task += n_warps) {
    const int t = (int)(task % p.n_tables);
    const long long g0 = (task / p.n_tables) << 5;
    const float4* __restrict__ w4 = reinterpret_cast<const float4*>(p.t[t].weights);
    const long long rows = p.t[t].rows;  // GLOBAL row count of the table
    const int col4 = p.t[t].out_col >> 2;
    long long local = -2;  // -2: not mine, -1: mine but out of range (write zeros), >= 0: local row
    if (g0 + lane < B_global) {
      const long long idx = (long long)reinterpret_cast<const IdxT*>(p.t[t].indices)[g0 + lane];
      const long long owner = ((idx % world) + world) % world;
      if (owner == rank) {
        if (idx >= 0 && idx < rows) local = idx / world;
        else {
          local = -1;
          if (oob_count) atomicAdd(oob_count, 1);
        }
      }
    }
    unsigned mine = __ballot_sync(0xffffffffu, local != -2);
    while (mine) {  // RPS owned samples per step
      int src_lane = -1;
      unsigned m = mine;
      for (int k = 0; k <= sub && m; ++k) {
        src_lane = __ffs(m) - 1;
        m &= m - 1;
        if (k < sub) src_lane = -1;
      }
      // drop the RPS lowest set bits
      for (int k = 0; k < RPS && mine; ++k) mine &= mine - 1;
      const long long l = __shfl_sync(0xffffffffu, local, src_lane < 0 ? 0 : src_lane);
      if (src_lane >= 0) {
        const long long g = g0 + src_lane;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (l >= 0) val = ldg_stream(w4 + l * VPR + v);
        float4* o = reinterpret_cast<float4*>(dst.ptr[g / B_local]) + (g % B_local) * (out_stride >> 2) + col4 + v;
        *o = val;  // peer (NVLink) or local store
      }
    }
  }
}

// deterministic initialiser of a row shard: local row l holds global row row0 + l*row_step
__global__ void init_uniform_hash_rows_kernel(float* __restrict__ w, long long local_rows, int D, unsigned long long seed,
                                              float lo, float span, long long row0, long long row_step) {
  const long long n = local_rows * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long l = i / D, d = i % D;
    const unsigned long long e = (unsigned long long)((row0 + l * row_step) * D + d);
    unsigned long long z = seed + (e + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    const float u = (float)(unsigned int)(z >> 40) * (1.0f / 16777216.0f);
    w[i] = __fadd_rn(lo, __fmul_rn(span, u));
  }
}

static int validate_tables(const mm_gather_table* tb, int n, const char* who) {
  MM_REQUIRE(tb != nullptr && n > 0 && n <= MM_MAX_TABLES, MM_ERR_ARG,
             "%s: n_tables=%d outside 1..%d or null table list", who, n, MM_MAX_TABLES);
  for (int t = 0; t < n; ++t)
    MM_REQUIRE(tb[t].weights && tb[t].indices && tb[t].rows > 0 && tb[t].dim > 0 &&
                   tb[t].out_col >= 0,
               MM_ERR_ARG, "%s: table %d has a null pointer or non-positive rows/dim", who, t);
  return MM_OK;
}

template <typename IdxT>
static int launch_gather(const GatherParams& p, int64_t B, float* out, int64_t out_stride,
                         int32_t* oob, cudaStream_t st) {
  // vector path: all dims equal, dim/4 a power of two <= 32, everything 16-B aligned
  bool vec = (out_stride % 4 == 0) && ((uintptr_t)out % 16 == 0);
  const int dim0 = p.t[0].dim;
  for (int t = 0; t < p.n_tables && vec; ++t)
    vec = p.t[t].dim == dim0 && p.t[t].out_col % 4 == 0 && ((uintptr_t)p.t[t].weights % 16 == 0);
  const int vpr = dim0 / 4;
  vec = vec && dim0 % 4 == 0 && vpr >= 1 && vpr <= 32 && (vpr & (vpr - 1)) == 0;
  const int threads = 256;
  const long long chunks = (B + 31) / 32;
  if (vec) {
    long long warps = chunks * p.n_tables;
    long long blocks = (warps * 32 + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 8 * 4;  // a few waves of 8 CTAs/SM
    if (blocks > cap) blocks = cap;
#define MM_LAUNCH_VEC(V, U)                                                                     \
  gather_rows_vec_kernel<IdxT, V, U><<<(unsigned)blocks, threads, 0, st>>>(p, B, out, out_stride, \
                                                                           oob)
    switch (vpr) {
      case 1: MM_LAUNCH_VEC(1, 1); break;
      case 2: MM_LAUNCH_VEC(2, 2); break;
      case 4: MM_LAUNCH_VEC(4, 4); break;
      case 8: MM_LAUNCH_VEC(8, 8); break;
      case 16: MM_LAUNCH_VEC(16, 8); break;
      default: MM_LAUNCH_VEC(32, 8); break;
    }
#undef MM_LAUNCH_VEC
  } else {
    long long warps = B * p.n_tables;
    long long blocks = (warps * 32 + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 8 * 8;
    if (blocks > cap) blocks = cap;
    gather_rows_generic_kernel<IdxT><<<(unsigned)blocks, threads, 0, st>>>(p, B, out, out_stride, oob);
  }
  return check_launch("mm_gather_multi");
}

}  // namespace mm

extern "C" {

int mm_gather_multi(const mm_gather_table* tables_host, int n_tables, int idx_dtype, int64_t B,
                    float* out, int64_t out_stride, int32_t* oob_count, void* stream) {
  int rc = mm::validate_tables(tables_host, n_tables, "mm_gather_multi");
  if (rc) return rc;
  MM_REQUIRE(out != nullptr && B >= 0 && out_stride > 0, MM_ERR_ARG,
             "mm_gather_multi: null out, B<0 or out_stride<=0");
  MM_REQUIRE(idx_dtype == MM_I32 || idx_dtype == MM_I64, MM_ERR_ARG,
             "mm_gather_multi: idx_dtype must be MM_I32 or MM_I64");
  for (int t = 0; t < n_tables; ++t)
    MM_REQUIRE((int64_t)tables_host[t].out_col + tables_host[t].dim <= out_stride, MM_ERR_ARG,
               "mm_gather_multi: table %d (out_col %d + dim %d) exceeds out_stride %lld", t,
               tables_host[t].out_col, tables_host[t].dim, (long long)out_stride);
  if (B == 0) return MM_OK;
  mm::GatherParams p;
  memset(&p, 0, sizeof(p));
  p.n_tables = n_tables;
  for (int t = 0; t < n_tables; ++t) p.t[t] = tables_host[t];
  return idx_dtype == MM_I32
             ? mm::launch_gather<int32_t>(p, B, out, out_stride, oob_count, (cudaStream_t)stream)
             : mm::launch_gather<int64_t>(p, B, out, out_stride, oob_count, (cudaStream_t)stream);
}

int mm_gather_bag(const float* weights, int64_t rows, int dim, const void* values, int idx_dtype,
                  const void* offsets, int off_dtype, int64_t B, int combiner, float* out,
                  int64_t out_stride, int out_col, int32_t* oob_count, void* stream) {
  MM_REQUIRE(weights && offsets && out && rows > 0 && dim > 0 && B >= 0, MM_ERR_ARG,
             "mm_gather_bag: null pointer or non-positive rows/dim");
  MM_REQUIRE(combiner == MM_COMBINER_MEAN || combiner == MM_COMBINER_SUM ||
                 combiner == MM_COMBINER_SQRTN,
             MM_ERR_ARG, "mm_gather_bag: combiner must be mean, sum or sqrtn");
  MM_REQUIRE((idx_dtype == MM_I32 || idx_dtype == MM_I64) &&
                 (off_dtype == MM_I32 || off_dtype == MM_I64),
             MM_ERR_ARG, "mm_gather_bag: bad index / offset dtype");
  MM_REQUIRE(out_col >= 0 && (int64_t)out_col + dim <= out_stride, MM_ERR_ARG,
             "mm_gather_bag: out_col + dim exceeds out_stride");
  if (B == 0) return MM_OK;
  const int threads = 256;
  long long blocks = (B * 32 + threads - 1) / threads;
  const long long cap = (long long)mm::sm_count() * 8 * 4;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
#define MM_BAG(IT, OT)                                                                        \
  mm::gather_bag_kernel<IT, OT><<<(unsigned)blocks, threads, 0, st>>>(                        \
      weights, rows, dim, (const IT*)values, (const OT*)offsets, B, combiner, out, out_stride, \
      out_col, oob_count)
  if (idx_dtype == MM_I32 && off_dtype == MM_I32) MM_BAG(int32_t, int32_t);
  else if (idx_dtype == MM_I32) MM_BAG(int32_t, int64_t);
  else if (off_dtype == MM_I32) MM_BAG(int64_t, int32_t);
  else MM_BAG(int64_t, int64_t);
#undef MM_BAG
  return mm::check_launch("mm_gather_bag");
}

int mm_gather_seq(const float* weights, int64_t rows, int dim, const void* ids, int idx_dtype,
                  int64_t B, int L, int combiner, float* out, int64_t out_stride, int out_col,
                  int32_t* oob_count, void* stream) {
  MM_REQUIRE(weights && ids && out && rows > 0 && dim > 0 && B >= 0 && L > 0, MM_ERR_ARG,
             "mm_gather_seq: null pointer or non-positive rows/dim/L");
  MM_REQUIRE(combiner == MM_COMBINER_MEAN || combiner == MM_COMBINER_SUM ||
                 combiner == MM_COMBINER_MAX,
             MM_ERR_ARG, "mm_gather_seq: combiner must be mean, sum or max");
  MM_REQUIRE(idx_dtype == MM_I32 || idx_dtype == MM_I64, MM_ERR_ARG, "mm_gather_seq: bad dtype");
  MM_REQUIRE(out_col >= 0 && (int64_t)out_col + dim <= out_stride, MM_ERR_ARG,
             "mm_gather_seq: out_col + dim exceeds out_stride");
  if (B == 0) return MM_OK;
  const int threads = 256;
  long long blocks = (B * 32 + threads - 1) / threads;
  const long long cap = (long long)mm::sm_count() * 8 * 4;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  if (idx_dtype == MM_I32)
    mm::gather_seq_kernel<int32_t><<<(unsigned)blocks, threads, 0, st>>>(
        weights, rows, dim, (const int32_t*)ids, B, L, combiner, out, out_stride, out_col, oob_count);
  else
    mm::gather_seq_kernel<int64_t><<<(unsigned)blocks, threads, 0, st>>>(
        weights, rows, dim, (const int64_t*)ids, B, L, combiner, out, out_stride, out_col, oob_count);
  return mm::check_launch("mm_gather_seq");
}


int mm_shard_gather_push(const mm_gather_table* tables_host, int n_tables, int idx_dtype, int64_t B_global,
                         int64_t B_local, int D, int rank, int world, void* const* dst_ptrs_host,
                         int64_t out_stride, int32_t* oob_count, void* stream) {
  int rc = mm::validate_tables(tables_host, n_tables, "mm_shard_gather_push");
  if (rc) return rc;
  MM_REQUIRE(world >= 1 && world <= 16 && rank >= 0 && rank < world && dst_ptrs_host, MM_ERR_ARG,
             "mm_shard_gather_push: world must be 1..16, 0 <= rank < world, dst_ptrs non-null");
  MM_REQUIRE(B_local > 0 && B_global == B_local * world, MM_ERR_ARG, "mm_shard_gather_push: B_global must equal B_local * world");
  MM_REQUIRE(idx_dtype == MM_I32 || idx_dtype == MM_I64, MM_ERR_ARG, "mm_shard_gather_push: bad idx_dtype");
  const int vpr = D / 4;
  MM_REQUIRE(D % 4 == 0 && vpr >= 1 && vpr <= 32 && (vpr & (vpr - 1)) == 0 && out_stride % 4 == 0, MM_ERR_UNSUPPORTED,
             "mm_shard_gather_push: D/4 must be a power of two <= 32 and out_stride a multiple of 4");
  mm::GatherParams p;
  memset(&p, 0, sizeof(p));
  p.n_tables = n_tables;
  for (int t = 0; t < n_tables; ++t) {
    MM_REQUIRE(tables_host[t].dim == D && tables_host[t].out_col % 4 == 0 && (int64_t)tables_host[t].out_col + D <= out_stride &&
                   ((uintptr_t)tables_host[t].weights % 16) == 0,
               MM_ERR_ARG, "mm_shard_gather_push: table %d: dim != D, bad out_col or misaligned shard", t);
    p.t[t] = tables_host[t];
  }
  mm::ShardDst dst;
  memset(&dst, 0, sizeof(dst));
  for (int r = 0; r < world; ++r) {
    MM_REQUIRE(dst_ptrs_host[r] && ((uintptr_t)dst_ptrs_host[r] % 16) == 0, MM_ERR_ARG,
               "mm_shard_gather_push: destination pointer of rank %d is null or misaligned", r);
    dst.ptr[r] = (float*)dst_ptrs_host[r];
  }
  const int threads = 256;
  long long warps = ((B_global + 31) / 32) * n_tables;
  long long blocks = (warps * 32 + threads - 1) / threads;
  const long long cap = (long long)mm::sm_count() * 8 * 4;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
#define MM_PUSH(IT, V)                                                                                              \
  mm::shard_gather_push_kernel<IT, V><<<(unsigned)blocks, threads, 0, st>>>(p, dst, B_global, B_local, rank, world, \
                                                                           out_stride, oob_count)
#define MM_PUSH_V(IT)        \
  switch (vpr) {             \
    case 1: MM_PUSH(IT, 1); break;   \
    case 2: MM_PUSH(IT, 2); break;   \
    case 4: MM_PUSH(IT, 4); break;   \
    case 8: MM_PUSH(IT, 8); break;   \
    case 16: MM_PUSH(IT, 16); break; \
    default: MM_PUSH(IT, 32); break; \
  }
  if (idx_dtype == MM_I32) { MM_PUSH_V(int32_t) } else { MM_PUSH_V(int64_t) }
#undef MM_PUSH_V
#undef MM_PUSH
  return mm::check_launch("mm_shard_gather_push");
}

int mm_init_uniform_hash_rows(float* w, int64_t local_rows, int D, uint64_t seed, float lo, float hi,
                              int64_t row0, int64_t row_step, void* stream) {
  MM_REQUIRE(w && local_rows >= 0 && D > 0 && row0 >= 0 && row_step >= 1, MM_ERR_ARG,
             "mm_init_uniform_hash_rows: null buffer or bad rows / D / row0 / row_step");
  if (local_rows == 0) return MM_OK;
  const long long n = local_rows * D;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)mm::sm_count() * 32;
  if (blocks > cap) blocks = cap;
  mm::init_uniform_hash_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      w, local_rows, D, (unsigned long long)seed, lo, hi - lo, row0, row_step);
  return mm::check_launch("mm_init_uniform_hash_rows");
}

}  // extern "C"
