// HBM-bound elementwise / gather / reduction kernels of the Llama forward path.
// All are 16-byte vectorised, fp32 math, one CTA per token row (or grid-stride over vectors), and
// launch on the caller's stream with no allocation or sync.
//
// Reference semantics (S/ = python/sglang/srt/, K/ = sgl-kernel/):
//   RMSNorm            S/layers/layernorm.py:59-76     (CUDA: flashinfer norm.cuh via K/csrc/elementwise)
//   SiluAndMul         S/layers/activation.py:42-44    (CUDA: K/csrc/elementwise/activation.cu:36-50)
//   RotaryEmbedding    S/layers/rotary_embedding.py:32-58,113-141 (CUDA: K/csrc/elementwise/rope.cu:22-89)
//   set_kv_buffer      S/mem_cache/memory_pool.py:316-346
//   embedding / last-token gather  S/layers/vocab_parallel_embedding.py:469-490, logits_processor.py:244
//   greedy argmax      S/layers/sampler.py:73-77
#include "common.cuh"
#include "ptx.cuh"

namespace spd {

// 16-byte vector of eight 16-bit elements
struct __align__(16) Vec8 {
  uint32_t w[4];
};

template <typename T>
__device__ __forceinline__ void vec8_to_f32(const Vec8& v, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 p = unpack2<T>(v.w[i]);
    f[2 * i] = p.x;
    f[2 * i + 1] = p.y;
  }
}
template <typename T>
__device__ __forceinline__ Vec8 f32_to_vec8(const float* f) {
  Vec8 v;
#pragma unroll
  for (int i = 0; i < 4; ++i) v.w[i] = pack2<T>(f[2 * i], f[2 * i + 1]);
  return v;
}

__device__ __forceinline__ float block_sum(float v, float* smem) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  if (lane_id() == 0) smem[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < nwarp) ? smem[threadIdx.x] : 0.f;
  if (warp == 0) {
    t = warp_sum(t);
    if (threadIdx.x == 0) smem[32] = t;
  }
  __syncthreads();
  return smem[32];
}

// ------------------------------------------------------------------------------------------------
// RMSNorm and fused residual-add RMSNorm.  One CTA per row; the row stays in registers
// (kMaxVec 16-byte vectors per thread) so HBM traffic is exactly 1 read + 1 write per tensor.
// ------------------------------------------------------------------------------------------------
constexpr int kNormThreads = 256;
// kMaxVec is a template parameter (2, 4 or 8: rows up to 4096 / 8192 / 16384 elements stay in
// registers) so that the common 4096-wide row costs ~40 registers and 8 CTAs fit per SM — the
// kernel is a pure HBM stream and needs the bytes in flight.

template <typename T, bool kFusedAdd, int kMaxVec>
__global__ void __launch_bounds__(kNormThreads)
rmsnorm_kernel(T* out, T* x, T* residual,  // out may alias x (fused variant is in place)
               const T* __restrict__ w, int64_t dim, int64_t x_stride, int64_t out_stride,
               float eps) {
  __shared__ float red[33];
  // PDL: release the dependents first (they block in their own griddepcontrol.wait until THIS grid
  // has completed, and a grid completes only after its own wait — so order is transitive), then
  // wait for the predecessor.  A decode GEMM two kernels downstream can then take over an SM the
  // moment the previous GEMM's CTA leaves it and stream its weights under this kernel.
  griddep_launch_dependents();
  griddep_wait();
  const int64_t row = blockIdx.x;
  const int nvec = static_cast<int>(dim >> 3);
  const Vec8* xr = reinterpret_cast<const Vec8*>(x + row * x_stride);
  Vec8* rr = kFusedAdd ? reinterpret_cast<Vec8*>(residual + row * x_stride) : nullptr;
  float vals[kMaxVec][8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int v = threadIdx.x + i * kNormThreads;
    if (v < nvec) {
      vec8_to_f32<T>(xr[v], vals[i]);
      if constexpr (kFusedAdd) {
        float r[8];
        vec8_to_f32<T>(rr[v], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) vals[i][j] += r[j];
        rr[v] = f32_to_vec8<T>(vals[i]);  // residual_out = (x + residual).to(dtype)
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += vals[i][j] * vals[i][j];
    }
  }
  // rows longer than the register budget: stream the remainder (re-read in the second pass)
  for (int v = threadIdx.x + kMaxVec * kNormThreads; v < nvec; v += kNormThreads) {
    float f[8];
    vec8_to_f32<T>(xr[v], f);
    if constexpr (kFusedAdd) {
      float r[8];
      vec8_to_f32<T>(rr[v], r);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += r[j];
      rr[v] = f32_to_vec8<T>(f);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
  }
  const float total = block_sum(ss, red);
  const float inv = rsqrtf(total / static_cast<float>(dim) + eps);
  Vec8* orow = reinterpret_cast<Vec8*>(out + row * out_stride);
  const Vec8* wv = reinterpret_cast<const Vec8*>(w);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int v = threadIdx.x + i * kNormThreads;
    if (v < nvec) {
      float wf[8], o[8];
      vec8_to_f32<T>(wv[v], wf);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = vals[i][j] * inv * wf[j];
      orow[v] = f32_to_vec8<T>(o);
    }
  }
  for (int v = threadIdx.x + kMaxVec * kNormThreads; v < nvec; v += kNormThreads) {
    float f[8], wf[8];
    // fused: the summed row was stored (rounded) to residual; unfused: re-read x.  The rounded
    // value differs from the fp32 sum by <= 1/2 ulp, inside the op's 1e-3 tolerance; rows this
    // long (> 8192) do not occur on the Llama path.
    vec8_to_f32<T>(kFusedAdd ? rr[v] : xr[v], f);
    vec8_to_f32<T>(wv[v], wf);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = f[j] * inv * wf[j];
    orow[v] = f32_to_vec8<T>(f);
  }
}

// ------------------------------------------------------------------------------------------------
// SiLU(x[:, :d]) * x[:, d:]
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
silu_and_mul_kernel(T* __restrict__ out, const T* __restrict__ x, int64_t tokens, int64_t d) {
  // PDL: release the dependents first (they block in their own griddepcontrol.wait until THIS grid
  // has completed, and a grid completes only after its own wait — so order is transitive), then
  // wait for the predecessor.  A decode GEMM two kernels downstream can then take over an SM the
  // moment the previous GEMM's CTA leaves it and stream its weights under this kernel.
  griddep_launch_dependents();
  griddep_wait();
  const int64_t nvec_row = d >> 3;
  const int64_t total = tokens * nvec_row;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t t = i / nvec_row, c = i - t * nvec_row;
    const Vec8* row = reinterpret_cast<const Vec8*>(x + t * 2 * d);
    float g[8], u[8], o[8];
    vec8_to_f32<T>(row[c], g);
    vec8_to_f32<T>(row[nvec_row + c], u);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = silu_f32(g[j]) * u[j];
    reinterpret_cast<Vec8*>(out + t * d)[c] = f32_to_vec8<T>(o);
  }
}

// ------------------------------------------------------------------------------------------------
// RoPE (+ optional KV-pool scatter).  One CTA per token; work items are 16-byte vectors:
//   rope items   : (head in q||k) x (rot_dim/16 vector pairs for NeoX | rot_dim/8 vectors for GPT-J)
//   pass items   : k dims >= rot_dim (copied to the pool unchanged)
//   v items      : Hk*D/8 vectors of v copied to the pool
// ------------------------------------------------------------------------------------------------
template <typename T, bool kWritePool>
__global__ void __launch_bounds__(256)
rope_kernel(T* __restrict__ q, T* __restrict__ k, const T* __restrict__ v, T* __restrict__ k_pool,
            T* __restrict__ v_pool, const float* __restrict__ cos_sin, const int64_t* __restrict__ pos_ids,
            const int64_t* __restrict__ loc, int Hq, int Hk, int D, int rot, int64_t q_stride,
            int64_t k_stride, int64_t v_stride, int interleave) {
  // PDL: release the dependents first (they block in their own griddepcontrol.wait until THIS grid
  // has completed, and a grid completes only after its own wait — so order is transitive), then
  // wait for the predecessor.  A decode GEMM two kernels downstream can then take over an SM the
  // moment the previous GEMM's CTA leaves it and stream its weights under this kernel.
  griddep_launch_dependents();
  griddep_wait();
  const int64_t t = blockIdx.x;
  const int64_t pos = pos_ids[t];
  const float* cs = cos_sin + pos * rot;  // [cos(rot/2) || sin(rot/2)]
  const int half = rot >> 1;
  T* qrow = q + t * q_stride;
  T* krow = k + t * k_stride;
  T* kdst = nullptr;
  T* vdst = nullptr;
  if constexpr (kWritePool) {
    const int64_t slot = loc[t];
    kdst = k_pool + slot * static_cast<int64_t>(Hk) * D;
    vdst = v_pool + slot * static_cast<int64_t>(Hk) * D;
  }
  const int H = Hq + Hk;
  if (!interleave) {
    const int per_head = half >> 3;  // vector pairs per head
    for (int it = threadIdx.x; it < H * per_head; it += blockDim.x) {
      const int h = it / per_head, c = it - h * per_head;
      const bool is_k = h >= Hq;
      T* base = is_k ? (krow + (h - Hq) * D) : (qrow + h * D);
      Vec8 a = reinterpret_cast<Vec8*>(base)[c];
      Vec8 b = reinterpret_cast<Vec8*>(base + half)[c];
      float x1[8], x2[8], o1[8], o2[8];
      vec8_to_f32<T>(a, x1);
      vec8_to_f32<T>(b, x2);
      const float4* cp = reinterpret_cast<const float4*>(cs + c * 8);
      const float4* sp = reinterpret_cast<const float4*>(cs + half + c * 8);
      float cv[8], sv[8];
      *reinterpret_cast<float4*>(cv) = cp[0];
      *reinterpret_cast<float4*>(cv + 4) = cp[1];
      *reinterpret_cast<float4*>(sv) = sp[0];
      *reinterpret_cast<float4*>(sv + 4) = sp[1];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o1[j] = x1[j] * cv[j] - x2[j] * sv[j];
        o2[j] = x2[j] * cv[j] + x1[j] * sv[j];
      }
      Vec8 r1 = f32_to_vec8<T>(o1), r2 = f32_to_vec8<T>(o2);
      reinterpret_cast<Vec8*>(base)[c] = r1;
      reinterpret_cast<Vec8*>(base + half)[c] = r2;
      if constexpr (kWritePool) {
        if (is_k) {
          T* d = kdst + (h - Hq) * D;
          reinterpret_cast<Vec8*>(d)[c] = r1;
          reinterpret_cast<Vec8*>(d + half)[c] = r2;
        }
      }
    }
  } else {
    const int per_head = rot >> 3;  // vectors per head; each holds 4 (even, odd) pairs
    for (int it = threadIdx.x; it < H * per_head; it += blockDim.x) {
      const int h = it / per_head, c = it - h * per_head;
      const bool is_k = h >= Hq;
      T* base = is_k ? (krow + (h - Hq) * D) : (qrow + h * D);
      Vec8 a = reinterpret_cast<Vec8*>(base)[c];
      float x[8], o[8];
      vec8_to_f32<T>(a, x);
      const float4 cv = *reinterpret_cast<const float4*>(cs + c * 4);
      const float4 sv = *reinterpret_cast<const float4*>(cs + half + c * 4);
      const float cc[4] = {cv.x, cv.y, cv.z, cv.w}, ssn[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o[2 * j] = x[2 * j] * cc[j] - x[2 * j + 1] * ssn[j];
        o[2 * j + 1] = x[2 * j + 1] * cc[j] + x[2 * j] * ssn[j];
      }
      Vec8 r = f32_to_vec8<T>(o);
      reinterpret_cast<Vec8*>(base)[c] = r;
      if constexpr (kWritePool) {
        if (is_k) reinterpret_cast<Vec8*>(kdst + (h - Hq) * D)[c] = r;
      }
    }
  }
  if constexpr (kWritePool) {
    // k dims beyond rot_dim pass through unchanged
    const int pass = (D - rot) >> 3;
    for (int it = threadIdx.x; it < Hk * pass; it += blockDim.x) {
      const int h = it / pass, c = it - h * pass;
      reinterpret_cast<Vec8*>(kdst + h * D + rot)[c] =
          reinterpret_cast<const Vec8*>(krow + h * D + rot)[c];
    }
    const int nv = (Hk * D) >> 3;
    const Vec8* vsrc = reinterpret_cast<const Vec8*>(v + t * v_stride);
    for (int it = threadIdx.x; it < nv; it += blockDim.x) reinterpret_cast<Vec8*>(vdst)[it] = vsrc[it];
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
kv_write_kernel(const T* __restrict__ k, const T* __restrict__ v, T* __restrict__ k_pool,
                T* __restrict__ v_pool, const int64_t* __restrict__ loc, int64_t row_elems,
                int64_t k_stride, int64_t v_stride) {
  // PDL: release the dependents first (they block in their own griddepcontrol.wait until THIS grid
  // has completed, and a grid completes only after its own wait — so order is transitive), then
  // wait for the predecessor.  A decode GEMM two kernels downstream can then take over an SM the
  // moment the previous GEMM's CTA leaves it and stream its weights under this kernel.
  griddep_launch_dependents();
  griddep_wait();
  const int64_t t = blockIdx.x;
  const int64_t slot = loc[t];
  const int nv = static_cast<int>(row_elems >> 3);
  const Vec8* ks = reinterpret_cast<const Vec8*>(k + t * k_stride);
  const Vec8* vs = reinterpret_cast<const Vec8*>(v + t * v_stride);
  Vec8* kd = reinterpret_cast<Vec8*>(k_pool + slot * row_elems);
  Vec8* vd = reinterpret_cast<Vec8*>(v_pool + slot * row_elems);
  if (v == nullptr) {  // single-buffer pools (MLA latent rows): only the k side exists
    for (int i = threadIdx.x; i < nv; i += blockDim.x) kd[i] = ks[i];
  } else {
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
      kd[i] = ks[i];
      vd[i] = vs[i];
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
gather_rows_kernel(T* __restrict__ out, const T* __restrict__ table, const int64_t* __restrict__ idx,
                   int64_t row_elems, int64_t table_rows) {
  // PDL: release the dependents first (they block in their own griddepcontrol.wait until THIS grid
  // has completed, and a grid completes only after its own wait — so order is transitive), then
  // wait for the predecessor.  A decode GEMM two kernels downstream can then take over an SM the
  // moment the previous GEMM's CTA leaves it and stream its weights under this kernel.
  griddep_launch_dependents();
  griddep_wait();
  const int64_t i = blockIdx.x;
  const int64_t r = idx[i];
  const int nv = static_cast<int>(row_elems >> 3);
  Vec8* d = reinterpret_cast<Vec8*>(out + i * row_elems);
  if (r < 0 || r >= table_rows) {  // out-of-vocabulary id: zero row (masked embedding)
    Vec8 z{{0, 0, 0, 0}};
    for (int c = threadIdx.x; c < nv; c += blockDim.x) d[c] = z;
  } else {
    const Vec8* s = reinterpret_cast<const Vec8*>(table + r * row_elems);
    for (int c = threadIdx.x; c < nv; c += blockDim.x) d[c] = s[c];
  }
}

// argmax over fp32 rows; ties resolve to the smallest index (torch.argmax on CUDA/CPU returns the
// first maximal element for distinct-valued logits; exact ties are measure-zero on this path).
constexpr int kArgmaxThreads = 1024;
__global__ void __launch_bounds__(kArgmaxThreads)
argmax_kernel(int32_t* __restrict__ ids, const float* __restrict__ logits, int64_t vocab, int64_t ld,
              int4* __restrict__ pairs, int idx_offset) {
  __shared__ float sv[32];
  __shared__ int si[32];
  // PDL: release the dependents first (they block in their own griddepcontrol.wait until THIS grid
  // has completed, and a grid completes only after its own wait — so order is transitive), then
  // wait for the predecessor.  A decode GEMM two kernels downstream can then take over an SM the
  // moment the previous GEMM's CTA leaves it and stream its weights under this kernel.
  griddep_launch_dependents();
  griddep_wait();
  const float* row = logits + blockIdx.x * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  const int64_t nv4 = ((reinterpret_cast<uintptr_t>(row) & 15u) == 0) ? (vocab >> 2) : 0;
  for (int64_t i = threadIdx.x; i < nv4; i += kArgmaxThreads) {
    const float4 f = reinterpret_cast<const float4*>(row)[i];
    const float a[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int id = static_cast<int>(i * 4 + j);
      if (a[j] > best || (a[j] == best && id < bi)) { best = a[j]; bi = id; }
    }
  }
  for (int64_t i = nv4 * 4 + threadIdx.x; i < vocab; i += kArgmaxThreads) {
    const float a = row[i];
    const int id = static_cast<int>(i);
    if (a > best || (a == best && id < bi)) { best = a; bi = id; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane_id() == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    best = sv[threadIdx.x];
    bi = si[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (threadIdx.x == 0) {
      bi = (bi == 0x7fffffff) ? 0 : bi;
      if (ids) ids[blockIdx.x] = bi;
      // vocab-sharded lm_head: (max, global index) of this rank's shard, one 16-byte row
      if (pairs) pairs[blockIdx.x] = make_int4(__float_as_int(best), bi + idx_offset, 0, 0);
    }
  }
}

// ids[b] = index of the largest of `parts` (max, index) candidates per row; ties resolve to the
// smallest index, so the result equals torch.argmax over the concatenated shards.
__global__ void argmax_merge_kernel(int32_t* __restrict__ ids, const int4* __restrict__ pairs, int parts,
                                    int64_t batch) {
  // PDL: release the dependents first (they block in their own griddepcontrol.wait until THIS grid
  // has completed, and a grid completes only after its own wait — so order is transitive), then
  // wait for the predecessor.  A decode GEMM two kernels downstream can then take over an SM the
  // moment the previous GEMM's CTA leaves it and stream its weights under this kernel.
  griddep_launch_dependents();
  griddep_wait();
  const int64_t b = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (b >= batch) return;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int r = 0; r < parts; ++r) {
    const int4 c = pairs[r * batch + b];
    const float v = __int_as_float(c.x);
    if (v > best || (v == best && c.y < bi)) { best = v; bi = c.y; }
  }
  ids[b] = (bi == 0x7fffffff) ? 0 : bi;
}

// ------------------------------------------------------------------------------------------------
// batch metadata (integer; exact)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
compute_positions_kernel(int64_t* __restrict__ positions, int32_t* __restrict__ extend_start_loc,
                         int32_t* __restrict__ token_req, const int32_t* __restrict__ prefix_lens,
                         const int32_t* __restrict__ extend_lens) {
  __shared__ long long red[8];
  __shared__ long long start_s;
  const int r = blockIdx.x;
  // exclusive prefix sum of extend_lens[0..r) — block-parallel (the Triton reference loops serially)
  long long part = 0;
  for (int i = threadIdx.x; i < r; i += blockDim.x) part += extend_lens[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane_id() == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long s = 0;
    for (int i = 0; i < 8; ++i) s += red[i];
    start_s = s;
    extend_start_loc[r] = static_cast<int32_t>(s);
  }
  __syncthreads();
  const long long start = start_s;
  const int pre = prefix_lens ? prefix_lens[r] : 0;
  const int len = extend_lens[r];
  for (int j = threadIdx.x; j < len; j += blockDim.x) {
    positions[start + j] = static_cast<int64_t>(pre) + j;
    if (token_req) token_req[start + j] = r;
  }
}

__global__ void decode_positions_kernel(int64_t* __restrict__ positions,
                                        const int64_t* __restrict__ seq_lens, int64_t batch) {
  // PDL: release the dependents first (they block in their own griddepcontrol.wait until THIS grid
  // has completed, and a grid completes only after its own wait — so order is transitive), then
  // wait for the predecessor.  A decode GEMM two kernels downstream can then take over an SM the
  // moment the previous GEMM's CTA leaves it and stream its weights under this kernel.
  griddep_launch_dependents();
  griddep_wait();
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < batch) {
    const int64_t p = seq_lens[i] - 1;
    positions[i] = p < 0 ? 0 : p;
  }
}

__global__ void __launch_bounds__(256)
build_kv_indices_kernel(int32_t* __restrict__ kv_indices, const int32_t* __restrict__ req_to_token,
                        const int64_t* __restrict__ req_pool_indices, const int32_t* __restrict__ lens,
                        const int32_t* __restrict__ kv_indptr, const int32_t* __restrict__ kv_start,
                        int64_t stride) {
  const int b = blockIdx.x;
  const int64_t req = req_pool_indices[b];
  const int start = kv_start ? kv_start[b] : 0;
  const int len = lens[b];
  const int32_t* src = req_to_token + req * stride + start;
  int32_t* dst = kv_indices + kv_indptr[b];
  for (int j = threadIdx.x; j < len; j += blockDim.x) dst[j] = src[j];
}

__global__ void __launch_bounds__(256)
write_req_to_token_kernel(int32_t* __restrict__ req_to_token, const int64_t* __restrict__ req_pool_indices,
                          const int64_t* __restrict__ pre_lens, const int64_t* __restrict__ seq_lens,
                          const int64_t* __restrict__ extend_lens, const int64_t* __restrict__ out_cache_loc,
                          int64_t stride) {
  __shared__ long long red[8];
  __shared__ long long start_s;
  const int b = blockIdx.x;
  long long part = 0;
  for (int i = threadIdx.x; i < b; i += blockDim.x) part += extend_lens[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane_id() == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long s = 0;
    for (int i = 0; i < 8; ++i) s += red[i];
    start_s = s;
  }
  __syncthreads();
  const int64_t req = req_pool_indices[b];
  const int64_t pre = pre_lens[b], seq = seq_lens[b];
  int32_t* dst = req_to_token + req * stride + pre;
  const int64_t* src = out_cache_loc + start_s;
  for (int64_t j = threadIdx.x; j < seq - pre; j += blockDim.x) dst[j] = static_cast<int32_t>(src[j]);
}

// ------------------------------------------------------------------------------------------------
// OPT-family ops (BASELINE configs[0], the plumbing model): LayerNorm with bias, and
// y = act(x + bias) + residual.  fp32 math, 16-bit I/O, 16-byte vectors.
// ------------------------------------------------------------------------------------------------
template <typename T, int kMaxVec>
__global__ void __launch_bounds__(kNormThreads)
layernorm_kernel(T* __restrict__ out, const T* __restrict__ x, const T* __restrict__ gamma,
                 const T* __restrict__ beta, int64_t dim, int64_t x_stride, int64_t out_stride, float eps) {
  __shared__ float red[33];
  // PDL: release the dependents first (they block in their own griddepcontrol.wait until THIS grid
  // has completed, and a grid completes only after its own wait — so order is transitive), then
  // wait for the predecessor.  A decode GEMM two kernels downstream can then take over an SM the
  // moment the previous GEMM's CTA leaves it and stream its weights under this kernel.
  griddep_launch_dependents();
  griddep_wait();
  const int64_t row = blockIdx.x;
  const int nvec = static_cast<int>(dim >> 3);
  const Vec8* xr = reinterpret_cast<const Vec8*>(x + row * x_stride);
  float vals[kMaxVec][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int v = threadIdx.x + i * kNormThreads;
    if (v < nvec) {
      vec8_to_f32<T>(xr[v], vals[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += vals[i][j];
    }
  }
  const float mean = block_sum(s, red) / static_cast<float>(dim);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int v = threadIdx.x + i * kNormThreads;
    if (v < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = vals[i][j] - mean;
        ss += d * d;
      }
    }
  }
  __syncthreads();  // `red` is reused
  const float inv = rsqrtf(block_sum(ss, red) / static_cast<float>(dim) + eps);
  Vec8* orow = reinterpret_cast<Vec8*>(out + row * out_stride);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int v = threadIdx.x + i * kNormThreads;
    if (v < nvec) {
      float g[8], b[8], o[8];
      vec8_to_f32<T>(reinterpret_cast<const Vec8*>(gamma)[v], g);
      vec8_to_f32<T>(reinterpret_cast<const Vec8*>(beta)[v], b);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (vals[i][j] - mean) * inv * g[j] + b[j];
      orow[v] = f32_to_vec8<T>(o);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
bias_act_residual_kernel(T* __restrict__ out, const T* __restrict__ x, const T* __restrict__ bias,
                         const T* __restrict__ residual, int64_t tokens, int64_t dim, int act) {
  // PDL: release the dependents first (they block in their own griddepcontrol.wait until THIS grid
  // has completed, and a grid completes only after its own wait — so order is transitive), then
  // wait for the predecessor.  A decode GEMM two kernels downstream can then take over an SM the
  // moment the previous GEMM's CTA leaves it and stream its weights under this kernel.
  griddep_launch_dependents();
  griddep_wait();
  const int64_t nvec_row = dim >> 3;
  const int64_t total = tokens * nvec_row;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t c = i % nvec_row;
    float f[8];
    vec8_to_f32<T>(reinterpret_cast<const Vec8*>(x)[i], f);
    if (bias) {
      float b[8];
      vec8_to_f32<T>(reinterpret_cast<const Vec8*>(bias)[c], b);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += b[j];
    }
    if (act == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    if (residual) {
      float r[8];
      vec8_to_f32<T>(reinterpret_cast<const Vec8*>(residual)[i], r);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += r[j];
    }
    reinterpret_cast<Vec8*>(out)[i] = f32_to_vec8<T>(f);
  }
}

}  // namespace spd

using namespace spd;

extern "C" {

int spd_rmsnorm(void* out, const void* x, const void* w, int64_t tokens, int64_t dim,
                int64_t x_stride, int64_t out_stride, float eps, int dtype, spd_stream_t stream) {
  SPD_REQUIRE(out && x && w, SPD_E_BADARG, "spd_rmsnorm: null pointer");
  SPD_REQUIRE(tokens >= 0 && dim > 0, SPD_E_BADARG, "spd_rmsnorm: bad sizes tokens=%lld dim=%lld",
              (long long)tokens, (long long)dim);
  SPD_REQUIRE(dim % 8 == 0 && x_stride % 8 == 0 && out_stride % 8 == 0 && aligned16(out) &&
                  aligned16(x) && aligned16(w),
              SPD_E_ALIGN, "spd_rmsnorm: dim/strides must be multiples of 8 and pointers 16-B aligned");
  if (tokens == 0) return 0;
#define SPD_NORM_LAUNCH(FUSED, VEC, ...)                                                         \
  SPD_CUDA(launch(rmsnorm_kernel<T, FUSED, VEC>, dim3((unsigned)tokens), dim3(kNormThreads), 0, \
                  as_stream(stream), pdl_enabled(), __VA_ARGS__))
#define SPD_NORM_DISPATCH(FUSED, ...)                                         \
  do {                                                                        \
    if (dim <= 2 * 8 * kNormThreads) SPD_NORM_LAUNCH(FUSED, 2, __VA_ARGS__);  \
    else if (dim <= 4 * 8 * kNormThreads) SPD_NORM_LAUNCH(FUSED, 4, __VA_ARGS__); \
    else SPD_NORM_LAUNCH(FUSED, 8, __VA_ARGS__);                              \
  } while (0)
  SPD_DISPATCH_DTYPE(dtype, {
    SPD_NORM_DISPATCH(false, (T*)out, (T*)x, (T*)nullptr, (const T*)w, dim, x_stride, out_stride,
                      eps);
  });
  return 0;
}

int spd_fused_add_rmsnorm(void* x, void* residual, const void* w, int64_t tokens, int64_t dim,
                          float eps, int dtype, spd_stream_t stream) {
  SPD_REQUIRE(x && residual && w, SPD_E_BADARG, "spd_fused_add_rmsnorm: null pointer");
  SPD_REQUIRE(tokens >= 0 && dim > 0, SPD_E_BADARG, "spd_fused_add_rmsnorm: bad sizes");
  SPD_REQUIRE(dim % 8 == 0 && aligned16(x) && aligned16(residual) && aligned16(w), SPD_E_ALIGN,
              "spd_fused_add_rmsnorm: dim must be a multiple of 8 and pointers 16-B aligned");
  if (tokens == 0) return 0;
  SPD_DISPATCH_DTYPE(dtype, {
    SPD_NORM_DISPATCH(true, (T*)x, (T*)x, (T*)residual, (const T*)w, dim, dim, dim, eps);
  });
  return 0;
}

int spd_silu_and_mul(void* out, const void* x, int64_t tokens, int64_t d, int dtype,
                     spd_stream_t stream) {
  SPD_REQUIRE(out && x, SPD_E_BADARG, "spd_silu_and_mul: null pointer");
  SPD_REQUIRE(tokens >= 0 && d > 0, SPD_E_BADARG, "spd_silu_and_mul: bad sizes");
  SPD_REQUIRE(d % 8 == 0 && aligned16(out) && aligned16(x), SPD_E_ALIGN,
              "spd_silu_and_mul: The pointers must be multiple of 16 bytes.");
  if (tokens == 0) return 0;
  const int64_t total = tokens * (d >> 3);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(device_sm_count_cached()) * 16;
  if (blocks > cap) blocks = cap;
  SPD_DISPATCH_DTYPE(dtype, {
    SPD_CUDA(launch(silu_and_mul_kernel<T>, dim3((unsigned)blocks), dim3(256), 0, as_stream(stream),
                    pdl_enabled(), (T*)out, (const T*)x, tokens, d));
  });
  return 0;
}

static int rope_common(void* q, void* k, const void* v, void* k_pool, void* v_pool,
                       const float* cos_sin_cache, const int64_t* pos_ids, const int64_t* loc,
                       int64_t tokens, int Hq, int Hk, int D, int rot, int64_t q_stride,
                       int64_t k_stride, int64_t v_stride, int interleave, int dtype,
                       spd_stream_t stream, bool write_pool) {
  SPD_REQUIRE(q && k && cos_sin_cache && pos_ids, SPD_E_BADARG, "spd_rope: null pointer");
  if (write_pool) SPD_REQUIRE(v && k_pool && v_pool && loc, SPD_E_BADARG, "spd_rope_kvwrite: null pointer");
  SPD_REQUIRE(tokens >= 0 && Hq > 0 && Hk > 0 && D > 0 && rot > 0 && rot <= D, SPD_E_BADARG,
              "spd_rope: bad sizes");
  SPD_REQUIRE(D % 8 == 0 && rot % 16 == 0 && q_stride % 8 == 0 && k_stride % 8 == 0 &&
                  (!write_pool || v_stride % 8 == 0),
              SPD_E_ALIGN, "spd_rope: head_dim %% 8, rot_dim %% 16 and strides %% 8 must be 0");
  SPD_REQUIRE(aligned16(q) && aligned16(k) && aligned16(cos_sin_cache) &&
                  (!write_pool || (aligned16(v) && aligned16(k_pool) && aligned16(v_pool))),
              SPD_E_ALIGN, "spd_rope: pointers must be 16-B aligned");
  if (tokens == 0) return 0;
  SPD_DISPATCH_DTYPE(dtype, {
    if (write_pool) {
      SPD_CUDA(launch(rope_kernel<T, true>, dim3((unsigned)tokens), dim3(256), 0, as_stream(stream),
                      pdl_enabled(), (T*)q, (T*)k, (const T*)v, (T*)k_pool, (T*)v_pool,
                      cos_sin_cache, pos_ids, loc, Hq, Hk, D, rot, q_stride, k_stride, v_stride,
                      interleave));
    } else {
      SPD_CUDA(launch(rope_kernel<T, false>, dim3((unsigned)tokens), dim3(256), 0, as_stream(stream),
                      pdl_enabled(), (T*)q, (T*)k, (const T*)nullptr, (T*)nullptr, (T*)nullptr,
                      cos_sin_cache, pos_ids, (const int64_t*)nullptr, Hq, Hk, D, rot, q_stride,
                      k_stride, (int64_t)0, interleave));
    }
  });
  return 0;
}

int spd_rope_inplace(void* q, void* k, const float* cos_sin_cache, const int64_t* pos_ids,
                     int64_t tokens, int num_q_heads, int num_k_heads, int head_dim, int rot_dim,
                     int64_t q_stride, int64_t k_stride, int interleave, int dtype,
                     spd_stream_t stream) {
  return rope_common(q, k, nullptr, nullptr, nullptr, cos_sin_cache, pos_ids, nullptr, tokens,
                     num_q_heads, num_k_heads, head_dim, rot_dim, q_stride, k_stride, 0, interleave,
                     dtype, stream, false);
}

int spd_rope_kvwrite(void* q, void* k, const void* v, void* k_pool, void* v_pool,
                     const float* cos_sin_cache, const int64_t* pos_ids, const int64_t* loc,
                     int64_t tokens, int num_q_heads, int num_k_heads, int head_dim, int rot_dim,
                     int64_t q_stride, int64_t k_stride, int64_t v_stride, int interleave,
                     int dtype, spd_stream_t stream) {
  return rope_common(q, k, v, k_pool, v_pool, cos_sin_cache, pos_ids, loc, tokens, num_q_heads,
                     num_k_heads, head_dim, rot_dim, q_stride, k_stride, v_stride, interleave,
                     dtype, stream, true);
}

int spd_kv_write(const void* k, const void* v, void* k_pool, void* v_pool, const int64_t* loc,
                 int64_t tokens, int64_t row_elems, int64_t k_stride, int64_t v_stride, int dtype,
                 spd_stream_t stream) {
  SPD_REQUIRE(k && k_pool && loc && ((v == nullptr) == (v_pool == nullptr)), SPD_E_BADARG,
              "spd_kv_write: null pointer (v and v_pool may both be NULL for a single-buffer pool)");
  SPD_REQUIRE(tokens >= 0 && row_elems > 0, SPD_E_BADARG, "spd_kv_write: bad sizes");
  SPD_REQUIRE(row_elems % 8 == 0 && k_stride % 8 == 0 && v_stride % 8 == 0 && aligned16(k) &&
                  aligned16(v) && aligned16(k_pool) && aligned16(v_pool),
              SPD_E_ALIGN, "spd_kv_write: rows must be multiples of 8 elements, pointers 16-B aligned");
  if (tokens == 0) return 0;
  SPD_DISPATCH_DTYPE(dtype, {
    SPD_CUDA(launch(kv_write_kernel<T>, dim3((unsigned)tokens), dim3(128), 0, as_stream(stream),
                    pdl_enabled(), (const T*)k, (const T*)v, (T*)k_pool, (T*)v_pool, loc, row_elems,
                    k_stride, v_stride));
  });
  return 0;
}

int spd_gather_rows(void* out, const void* table, const int64_t* idx, int64_t n, int64_t row_elems,
                    int64_t table_rows, int dtype, spd_stream_t stream) {
  SPD_REQUIRE(out && table && idx, SPD_E_BADARG, "spd_gather_rows: null pointer");
  SPD_REQUIRE(n >= 0 && row_elems > 0 && table_rows > 0, SPD_E_BADARG, "spd_gather_rows: bad sizes");
  SPD_REQUIRE(row_elems % 8 == 0 && aligned16(out) && aligned16(table), SPD_E_ALIGN,
              "spd_gather_rows: rows must be multiples of 8 elements, pointers 16-B aligned");
  if (n == 0) return 0;
  SPD_DISPATCH_DTYPE(dtype, {
    SPD_CUDA(launch(gather_rows_kernel<T>, dim3((unsigned)n), dim3(256), 0, as_stream(stream),
                    pdl_enabled(), (T*)out, (const T*)table, idx, row_elems, table_rows));
  });
  return 0;
}

int spd_argmax_f32(int32_t* ids, const float* logits, int64_t batch, int64_t vocab, int64_t ld,
                   spd_stream_t stream) {
  SPD_REQUIRE(ids && logits, SPD_E_BADARG, "spd_argmax_f32: null pointer");
  SPD_REQUIRE(batch >= 0 && vocab > 0 && ld >= vocab && vocab < 0x7fffffff, SPD_E_BADARG,
              "spd_argmax_f32: bad sizes");
  if (batch == 0) return 0;
  SPD_CUDA(launch(argmax_kernel, dim3((unsigned)batch), dim3(kArgmaxThreads), 0, as_stream(stream),
                  pdl_enabled(), ids, logits, vocab, ld, static_cast<int4*>(nullptr), 0));
  return 0;
}

int spd_argmax_partial_f32(void* pairs, const float* logits, int64_t batch, int64_t vocab, int64_t ld,
                           int64_t index_offset, spd_stream_t stream) {
  SPD_REQUIRE(pairs && logits, SPD_E_BADARG, "spd_argmax_partial_f32: null pointer");
  SPD_REQUIRE(aligned16(pairs), SPD_E_ALIGN, "spd_argmax_partial_f32: pairs must be 16-byte aligned");
  SPD_REQUIRE(batch >= 0 && vocab > 0 && ld >= vocab && index_offset >= 0 &&
                  index_offset + vocab < 0x7fffffff, SPD_E_BADARG, "spd_argmax_partial_f32: bad sizes");
  if (batch == 0) return 0;
  SPD_CUDA(launch(argmax_kernel, dim3((unsigned)batch), dim3(kArgmaxThreads), 0, as_stream(stream),
                  pdl_enabled(), static_cast<int32_t*>(nullptr), logits, vocab, ld,
                  static_cast<int4*>(pairs), static_cast<int>(index_offset)));
  return 0;
}

int spd_argmax_merge(int32_t* ids, const void* pairs, int parts, int64_t batch, spd_stream_t stream) {
  SPD_REQUIRE(ids && pairs, SPD_E_BADARG, "spd_argmax_merge: null pointer");
  SPD_REQUIRE(aligned16(pairs), SPD_E_ALIGN, "spd_argmax_merge: pairs must be 16-byte aligned");
  SPD_REQUIRE(parts > 0 && batch >= 0, SPD_E_BADARG, "spd_argmax_merge: bad sizes");
  if (batch == 0) return 0;
  SPD_CUDA(launch(argmax_merge_kernel, dim3((unsigned)((batch + 127) / 128)), dim3(128), 0,
                  as_stream(stream), pdl_enabled(), ids, static_cast<const int4*>(pairs), parts, batch));
  return 0;
}

int spd_compute_positions(int64_t* positions, int32_t* extend_start_loc, int32_t* token_req,
                          const int32_t* prefix_lens, const int32_t* extend_lens, int64_t batch,
                          spd_stream_t stream) {
  SPD_REQUIRE(positions && extend_start_loc && extend_lens, SPD_E_BADARG,
              "spd_compute_positions: null pointer");
  SPD_REQUIRE(batch >= 0, SPD_E_BADARG, "spd_compute_positions: bad batch");
  if (batch == 0) return 0;
  compute_positions_kernel<<<(unsigned)batch, 256, 0, as_stream(stream)>>>(
      positions, extend_start_loc, token_req, prefix_lens, extend_lens);
  SPD_CHECK_LAUNCH("compute_positions_kernel");
  return 0;
}

int spd_decode_positions(int64_t* positions, const int64_t* seq_lens, int64_t batch,
                         spd_stream_t stream) {
  SPD_REQUIRE(positions && seq_lens, SPD_E_BADARG, "spd_decode_positions: null pointer");
  SPD_REQUIRE(batch >= 0, SPD_E_BADARG, "spd_decode_positions: bad batch");
  if (batch == 0) return 0;
  SPD_CUDA(launch(decode_positions_kernel, dim3((unsigned)((batch + 127) / 128)), dim3(128), 0,
                  as_stream(stream), pdl_enabled(), positions, seq_lens, batch));
  return 0;
}

int spd_build_kv_indices(int32_t* kv_indices, const int32_t* req_to_token,
                         const int64_t* req_pool_indices, const int32_t* lens,
                         const int32_t* kv_indptr, const int32_t* kv_start, int64_t batch,
                         int64_t req_to_token_stride, spd_stream_t stream) {
  SPD_REQUIRE(kv_indices && req_to_token && req_pool_indices && lens && kv_indptr, SPD_E_BADARG,
              "spd_build_kv_indices: null pointer");
  SPD_REQUIRE(batch >= 0 && req_to_token_stride > 0, SPD_E_BADARG, "spd_build_kv_indices: bad sizes");
  if (batch == 0) return 0;
  build_kv_indices_kernel<<<(unsigned)batch, 256, 0, as_stream(stream)>>>(
      kv_indices, req_to_token, req_pool_indices, lens, kv_indptr, kv_start, req_to_token_stride);
  SPD_CHECK_LAUNCH("build_kv_indices_kernel");
  return 0;
}

int spd_write_req_to_token(int32_t* req_to_token, const int64_t* req_pool_indices,
                           const int64_t* pre_lens, const int64_t* seq_lens,
                           const int64_t* extend_lens, const int64_t* out_cache_loc, int64_t batch,
                           int64_t req_to_token_stride, spd_stream_t stream) {
  SPD_REQUIRE(req_to_token && req_pool_indices && pre_lens && seq_lens && extend_lens && out_cache_loc,
              SPD_E_BADARG, "spd_write_req_to_token: null pointer");
  SPD_REQUIRE(batch >= 0 && req_to_token_stride > 0, SPD_E_BADARG, "spd_write_req_to_token: bad sizes");
  if (batch == 0) return 0;
  write_req_to_token_kernel<<<(unsigned)batch, 256, 0, as_stream(stream)>>>(
      req_to_token, req_pool_indices, pre_lens, seq_lens, extend_lens, out_cache_loc,
      req_to_token_stride);
  SPD_CHECK_LAUNCH("write_req_to_token_kernel");
  return 0;
}

int spd_layernorm(void* out, const void* x, const void* gamma, const void* beta, int64_t tokens, int64_t dim,
                  int64_t x_stride, int64_t out_stride, float eps, int dtype, spd_stream_t stream) {
  SPD_REQUIRE(out && x && gamma && beta, SPD_E_BADARG, "spd_layernorm: null pointer");
  SPD_REQUIRE(tokens >= 0 && dim > 0 && dim <= 8192, SPD_E_SHAPE,
              "spd_layernorm: bad sizes tokens=%lld dim=%lld (dim <= 8192)", (long long)tokens, (long long)dim);
  SPD_REQUIRE(dim % 8 == 0 && x_stride % 8 == 0 && out_stride % 8 == 0 && aligned16(out) && aligned16(x) &&
                  aligned16(gamma) && aligned16(beta),
              SPD_E_ALIGN, "spd_layernorm: dim/strides must be multiples of 8 and pointers 16-B aligned");
  if (tokens == 0) return 0;
  SPD_DISPATCH_DTYPE(dtype, {
    SPD_CUDA(launch(layernorm_kernel<T, 4>, dim3((unsigned)tokens), dim3(kNormThreads), 0, as_stream(stream),
                    pdl_enabled(), (T*)out, (const T*)x, (const T*)gamma, (const T*)beta, dim, x_stride,
                    out_stride, eps));
  });
  return 0;
}

int spd_bias_act_residual(void* out, const void* x, const void* bias, const void* residual, int64_t tokens,
                          int64_t dim, int act, int dtype, spd_stream_t stream) {
  SPD_REQUIRE(out && x, SPD_E_BADARG, "spd_bias_act_residual: null pointer");
  SPD_REQUIRE(tokens >= 0 && dim > 0 && (act == 0 || act == 1), SPD_E_BADARG, "spd_bias_act_residual: bad arguments");
  SPD_REQUIRE(dim % 8 == 0 && aligned16(out) && aligned16(x) && (!bias || aligned16(bias)) &&
                  (!residual || aligned16(residual)),
              SPD_E_ALIGN, "spd_bias_act_residual: dim must be a multiple of 8 and pointers 16-B aligned");
  if (tokens == 0) return 0;
  int64_t blocks = (tokens * (dim >> 3) + 255) / 256;
  const int64_t cap = static_cast<int64_t>(device_sm_count_cached()) * 16;
  if (blocks > cap) blocks = cap;
  SPD_DISPATCH_DTYPE(dtype, {
    SPD_CUDA(launch(bias_act_residual_kernel<T>, dim3((unsigned)blocks), dim3(256), 0, as_stream(stream),
                    pdl_enabled(), (T*)out, (const T*)x, (const T*)bias, (const T*)residual, tokens, dim, act));
  });
  return 0;
}

}  // extern "C"
