// Mixture-of-experts plumbing for the DeepSeek-V2-Lite path (BASELINE configs[3]): routing,
// token -> expert alignment, gather / combine, and the dense-decode activation.  The expert
// GEMMs themselves run on the tcgen05 kernels of gemm.cu (spd_gemm_grouped rows mode for
// prefill-sized batches, plain spd_gemm over the stacked expert weights for decode-sized ones).
//
// Reference: DeepseekV2MoE.forward (python/sglang/srt/models/deepseek_v2.py:194-218) →
//   router       F.linear + softmax + greedy top-k, no renormalisation for V2-Lite
//                (python/sglang/srt/layers/moe/topk.py:79-120)
//   fused experts Triton fused_moe_kernel x 2 over tokens sorted into expert-aligned blocks
//                (layers/moe/fused_moe_triton/fused_moe.py:961-1168) with
//                moe_align_block_size (sgl-kernel/csrc/moe/moe_align_kernel.cu:28,44)
//   x routed_scaling_factor + shared experts (deepseek_v2.py:205-218)
#include "common.cuh"
#include "ptx.cuh"

namespace spd {

struct __align__(16) MVec8 {
  uint32_t w[4];
};

constexpr int kMoeMaxExperts = 256;
constexpr int kMoeMaxTopK = 8;
constexpr int kMoeTile = 128;  // rows per expert-aligned tile == UMMA M of the grouped GEMM

// One warp per token: softmax over E logits (fp32), then top-k by repeated arg-max (ties -> the
// smaller expert id).  weights = prob (optionally renormalised over the k) x scale.
template <typename T>
__global__ void __launch_bounds__(128)
moe_route_kernel(int32_t* __restrict__ topk_ids, float* __restrict__ topk_w, const T* __restrict__ logits,
                 int64_t tokens, int E, int k, float scale, int renorm) {
  griddep_wait();
  const int lane = threadIdx.x & 31;
  const int64_t t = blockIdx.x * 4ll + (threadIdx.x >> 5);
  if (t >= tokens) return;
  constexpr int PER = kMoeMaxExperts / 32;
  float v[PER];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int e = i * 32 + lane;
    v[i] = e < E ? to_f32<T>(logits[t * E + e]) : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    v[i] = (i * 32 + lane) < E ? __expf(v[i] - mx) : -1.f;  // -1: not an expert / already taken
    sum += fmaxf(v[i], 0.f);
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  float picked_w[kMoeMaxTopK];
  int picked_e[kMoeMaxTopK];
  float wsum = 0.f;
#pragma unroll
  for (int j = 0; j < kMoeMaxTopK; ++j) {
    if (j < k) {
      float best = -1.f;
      int be = 0x7fffffff;
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int e = i * 32 + lane;
        if (v[i] > best) { best = v[i]; be = e; }  // ascending e within a lane: first max wins
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oe = __shfl_xor_sync(0xffffffffu, be, o);
        if (ob > best || (ob == best && oe < be)) { best = ob; be = oe; }
      }
      picked_e[j] = be;
      picked_w[j] = best * inv;
      wsum += picked_w[j];
#pragma unroll
      for (int i = 0; i < PER; ++i)
        if (i * 32 + lane == be) v[i] = -1.f;
    }
  }
  if (lane == 0) {
    const float r = renorm ? scale / wsum : scale;
#pragma unroll
    for (int j = 0; j < kMoeMaxTopK; ++j)
      if (j < k) {
        topk_ids[t * k + j] = picked_e[j];
        topk_w[t * k + j] = picked_w[j] * r;
      }
  }
}

// Single CTA: rows sorted by expert, every expert's rows padded to whole 128-row tiles.
//   row_token[r]  token feeding sorted row r (-1 = padding row)
//   inv_pos[i]    sorted row of routing entry i = token * k + j
//   tile_group[x] expert of 128-row tile x (-1 = unused tile)
__global__ void __launch_bounds__(1024)
moe_align_kernel(int32_t* __restrict__ row_token, int32_t* __restrict__ inv_pos, int32_t* __restrict__ tile_group,
                 const int32_t* __restrict__ topk_ids, int64_t entries, int k, int E, int64_t max_rows) {
  __shared__ int cnt[kMoeMaxExperts];
  __shared__ int off[kMoeMaxExperts + 1];
  __shared__ int cur[kMoeMaxExperts];
  griddep_wait();
  const int tid = threadIdx.x;
  for (int e = tid; e < E; e += 1024) { cnt[e] = 0; cur[e] = 0; }
  for (int64_t r = tid; r < max_rows; r += 1024) row_token[r] = -1;
  const int64_t max_tiles = max_rows / kMoeTile;
  for (int64_t x = tid; x < max_tiles; x += 1024) tile_group[x] = -1;
  __syncthreads();
  for (int64_t i = tid; i < entries; i += 1024) atomicAdd(&cnt[topk_ids[i]], 1);
  __syncthreads();
  if (tid == 0) {
    int o = 0;
    for (int e = 0; e < E; ++e) {
      off[e] = o;
      o += (cnt[e] + kMoeTile - 1) / kMoeTile * kMoeTile;
    }
    off[E] = o;
  }
  __syncthreads();
  for (int e = tid; e < E; e += 1024)
    for (int x = off[e] / kMoeTile; x < off[e + 1] / kMoeTile; ++x) tile_group[x] = e;
  for (int64_t i = tid; i < entries; i += 1024) {
    const int e = topk_ids[i];
    const int pos = off[e] + atomicAdd(&cur[e], 1);
    inv_pos[i] = pos;
    row_token[pos] = static_cast<int32_t>(i / k);
  }
}

// xs[r, :] = x[row_token[r], :]  (zeros for padding rows)
template <typename T>
__global__ void __launch_bounds__(256)
moe_gather_kernel(T* __restrict__ xs, const T* __restrict__ x, const int32_t* __restrict__ row_token, int64_t dim,
                  int64_t x_stride) {
  griddep_wait();
  const int64_t r = blockIdx.x;
  const int tok = row_token[r];
  MVec8* dst = reinterpret_cast<MVec8*>(xs + r * dim);
  const int nv = static_cast<int>(dim >> 3);
  if (tok < 0) {
    MVec8 z;
    z.w[0] = z.w[1] = z.w[2] = z.w[3] = 0u;
    for (int i = threadIdx.x; i < nv; i += 256) dst[i] = z;
  } else {
    const MVec8* src = reinterpret_cast<const MVec8*>(x + tok * x_stride);
    for (int i = threadIdx.x; i < nv; i += 256) dst[i] = src[i];
  }
}

// out[t, :] = (shared ? shared[t, :] : 0) + sum_j w[t, j] * y[inv_pos[t*k + j], :]   (fp32 sum, j order)
template <typename T>
__global__ void __launch_bounds__(256)
moe_combine_kernel(T* __restrict__ out, const T* __restrict__ y, const T* __restrict__ shared,
                   const int32_t* __restrict__ inv_pos, const float* __restrict__ topk_w, int k, int64_t dim) {
  griddep_wait();
  const int64_t t = blockIdx.x;
  const int nv = static_cast<int>(dim >> 3);
  for (int i = threadIdx.x; i < nv; i += 256) {
    float acc[8];
    if (shared) {
      const MVec8 s = reinterpret_cast<const MVec8*>(shared + t * dim)[i];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = unpack2<T>(s.w[q]);
        acc[2 * q] = f.x;
        acc[2 * q + 1] = f.y;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    }
    for (int j = 0; j < k; ++j) {
      const float w = topk_w[t * k + j];
      const MVec8 v = reinterpret_cast<const MVec8*>(y + static_cast<int64_t>(inv_pos[t * k + j]) * dim)[i];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = unpack2<T>(v.w[q]);
        acc[2 * q] += w * f.x;
        acc[2 * q + 1] += w * f.y;
      }
    }
    MVec8 o;
#pragma unroll
    for (int q = 0; q < 4; ++q) o.w[q] = pack2<T>(acc[2 * q], acc[2 * q + 1]);
    reinterpret_cast<MVec8*>(out + t * dim)[i] = o;
  }
}

// Decode-sized batches run ALL experts as two plain GEMMs over the stacked weights (every expert's
// weights are streamed anyway once ~16 tokens x top-6 are spread over 64 experts); this kernel
// sits between them: act[t, e*I + i] = w(t, e) * silu(gate) * up, w(t, e) = 0 for unselected e.
template <typename T>
__global__ void __launch_bounds__(256)
moe_dense_act_kernel(T* __restrict__ act, const T* __restrict__ gate_up, const int32_t* __restrict__ topk_ids,
                     const float* __restrict__ topk_w, int k, int E, int64_t inter) {
  __shared__ float w_e[kMoeMaxExperts];
  griddep_wait();
  const int64_t t = blockIdx.x;
  for (int e = threadIdx.x; e < E; e += 256) w_e[e] = 0.f;
  __syncthreads();
  if (threadIdx.x < k) w_e[topk_ids[t * k + threadIdx.x]] = topk_w[t * k + threadIdx.x];
  __syncthreads();
  const int nv = static_cast<int>(inter >> 3);
  const T* src = gate_up + t * (2 * inter * E);
  T* dst = act + t * (inter * E);
  for (int idx = threadIdx.x; idx < E * nv; idx += 256) {
    const int e = idx / nv, i = idx - e * nv;
    const float w = w_e[e];
    MVec8 o;
    if (w == 0.f) {
      o.w[0] = o.w[1] = o.w[2] = o.w[3] = 0u;
    } else {
      const MVec8 g = reinterpret_cast<const MVec8*>(src + e * 2 * inter)[i];
      const MVec8 u = reinterpret_cast<const MVec8*>(src + e * 2 * inter + inter)[i];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 gf = unpack2<T>(g.w[q]), uf = unpack2<T>(u.w[q]);
        // silu(gate) * up rounded to T (what the expert's down projection reads), then the weight
        const float a = to_f32<T>(from_f32<T>(silu_f32(gf.x) * uf.x));
        const float b = to_f32<T>(from_f32<T>(silu_f32(gf.y) * uf.y));
        o.w[q] = pack2<T>(w * a, w * b);
      }
    }
    reinterpret_cast<MVec8*>(dst + e * inter)[i] = o;
  }
}

}  // namespace spd

using namespace spd;

extern "C" {

int spd_moe_route(int32_t* topk_ids, float* topk_w, const void* logits, int64_t tokens, int num_experts,
                  int top_k, float scale, int renormalize, int dtype, spd_stream_t stream) {
  SPD_REQUIRE(topk_ids && topk_w && logits, SPD_E_BADARG, "spd_moe_route: null pointer");
  SPD_REQUIRE(tokens >= 0 && num_experts > 0 && num_experts <= kMoeMaxExperts && top_k > 0 &&
                  top_k <= kMoeMaxTopK && top_k <= num_experts,
              SPD_E_SHAPE, "spd_moe_route: need experts <= %d and top_k <= %d", kMoeMaxExperts, kMoeMaxTopK);
  if (tokens == 0) return 0;
  SPD_DISPATCH_DTYPE(dtype, {
    SPD_CUDA(launch(moe_route_kernel<T>, dim3((unsigned)((tokens + 3) / 4)), dim3(128), 0, as_stream(stream),
                    pdl_enabled(), topk_ids, topk_w, (const T*)logits, tokens, num_experts, top_k, scale,
                    renormalize));
  });
  return 0;
}

int64_t spd_moe_align_rows(int64_t tokens, int top_k, int num_experts) {
  // sum over experts of ceil128(count) <= entries + experts * 127, rounded up to whole tiles
  const int64_t bound = tokens * top_k + static_cast<int64_t>(num_experts) * (kMoeTile - 1);
  return (bound + kMoeTile - 1) / kMoeTile * kMoeTile;
}

int spd_moe_align(int32_t* row_token, int32_t* inv_pos, int32_t* tile_group, const int32_t* topk_ids,
                  int64_t tokens, int top_k, int num_experts, int64_t max_rows, spd_stream_t stream) {
  SPD_REQUIRE(row_token && inv_pos && tile_group && topk_ids, SPD_E_BADARG, "spd_moe_align: null pointer");
  SPD_REQUIRE(num_experts > 0 && num_experts <= kMoeMaxExperts && top_k > 0, SPD_E_SHAPE, "spd_moe_align: bad sizes");
  SPD_REQUIRE(max_rows >= spd_moe_align_rows(tokens, top_k, num_experts) && max_rows % kMoeTile == 0,
              SPD_E_WORKSPACE, "spd_moe_align: max_rows too small (need spd_moe_align_rows())");
  SPD_CUDA(launch(moe_align_kernel, dim3(1), dim3(1024), 0, as_stream(stream), pdl_enabled(), row_token, inv_pos,
                  tile_group, topk_ids, tokens * top_k, top_k, num_experts, max_rows));
  return 0;
}

int spd_moe_gather(void* xs, const void* x, const int32_t* row_token, int64_t rows, int64_t dim,
                   int64_t x_stride, int dtype, spd_stream_t stream) {
  SPD_REQUIRE(xs && x && row_token, SPD_E_BADARG, "spd_moe_gather: null pointer");
  SPD_REQUIRE(dim > 0 && dim % 8 == 0 && x_stride % 8 == 0 && aligned16(xs) && aligned16(x), SPD_E_ALIGN,
              "spd_moe_gather: rows must be multiples of 8 elements, pointers 16-B aligned");
  if (rows == 0) return 0;
  SPD_DISPATCH_DTYPE(dtype, {
    SPD_CUDA(launch(moe_gather_kernel<T>, dim3((unsigned)rows), dim3(256), 0, as_stream(stream), pdl_enabled(),
                    (T*)xs, (const T*)x, row_token, dim, x_stride));
  });
  return 0;
}

int spd_moe_combine(void* out, const void* y, const void* shared, const int32_t* inv_pos, const float* topk_w,
                    int64_t tokens, int top_k, int64_t dim, int dtype, spd_stream_t stream) {
  SPD_REQUIRE(out && y && inv_pos && topk_w, SPD_E_BADARG, "spd_moe_combine: null pointer");
  SPD_REQUIRE(dim > 0 && dim % 8 == 0 && aligned16(out) && aligned16(y) && (!shared || aligned16(shared)),
              SPD_E_ALIGN, "spd_moe_combine: rows must be multiples of 8 elements, pointers 16-B aligned");
  if (tokens == 0) return 0;
  SPD_DISPATCH_DTYPE(dtype, {
    SPD_CUDA(launch(moe_combine_kernel<T>, dim3((unsigned)tokens), dim3(256), 0, as_stream(stream), pdl_enabled(),
                    (T*)out, (const T*)y, (const T*)shared, inv_pos, topk_w, top_k, dim));
  });
  return 0;
}

int spd_moe_dense_act(void* act, const void* gate_up, const int32_t* topk_ids, const float* topk_w,
                      int64_t tokens, int top_k, int num_experts, int64_t inter, int dtype,
                      spd_stream_t stream) {
  SPD_REQUIRE(act && gate_up && topk_ids && topk_w, SPD_E_BADARG, "spd_moe_dense_act: null pointer");
  SPD_REQUIRE(num_experts > 0 && num_experts <= kMoeMaxExperts && top_k > 0 && top_k <= kMoeMaxTopK, SPD_E_SHAPE,
              "spd_moe_dense_act: bad sizes");
  SPD_REQUIRE(inter > 0 && inter % 8 == 0 && aligned16(act) && aligned16(gate_up), SPD_E_ALIGN,
              "spd_moe_dense_act: inter must be a multiple of 8, pointers 16-B aligned");
  if (tokens == 0) return 0;
  SPD_DISPATCH_DTYPE(dtype, {
    SPD_CUDA(launch(moe_dense_act_kernel<T>, dim3((unsigned)tokens), dim3(256), 0, as_stream(stream),
                    pdl_enabled(), (T*)act, (const T*)gate_up, topk_ids, topk_w, top_k, num_experts, inter));
  });
  return 0;
}

}  // extern "C"
