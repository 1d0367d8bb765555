// Absorbed multi-head latent attention (MLA) over the token-granular latent KV pool — decode and
// ragged extend in one kernel.  DeepSeek-V2-Lite, BASELINE configs[3].
//
// Reference: DeepseekV2AttentionMLA.forward_absorb (python/sglang/srt/models/deepseek_v2.py:633-706)
// over MLATokenToKVPool rows [kv_a_normed(kv_lora) || k_pe_roped(rope)] (mem_cache/memory_pool.py:
// 379-452), computed there by the Triton kernels decode_attention_fwd_grouped / extend_attention_fwd
// (layers/attention/triton_ops/decode_attention.py:235-470, extend_attention.py:291-410): all H
// query heads share ONE latent K/V stream (MQA), K = the whole latent row (DL = kv_lora + rope),
// V = its first DV = kv_lora columns.  The reference switches to a plain MHA (forward_normal) for
// prefill batches without a cached prefix; the absorbed form used here for every batch is the same
// function of the same inputs (oracle/deepseek_oracle.py checks both against each other).
//
// One CTA = one query token x one group of 16 heads x one KV split.  Both products run on the
// tensor cores through warp-level mma.sync m16n8k16 (M = the 16 heads): this is the first MLA
// kernel of the repo; the tcgen05 swap-AB form of attn_decode.cu with N = 16 real heads is the
// planned successor.  Per 32-token chunk:
//   K rows (= V rows) -> smem by 16-byte cp.async through req_to_token, double buffered;
//   S[16 x 32]  = Q[16 x DL] K^T      k-steps split over the 8 warps, partials reduced in smem;
//   online softmax, one warp per pair of heads, exp2 domain;
//   O[16 x DV] += P[16 x 32] V        warp w owns output columns [64 w, 64 w + 64).
#include "common.cuh"
#include "ptx.cuh"

namespace spd {

constexpr int kMlaThreads = 256;
constexpr int kMlaChunk = 32;
constexpr int kMlaHeads = 16;
constexpr int kMlaMaxSplits = 8;

struct MlaParams {
  void* out;             // [items, H, DV]
  const void* q;         // [items, H, DL]
  const void* kv_pool;   // [slots, DL]
  const int32_t* req_to_token;
  const int64_t* req_pool_indices;
  const int64_t* seq_lens;          // decode (extend: nullptr)
  const int32_t* prefix_lens;       // extend
  const int32_t* extend_start_loc;  // extend
  int batch, H, splits;
  int64_t q_stride, out_stride, rtt_stride;
  float scale_log2;
  float* o_part;   // [item][H][splits][DV]
  float* ml_part;  // [item][H][splits][2]
};

template <typename T>
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma_16816<__nv_bfloat16>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma_16816<__half>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// DL, DV: latent row width and value width (elements).  Row stride in smem = DL + 8 elements:
// (DL + 8) / 2 = 292 words = 4 (mod 32), so the 8 rows x 4 words an MMA fragment load touches fall
// into 32 distinct banks.
template <typename T, int DL, int DV, bool kExtend>
__global__ void __launch_bounds__(kMlaThreads)
attn_mla_kernel(const MlaParams p) {
  constexpr int LD = DL + 8;
  constexpr int KSTEPS = DL / 16;
  constexpr int VEC_PER_ROW = DL / 8;
  extern __shared__ __align__(16) uint8_t mla_smem[];
  T* sQ = reinterpret_cast<T*>(mla_smem);                       // [16][LD]
  T* sK = sQ + kMlaHeads * LD;                                  // [2][32][LD]
  float* sS = reinterpret_cast<float*>(sK + 2 * kMlaChunk * LD);  // [8][16][32] partial scores
  T* sP = reinterpret_cast<T*>(sS + 8 * kMlaHeads * kMlaChunk);   // [16][40]
  float* s_m = reinterpret_cast<float*>(sP + kMlaHeads * 40);     // [16]
  float* s_l = s_m + 16;
  float* s_alpha = s_l + 16;

  griddep_wait();
  const int item = blockIdx.z, hg = blockIdx.y, split = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int h0 = hg * kMlaHeads;
  const int nh = min(kMlaHeads, p.H - h0);
  int req, L;
  if constexpr (kExtend) {
    int lo = 0, hi = p.batch - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (p.extend_start_loc[mid] <= item) lo = mid; else hi = mid - 1;
    }
    req = lo;
    L = p.prefix_lens[req] + (item - p.extend_start_loc[req]) + 1;
  } else {
    req = item;
    L = static_cast<int>(p.seq_lens[req]);
  }
  const int32_t* rtt = p.req_to_token + p.req_pool_indices[req] * p.rtt_stride;
  int per = (L + p.splits - 1) / p.splits;
  per = (per + kMlaChunk - 1) / kMlaChunk * kMlaChunk;
  const int t0 = split * per;
  const int t1 = min(L, t0 + per);
  const int n_chunks = t1 > t0 ? (t1 - t0 + kMlaChunk - 1) / kMlaChunk : 0;

  // Q rows of this head group (rows >= nh are zero)
  {
    const T* qb = reinterpret_cast<const T*>(p.q) + item * p.q_stride + static_cast<int64_t>(h0) * DL;
    for (int v = tid; v < kMlaHeads * VEC_PER_ROW; v += kMlaThreads) {
      const int r = v / VEC_PER_ROW, c = v - r * VEC_PER_ROW;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (r < nh) val = *reinterpret_cast<const uint4*>(qb + r * DL + c * 8);
      *reinterpret_cast<uint4*>(sQ + r * LD + c * 8) = val;
    }
    if (tid < 16) { s_m[tid] = -INFINITY; s_l[tid] = 0.f; }
  }
  const T* pool = reinterpret_cast<const T*>(p.kv_pool);
  auto load_chunk = [&](int ci) {
    T* dst = sK + (ci & 1) * kMlaChunk * LD;
    const int base = t0 + ci * kMlaChunk;
    for (int v = tid; v < kMlaChunk * VEC_PER_ROW; v += kMlaThreads) {
      const int r = v / VEC_PER_ROW, c = v - r * VEC_PER_ROW;
      const bool valid = base + r < t1;
      const int64_t slot = valid ? rtt[base + r] : 0;
      cp_async16(smem_u32(dst + r * LD + c * 8), pool + slot * DL + c * 8, valid);
    }
    cp_async_commit();
  };

  float o_acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o_acc[i][j] = 0.f;

  if (n_chunks > 0) load_chunk(0);
  for (int ci = 0; ci < n_chunks; ++ci) {
    if (ci + 1 < n_chunks) {
      load_chunk(ci + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const T* kb = sK + (ci & 1) * kMlaChunk * LD;
    const int n_valid = min(kMlaChunk, t1 - (t0 + ci * kMlaChunk));
    // ---- S partials: this warp's k-steps ----
    float s_acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s_acc[i][j] = 0.f;
    for (int ks = warp; ks < KSTEPS; ks += 8) {
      const int k0 = ks * 16;
      uint32_t a[4];
      a[0] = *reinterpret_cast<const uint32_t*>(sQ + g * LD + k0 + 2 * tq);
      a[1] = *reinterpret_cast<const uint32_t*>(sQ + (g + 8) * LD + k0 + 2 * tq);
      a[2] = *reinterpret_cast<const uint32_t*>(sQ + g * LD + k0 + 8 + 2 * tq);
      a[3] = *reinterpret_cast<const uint32_t*>(sQ + (g + 8) * LD + k0 + 8 + 2 * tq);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const T* kr = kb + (nt * 8 + g) * LD + k0 + 2 * tq;
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kr);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kr + 8);
        mma_16816<T>(s_acc[nt], a, b0, b1);
      }
    }
    {
      float* my = sS + warp * (kMlaHeads * kMlaChunk);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = nt * 8 + 2 * tq;
        my[g * kMlaChunk + col] = s_acc[nt][0];
        my[g * kMlaChunk + col + 1] = s_acc[nt][1];
        my[(g + 8) * kMlaChunk + col] = s_acc[nt][2];
        my[(g + 8) * kMlaChunk + col + 1] = s_acc[nt][3];
      }
    }
    __syncthreads();
    // ---- reduce the 8 partials + online softmax: warp w owns heads w and w + 8, lane = token ----
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int h = warp + hh * 8;
      float s = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) s += sS[w8 * (kMlaHeads * kMlaChunk) + h * kMlaChunk + lane];
      s = lane < n_valid ? s * p.scale_log2 : -INFINITY;
      const float m_old = s_m[h];
      const float m_new = fmaxf(m_old, warp_max(s));
      const float pj = exp2f(s - m_new);
      const T pt = from_f32<T>(pj);
      const float ps = warp_sum(to_f32<T>(pt));  // sum what the tensor core multiplies
      const float alpha = (m_old == -INFINITY) ? 0.f : exp2f(m_old - m_new);
      sP[h * 40 + lane] = pt;
      __syncwarp();
      if (lane == 0) { s_m[h] = m_new; s_l[h] = s_l[h] * alpha + ps; s_alpha[h] = alpha; }
    }
    __syncthreads();
    // ---- O += P V : this warp's 64 output columns ----
    {
      const float al0 = s_alpha[g], al1 = s_alpha[g + 8];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        o_acc[nt][0] *= al0; o_acc[nt][1] *= al0; o_acc[nt][2] *= al1; o_acc[nt][3] *= al1;
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        uint32_t a[4];
        a[0] = *reinterpret_cast<const uint32_t*>(sP + g * 40 + kk * 16 + 2 * tq);
        a[1] = *reinterpret_cast<const uint32_t*>(sP + (g + 8) * 40 + kk * 16 + 2 * tq);
        a[2] = *reinterpret_cast<const uint32_t*>(sP + g * 40 + kk * 16 + 8 + 2 * tq);
        a[3] = *reinterpret_cast<const uint32_t*>(sP + (g + 8) * 40 + kk * 16 + 8 + 2 * tq);
        const unsigned short* vb = reinterpret_cast<const unsigned short*>(kb);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int d = warp * 64 + nt * 8 + g;
          const int r0 = kk * 16 + 2 * tq;
          const uint32_t b0 = static_cast<uint32_t>(vb[r0 * LD + d]) |
                              (static_cast<uint32_t>(vb[(r0 + 1) * LD + d]) << 16);
          const uint32_t b1 = static_cast<uint32_t>(vb[(r0 + 8) * LD + d]) |
                              (static_cast<uint32_t>(vb[(r0 + 9) * LD + d]) << 16);
          mma_16816<T>(o_acc[nt], a, b0, b1);
        }
      }
    }
    __syncthreads();  // sS / sP / this K buffer are rewritten by the next iterations
  }

  // ---- epilogue ----
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int hr = g + half * 8;
    if (hr < nh) {
      const int h = h0 + hr;
      const float l = s_l[hr];
      if (p.splits == 1) {
        T* dst = reinterpret_cast<T*>(p.out) + item * p.out_stride + static_cast<int64_t>(h) * DV;
        const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int d = warp * 64 + nt * 8 + 2 * tq;
          *reinterpret_cast<uint32_t*>(dst + d) = pack2<T>(o_acc[nt][half * 2] * inv, o_acc[nt][half * 2 + 1] * inv);
        }
      } else {
        const int64_t idx = (static_cast<int64_t>(item) * p.H + h) * p.splits + split;
        float* dst = p.o_part + idx * DV;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int d = warp * 64 + nt * 8 + 2 * tq;
          dst[d] = o_acc[nt][half * 2];
          dst[d + 1] = o_acc[nt][half * 2 + 1];
        }
        if (warp == 0 && tq == 0) { p.ml_part[idx * 2] = s_m[hr]; p.ml_part[idx * 2 + 1] = l; }
      }
    }
  }
  griddep_launch_dependents();
}

template <typename T, int DV>
__global__ void __launch_bounds__(256)
mla_merge_kernel(T* __restrict__ out, const float* __restrict__ o_part, const float* __restrict__ ml_part, int H,
                 int splits, int64_t out_stride) {
  griddep_wait();
  const int item = blockIdx.y, h = blockIdx.x;
  const int64_t base = (static_cast<int64_t>(item) * H + h) * splits;
  float M = -INFINITY;
  for (int s = 0; s < splits; ++s)
    if (ml_part[(base + s) * 2 + 1] > 0.f) M = fmaxf(M, ml_part[(base + s) * 2]);
  for (int t = threadIdx.x; t < DV; t += 256) {
    float num = 0.f, den = 0.f;
    for (int s = 0; s < splits; ++s) {
      const float m = ml_part[(base + s) * 2], l = ml_part[(base + s) * 2 + 1];
      if (l > 0.f) {
        const float w = exp2f(m - M);
        num += w * o_part[(base + s) * DV + t];
        den += w * l;
      }
    }
    out[item * out_stride + static_cast<int64_t>(h) * DV + t] = from_f32<T>(den > 0.f ? num / den : 0.f);
  }
}

template <int DL>
constexpr int mla_smem_bytes() {
  return (kMlaHeads + 2 * kMlaChunk) * (DL + 8) * 2 + 8 * kMlaHeads * kMlaChunk * 4 + kMlaHeads * 40 * 2 + 48 * 4;
}

static int mla_splits(int64_t items, int head_groups, int sms) {
  const int64_t ctas = items * head_groups;
  int s = static_cast<int>((2ll * sms + ctas - 1) / ctas);
  if (s < 1) s = 1;
  if (s > kMlaMaxSplits) s = kMlaMaxSplits;
  return s;
}

}  // namespace spd

using namespace spd;

extern "C" {

int64_t spd_attn_mla_workspace_bytes(int num_heads, int kv_lora_rank) {
  // splits are only used while items x head groups < 2 x SMs: at most 2 x SMs items
  const int64_t items = 2ll * device_sm_count_cached();
  return items * num_heads * kMlaMaxSplits * (kv_lora_rank + 2) * 4 + 256;
}

int spd_attn_mla(void* out, const void* q, const void* kv_pool, const int32_t* req_to_token,
                 const int64_t* req_pool_indices, const int64_t* seq_lens, const int32_t* prefix_lens,
                 const int32_t* extend_start_loc, int64_t batch, int64_t total_tokens, int num_heads,
                 int kv_lora_rank, int rope_dim, int64_t q_stride, int64_t out_stride,
                 int64_t req_to_token_stride, float sm_scale, int dtype, void* workspace,
                 int64_t workspace_bytes, int sm_budget, spd_stream_t stream) {
  SPD_REQUIRE(out && q && kv_pool && req_to_token && req_pool_indices, SPD_E_BADARG, "spd_attn_mla: null pointer");
  const bool extend = seq_lens == nullptr;
  SPD_REQUIRE(!extend || (prefix_lens && extend_start_loc), SPD_E_BADARG,
              "spd_attn_mla: give seq_lens (decode) or prefix_lens + extend_start_loc (extend)");
  SPD_REQUIRE(kv_lora_rank == 512 && rope_dim == 64, SPD_E_SHAPE,
              "spd_attn_mla: built for kv_lora_rank 512 + rope 64 (got %d + %d)", kv_lora_rank, rope_dim);
  SPD_REQUIRE(num_heads > 0 && batch >= 0 && total_tokens >= 0, SPD_E_BADARG, "spd_attn_mla: bad sizes");
  SPD_REQUIRE(q_stride % 8 == 0 && out_stride % 8 == 0 && aligned16(out) && aligned16(q) && aligned16(kv_pool),
              SPD_E_ALIGN, "spd_attn_mla: strides must be multiples of 8 and pointers 16-B aligned");
  const int64_t items = extend ? total_tokens : batch;
  if (items == 0) return 0;
  SPD_REQUIRE(items < 65536, SPD_E_SHAPE, "spd_attn_mla: more than 65535 query tokens in one call");
  constexpr int DL = 576, DV = 512;
  const int sms = sm_budget > 0 ? sm_budget : device_sm_count_cached();
  const int hgs = (num_heads + kMlaHeads - 1) / kMlaHeads;
  MlaParams p{};
  p.out = out;
  p.q = q;
  p.kv_pool = kv_pool;
  p.req_to_token = req_to_token;
  p.req_pool_indices = req_pool_indices;
  p.seq_lens = seq_lens;
  p.prefix_lens = prefix_lens;
  p.extend_start_loc = extend_start_loc;
  p.batch = static_cast<int>(batch);
  p.H = num_heads;
  p.q_stride = q_stride;
  p.out_stride = out_stride;
  p.rtt_stride = req_to_token_stride;
  p.scale_log2 = sm_scale * 1.4426950408889634f;
  p.splits = mla_splits(items, hgs, sms);
  if (p.splits > 1) {
    const int64_t need = items * num_heads * p.splits * (DV + 2) * 4;
    SPD_REQUIRE(workspace && workspace_bytes >= need, SPD_E_WORKSPACE,
                "spd_attn_mla: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need);
    p.o_part = reinterpret_cast<float*>(workspace);
    p.ml_part = p.o_part + items * num_heads * p.splits * DV;
  }
  const dim3 grid(p.splits, hgs, static_cast<unsigned>(items));
  constexpr int smem = mla_smem_bytes<DL>();
  cudaStream_t st = as_stream(stream);
  SPD_DISPATCH_DTYPE(dtype, {
    if (extend) {
      auto kern = attn_mla_kernel<T, DL, DV, true>;
      static bool configured = false;
      if (!configured) {
        SPD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
      }
      SPD_CUDA(launch(kern, grid, dim3(kMlaThreads), smem, st, pdl_enabled(), p));
    } else {
      auto kern = attn_mla_kernel<T, DL, DV, false>;
      static bool configured = false;
      if (!configured) {
        SPD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
      }
      SPD_CUDA(launch(kern, grid, dim3(kMlaThreads), smem, st, pdl_enabled(), p));
    }
    if (p.splits > 1) {
      SPD_CUDA(launch(mla_merge_kernel<T, DV>, dim3(num_heads, static_cast<unsigned>(items)), dim3(256), 0, st,
                      pdl_enabled(), reinterpret_cast<T*>(out), (const float*)p.o_part, (const float*)p.ml_part,
                      num_heads, p.splits, out_stride));
    }
  });
  return 0;
}

}  // extern "C"
