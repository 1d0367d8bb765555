// Plain SIMT attention over the token-granular paged KV pool — bring-up implementation and
// on-device cross-check for the tcgen05 kernels (attn_prefill.cu / attn_decode.cu).  The product
// path never calls these.  Semantics: S/layers/attention/torch_native_backend.py:26-179.
//
// One CTA handles one (query item, kv head) pair — all G = Hq/Hk query heads of the GQA group —
// and one KV split.  A query item is a decode request (keys 0..seq_len) or an extend token
// (keys 0..prefix+j).  Tokens are visited in chunks of 64: scores (16 lanes per token, 16-byte
// K loads), online softmax per head, then PV with each thread owning 8 output dims.
#include "common.cuh"
#include "ptx.cuh"

namespace spd {

constexpr int kSimtThreads = 128;
constexpr int kSimtChunk = 64;
constexpr int kMaxGroup = 8;
constexpr int kMaxSplits = 32;

struct AttnSimtParams {
  void* out;
  const void* q;
  const void* k_pool;
  const void* v_pool;
  const int32_t* req_to_token;
  const int64_t* req_pool_indices;
  const int64_t* seq_lens;          // decode
  const int32_t* prefix_lens;       // extend
  const int32_t* extend_lens;       // extend
  const int32_t* extend_start_loc;  // extend
  int batch;
  int Hq, Hk;
  int64_t q_stride, out_stride, rtt_stride;
  float scale_log2;  // sm_scale * log2(e)
  int splits;
  float* o_part;  // [item][Hq][splits][D]
  float* ml_part; // [item][Hq][splits][2]
};

template <typename T, int D, bool kExtend>
__global__ void __launch_bounds__(kSimtThreads)
attn_simt_kernel(const AttnSimtParams p) {
  constexpr int LPT = D / 8;                 // lanes per token
  constexpr int TPI = kSimtThreads / LPT;    // tokens per iteration
  __shared__ int slots[kSimtChunk];
  __shared__ float s_sc[kMaxGroup][kSimtChunk];
  __shared__ float s_m[kMaxGroup], s_l[kMaxGroup], s_alpha[kMaxGroup];
  __shared__ float s_red[TPI][D];

  griddep_wait();
  const int item = blockIdx.z, hk = blockIdx.y, split = blockIdx.x;
  const int G = p.Hq / p.Hk;
  const int t = threadIdx.x;
  int req, L;
  if constexpr (kExtend) {
    // request of this token: last r with extend_start_loc[r] <= item
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
  // this split's token range (multiples of the chunk so splits never share a chunk)
  int per = (L + p.splits - 1) / p.splits;
  per = (per + kSimtChunk - 1) / kSimtChunk * kSimtChunk;
  const int t0 = split * per;
  const int t1 = min(L, t0 + per);

  const int c = t % LPT;   // 8-dim slice
  const int tg = t / LPT;  // token group
  const T* qbase = reinterpret_cast<const T*>(p.q) + item * p.q_stride + static_cast<int64_t>(hk) * G * D;
  float qf[kMaxGroup][8];
#pragma unroll
  for (int g = 0; g < kMaxGroup; ++g) {
    if (g < G) {
      const uint4 v = *reinterpret_cast<const uint4*>(qbase + g * D + c * 8);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack2<T>(w[i]);
        qf[g][2 * i] = f.x * p.scale_log2;
        qf[g][2 * i + 1] = f.y * p.scale_log2;
      }
    }
  }
  float acc[kMaxGroup][8];
#pragma unroll
  for (int g = 0; g < kMaxGroup; ++g)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[g][i] = 0.f;
  if (t < kMaxGroup) { s_m[t] = -INFINITY; s_l[t] = 0.f; }
  __syncthreads();

  const T* kp = reinterpret_cast<const T*>(p.k_pool);
  const T* vp = reinterpret_cast<const T*>(p.v_pool);
  const int64_t row = static_cast<int64_t>(p.Hk) * D;

  for (int base = t0; base < t1; base += kSimtChunk) {
    const int n = min(kSimtChunk, t1 - base);
    if (t < n) slots[t] = rtt[base + t];
    __syncthreads();
    // scores
    for (int j = tg; j < kSimtChunk; j += TPI) {
      float part[kMaxGroup];
#pragma unroll
      for (int g = 0; g < kMaxGroup; ++g) part[g] = 0.f;
      if (j < n) {
        const uint4 v = *reinterpret_cast<const uint4*>(kp + slots[j] * row + hk * D + c * 8);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        float kf[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = unpack2<T>(w[i]);
          kf[2 * i] = f.x;
          kf[2 * i + 1] = f.y;
        }
#pragma unroll
        for (int g = 0; g < kMaxGroup; ++g)
          if (g < G) {
#pragma unroll
            for (int i = 0; i < 8; ++i) part[g] += qf[g][i] * kf[i];
          }
      }
#pragma unroll
      for (int g = 0; g < kMaxGroup; ++g)
        if (g < G) {
#pragma unroll
          for (int o = LPT / 2; o > 0; o >>= 1) part[g] += __shfl_xor_sync(0xffffffffu, part[g], o);
          if (c == 0) s_sc[g][j] = (j < n) ? part[g] : -INFINITY;
        }
    }
    __syncthreads();
    // online softmax: warp w owns heads w, w+4
    {
      const int warp = t >> 5, lane = t & 31;
      for (int g = warp; g < G; g += kSimtThreads / 32) {
        const float a = s_sc[g][lane], b = s_sc[g][lane + 32];
        const float cmax = warp_max(fmaxf(a, b));
        const float m_old = s_m[g];
        const float m_new = fmaxf(m_old, cmax);
        const float pa = exp2f(a - m_new), pb = exp2f(b - m_new);
        const float ps = warp_sum(pa + pb);
        const float alpha = (m_old == -INFINITY) ? 0.f : exp2f(m_old - m_new);
        __syncwarp();
        s_sc[g][lane] = pa;
        s_sc[g][lane + 32] = pb;
        if (lane == 0) { s_m[g] = m_new; s_l[g] = s_l[g] * alpha + ps; s_alpha[g] = alpha; }
      }
    }
    __syncthreads();
    // PV
#pragma unroll
    for (int g = 0; g < kMaxGroup; ++g)
      if (g < G) {
        const float al = s_alpha[g];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[g][i] *= al;
      }
    for (int j = tg; j < n; j += TPI) {
      const uint4 v = *reinterpret_cast<const uint4*>(vp + slots[j] * row + hk * D + c * 8);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      float vf[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack2<T>(w[i]);
        vf[2 * i] = f.x;
        vf[2 * i + 1] = f.y;
      }
#pragma unroll
      for (int g = 0; g < kMaxGroup; ++g)
        if (g < G) {
          const float pj = s_sc[g][j];
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[g][i] += pj * vf[i];
        }
    }
    __syncthreads();
  }

  // reduce the TPI token groups, one head at a time (unrolled so acc[] stays in registers)
#pragma unroll
  for (int g = 0; g < kMaxGroup; ++g) {
    if (g < G) {  // G is block-uniform, so the barriers below are reached by every thread
#pragma unroll
      for (int i = 0; i < 8; ++i) s_red[tg][c * 8 + i] = acc[g][i];
      __syncthreads();
      if (t < D) {
        float o = 0.f;
#pragma unroll
        for (int r = 0; r < TPI; ++r) o += s_red[r][t];
        const int h = hk * G + g;
        if (p.splits == 1) {
          const float l = s_l[g];
          reinterpret_cast<T*>(p.out)[item * p.out_stride + static_cast<int64_t>(h) * D + t] =
              from_f32<T>(l > 0.f ? o / l : 0.f);
        } else {
          const int64_t idx = (static_cast<int64_t>(item) * p.Hq + h) * p.splits + split;
          p.o_part[idx * D + t] = o;
          if (t == 0) { p.ml_part[idx * 2] = s_m[g]; p.ml_part[idx * 2 + 1] = s_l[g]; }
        }
      }
      __syncthreads();
    }
  }
  griddep_launch_dependents();
}

template <typename T, int D>
__global__ void __launch_bounds__(D)
attn_merge_kernel(T* __restrict__ out, const float* __restrict__ o_part,
                  const float* __restrict__ ml_part, int Hq, int splits, int64_t out_stride) {
  griddep_wait();
  const int item = blockIdx.y, h = blockIdx.x, t = threadIdx.x;
  const int64_t base = (static_cast<int64_t>(item) * Hq + h) * splits;
  float M = -INFINITY;
  for (int s = 0; s < splits; ++s) M = fmaxf(M, ml_part[(base + s) * 2]);
  float num = 0.f, den = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float m = ml_part[(base + s) * 2], l = ml_part[(base + s) * 2 + 1];
    if (l > 0.f) {
      const float w = exp2f(m - M);
      num += w * o_part[(base + s) * D + t];
      den += w * l;
    }
  }
  out[item * out_stride + static_cast<int64_t>(h) * D + t] = from_f32<T>(den > 0.f ? num / den : 0.f);
  griddep_launch_dependents();
}

int choose_kv_splits(int64_t batch, int num_kv_heads, int sms) {
  int64_t blocks = batch * num_kv_heads;
  int splits = static_cast<int>((2ll * sms + blocks - 1) / blocks);
  if (splits < 1) splits = 1;
  if (splits > kMaxSplits) splits = kMaxSplits;
  return splits;
}

template <typename T, int D>
static int launch_simt(bool extend, const AttnSimtParams& p, int64_t items, cudaStream_t st) {
  dim3 grid(p.splits, p.Hk, static_cast<unsigned>(items));
  if (extend)
    SPD_CUDA(launch(attn_simt_kernel<T, D, true>, grid, dim3(kSimtThreads), 0, st, pdl_enabled(), p));
  else
    SPD_CUDA(launch(attn_simt_kernel<T, D, false>, grid, dim3(kSimtThreads), 0, st, pdl_enabled(), p));
  if (p.splits > 1) {
    SPD_CUDA(launch(attn_merge_kernel<T, D>, dim3(p.Hq, static_cast<unsigned>(items)), dim3(D), 0, st,
                    pdl_enabled(), reinterpret_cast<T*>(p.out), (const float*)p.o_part,
                    (const float*)p.ml_part, p.Hq, p.splits, p.out_stride));
  }
  return 0;
}

int check_attn_args(const char* fn, const void* out, const void* q, const void* k_pool,
                    const void* v_pool, const void* rtt, const void* rpi, int Hq, int Hk, int D,
                    int64_t q_stride, int64_t out_stride) {
  SPD_REQUIRE(out && q && k_pool && v_pool && rtt && rpi, SPD_E_BADARG, "%s: null pointer", fn);
  SPD_REQUIRE(Hq > 0 && Hk > 0 && Hq % Hk == 0 && Hq / Hk <= kMaxGroup, SPD_E_SHAPE,
              "%s: need Hq %% Hk == 0 and group <= %d (Hq=%d Hk=%d)", fn, kMaxGroup, Hq, Hk);
  SPD_REQUIRE(D == 64 || D == 128, SPD_E_SHAPE, "%s: head_dim must be 64 or 128 (got %d)", fn, D);
  SPD_REQUIRE(q_stride % 8 == 0 && out_stride % 8 == 0 && aligned16(out) && aligned16(q) &&
                  aligned16(k_pool) && aligned16(v_pool),
              SPD_E_ALIGN, "%s: strides must be multiples of 8 and pointers 16-B aligned", fn);
  return 0;
}

}  // namespace spd

using namespace spd;

extern "C" {

int64_t spd_attn_decode_workspace_bytes(int64_t max_batch, int num_q_heads, int head_dim) {
  // [64 KB of arrival counters, one per (request, kv head) — at a FIXED offset so that they stay
  // zero / re-armed whatever shapes the workspace is reused with] [split partials o | m, l]
  return 65536 + max_batch * num_q_heads * kMaxSplits * (head_dim + 2) * 4 + 256;
}

int spd_attn_decode_simt(void* out, const void* q, const void* k_pool, const void* v_pool,
                         const int32_t* req_to_token, const int64_t* req_pool_indices,
                         const int64_t* seq_lens, int64_t batch, int num_q_heads, int num_kv_heads,
                         int head_dim, int64_t q_stride, int64_t out_stride,
                         int64_t req_to_token_stride, float sm_scale, int dtype, void* workspace,
                         int64_t workspace_bytes, spd_stream_t stream) {
  int rc = check_attn_args("spd_attn_decode_simt", out, q, k_pool, v_pool, req_to_token,
                           req_pool_indices, num_q_heads, num_kv_heads, head_dim, q_stride, out_stride);
  if (rc) return rc;
  SPD_REQUIRE(seq_lens && batch >= 0, SPD_E_BADARG, "spd_attn_decode_simt: bad seq_lens/batch");
  if (batch == 0) return 0;
  AttnSimtParams p{};
  p.out = out; p.q = q; p.k_pool = k_pool; p.v_pool = v_pool;
  p.req_to_token = req_to_token; p.req_pool_indices = req_pool_indices; p.seq_lens = seq_lens;
  p.batch = static_cast<int>(batch); p.Hq = num_q_heads; p.Hk = num_kv_heads;
  p.q_stride = q_stride; p.out_stride = out_stride; p.rtt_stride = req_to_token_stride;
  p.scale_log2 = sm_scale * 1.4426950408889634f;
  p.splits = choose_kv_splits(batch, num_kv_heads, device_sm_count_cached());
  if (p.splits > 1) {
    SPD_REQUIRE(workspace && workspace_bytes >= spd_attn_decode_workspace_bytes(batch, num_q_heads, head_dim),
                SPD_E_WORKSPACE, "spd_attn_decode_simt: workspace too small");
    p.o_part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + 65536);
    p.ml_part = p.o_part + batch * num_q_heads * kMaxSplits * head_dim;
  }
  SPD_DISPATCH_DTYPE(dtype, {
    if (head_dim == 128) return launch_simt<T, 128>(false, p, batch, as_stream(stream));
    return launch_simt<T, 64>(false, p, batch, as_stream(stream));
  });
  return 0;
}

int spd_attn_prefill_simt(void* out, const void* q, const void* k_pool, const void* v_pool,
                          const int32_t* req_to_token, const int64_t* req_pool_indices,
                          const int32_t* prefix_lens, const int32_t* extend_lens,
                          const int32_t* extend_start_loc, int64_t batch, int64_t total_tokens,
                          int64_t max_extend_len, int num_q_heads, int num_kv_heads, int head_dim,
                          int64_t q_stride, int64_t out_stride, int64_t req_to_token_stride,
                          float sm_scale, int dtype, spd_stream_t stream) {
  (void)max_extend_len;
  int rc = check_attn_args("spd_attn_prefill_simt", out, q, k_pool, v_pool, req_to_token,
                           req_pool_indices, num_q_heads, num_kv_heads, head_dim, q_stride, out_stride);
  if (rc) return rc;
  SPD_REQUIRE(prefix_lens && extend_lens && extend_start_loc && batch >= 0 && total_tokens >= 0,
              SPD_E_BADARG, "spd_attn_prefill_simt: bad arguments");
  if (batch == 0 || total_tokens == 0) return 0;
  SPD_REQUIRE(total_tokens <= 65535, SPD_E_SHAPE, "spd_attn_prefill_simt: at most 65535 tokens");
  AttnSimtParams p{};
  p.out = out; p.q = q; p.k_pool = k_pool; p.v_pool = v_pool;
  p.req_to_token = req_to_token; p.req_pool_indices = req_pool_indices;
  p.prefix_lens = prefix_lens; p.extend_lens = extend_lens; p.extend_start_loc = extend_start_loc;
  p.batch = static_cast<int>(batch); p.Hq = num_q_heads; p.Hk = num_kv_heads;
  p.q_stride = q_stride; p.out_stride = out_stride; p.rtt_stride = req_to_token_stride;
  p.scale_log2 = sm_scale * 1.4426950408889634f;
  p.splits = 1;
  SPD_DISPATCH_DTYPE(dtype, {
    if (head_dim == 128) return launch_simt<T, 128>(true, p, total_tokens, as_stream(stream));
    return launch_simt<T, 64>(true, p, total_tokens, as_stream(stream));
  });
  return 0;
}

}  // extern "C"
