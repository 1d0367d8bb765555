// Ragged chunked-prefill ("extend") attention on tcgen05: QK^T and PV both run on the tensor
// cores with S and O accumulators in TMEM.  One pass over [prefix pages || new tokens] with the
// causal mask applied only on the diagonal tiles — no separate ragged / paged kernels and no LSE
// merge as in the flashinfer backend (S/layers/attention/flashinfer_backend.py:425-445).
// Semantics: S/layers/attention/torch_native_backend.py:26-109.
//
// CTA = 128 query rows of one request x one query head (grid: q-tile (heaviest first), head,
// request; CTAs past a request's length exit).  KV is visited in 128-token tiles:
//   loaders  (warps 4-7): gather K/V rows through req_to_token with 16-byte cp.async into
//                         128-byte-swizzled smem (page size is 1 token, so rows are scattered);
//   MMA      (warp 8)   : S_j = Q K_j^T (K-major operands), O += P_j V_j (V as MN-major operand);
//                         S is double-buffered in TMEM so S_{j+1} is computed while the softmax
//                         of tile j runs;
//   softmax  (warps 0-3): one thread per query row — TMEM -> registers, mask, online softmax in
//                         the exp2 domain with LAZY rescaling (O in TMEM is only rescaled when a
//                         row max grows by > 2^8), P -> smem as the A operand of the PV MMA,
//                         final O / l -> global.
// Q arrives by TMA.  TMEM: S0 | S1 | O = 384 of 512 columns.
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace spd {

constexpr int kPfThreads = 288;
constexpr int kPfM = 128;    // query rows per CTA
constexpr int kPfN = 128;    // kv tokens per tile
constexpr int kPfD = 128;    // head dim
constexpr int kTileBytes = 128 * 128 * 2;  // one [128 x 128] 16-bit tile = two 16 KB swizzle blocks
constexpr int kPfSmem = kTileBytes * 6 + 1024 + 256;  // Q, K0, K1, V0, V1, P + align + barriers
constexpr float kRescaleThreshold = 8.0f;  // log2 units

struct PrefillParams {
  void* out;
  const void* k_pool;
  const void* v_pool;
  const int32_t* req_to_token;
  const int64_t* req_pool_indices;
  const int32_t* prefix_lens;
  const int32_t* extend_lens;
  const int32_t* extend_start_loc;
  int Hq, Hk;
  int64_t out_stride, rtt_stride;
  float scale_log2;
  int q_tiles;
};

// byte offset of (row, 16-byte chunk) inside a [128 x 128] 16-bit tile stored as two
// [128 rows x 128 B] blocks with the 128-byte swizzle (chunk index XOR row%8)
__device__ __forceinline__ uint32_t tile_off(int row, int chunk16) {
  return (chunk16 >> 3) * 16384 + row * 128 + (((chunk16 & 7) ^ (row & 7)) << 4);
}

template <typename T>
__global__ void __launch_bounds__(kPfThreads, 1)
attn_prefill_kernel(const __grid_constant__ CUtensorMap tmQ, const PrefillParams p) {
  const int req = blockIdx.z, head = blockIdx.y;
  const int qt = p.q_tiles - 1 - blockIdx.x;  // heaviest (latest) tiles first
  const int m0 = qt * kPfM;
  const int ext = p.extend_lens[req];
  if (m0 >= ext) return;
  const int pre = p.prefix_lens[req];
  const int q_row0 = p.extend_start_loc[req] + m0;
  const int rows_valid = min(kPfM, ext - m0);
  const int kv_len = pre + m0 + rows_valid;  // keys visible to the last valid row of this tile
  const int n_tiles = (kv_len + kPfN - 1) / kPfN;
  const int hk = head / (p.Hq / p.Hk);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTileBytes;      // 2 stages
  uint8_t* sV = smem + 3 * kTileBytes;  // 2 stages
  uint8_t* sP = smem + 5 * kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * kTileBytes);
  uint64_t* q_full = bars;          // 1
  uint64_t* kv_full = bars + 1;     // 2
  uint64_t* kv_empty = bars + 3;    // 2
  uint64_t* s_full = bars + 5;      // 2
  uint64_t* s_empty = bars + 7;     // 2
  uint64_t* p_full = bars + 9;      // 1
  uint64_t* o_done = bars + 10;     // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      mbar_init(q_full, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&kv_full[i], 128);
        mbar_init(&kv_empty[i], 1);
        mbar_init(&s_full[i], 1);
        mbar_init(&s_empty[i], 4);
      }
      mbar_init(p_full, 4);
      mbar_init(o_done, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;        // 2 x 128 columns
  const uint32_t tmem_O = tmem_base + 256;  // 128 columns

  griddep_wait();

  if (warp == 8) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      mbar_expect_tx(q_full, kTileBytes);
      tma_load_2d(sQ, &tmQ, q_full, head * kPfD, q_row0, kEvictFirst);
      tma_load_2d(sQ + 16384, &tmQ, q_full, head * kPfD + 64, q_row0, kEvictFirst);
      constexpr uint32_t idesc_qk = umma_idesc<T>(kPfM, kPfN, 0, 0);  // A, B K-major
      constexpr uint32_t idesc_pv = umma_idesc<T>(kPfM, kPfD, 0, 1);  // B (V) MN-major
      const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
      auto issue_qk = [&](int j) {
        const int st = j & 1;
        mbar_wait(&kv_full[st], (j >> 1) & 1);
        mbar_wait(&s_empty[st], ((j >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t k_addr = smem_u32(sK + st * kTileBytes);
#pragma unroll
        for (int k = 0; k < kPfD / 16; ++k) {
          const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
          umma_ss(tmem_S + st * kPfN, umma_desc_sw128(q_addr + off, 16, 1024),
                  umma_desc_sw128(k_addr + off, 16, 1024), idesc_qk, k > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[st]);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) issue_qk(j + 1);
        mbar_wait(p_full, j & 1);
        tc_fence_after_sync();
        const uint32_t v_addr = smem_u32(sV + (j & 1) * kTileBytes);
#pragma unroll
        for (int k = 0; k < kPfN / 16; ++k) {
          // A = P[128 x 16 kv] (K-major), B = V[16 kv x 128 d] (MN-major: 2 blocks of 64 d,
          // LBO = 16 KB; kv rows 128 B apart, 8-row groups SBO = 1 KB)
          const uint32_t a_off = (k >> 2) * 16384 + (k & 3) * 32;
          umma_ss(tmem_O, umma_desc_sw128(p_addr + a_off, 16, 1024),
                  umma_desc_sw128(v_addr + k * 2048, 16384, 1024), idesc_pv,
                  (j > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&kv_empty[j & 1]);
        umma_commit(o_done);
      }
    }
  } else if (warp >= 4) {
    // ===================== K/V gather loaders (128 threads) =====================
    const int lt = threadIdx.x - 128;
    const int chunk = lt & 15;   // 16-byte chunk of the 256-byte head row
    const int rsub = lt >> 4;    // 0..7
    const T* kp = reinterpret_cast<const T*>(p.k_pool);
    const T* vp = reinterpret_cast<const T*>(p.v_pool);
    const int32_t* rtt = p.req_to_token + p.req_pool_indices[req] * p.rtt_stride;
    const int64_t row_elems = static_cast<int64_t>(p.Hk) * kPfD;
    const int64_t col = static_cast<int64_t>(hk) * kPfD + chunk * 8;
    // Double buffering: tile j is gathered while tile j-1 is being consumed.  Tile j-1 is
    // published BEFORE waiting for tile j's stage, otherwise the MMA warp (which needs K_{j-1}
    // before it can retire PV_{j-2} and free that stage) and the loaders would wait on each other.
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j & 1;
      // page-table lookups first, all in flight together and overlapped with the waits below (the
      // cp.async asm is a compiler barrier; interleaving the lookups with it would serialise 16
      // dependent global-load latencies per tile)
      int slot[16];
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int kv = j * kPfN + it * 8 + rsub;
        slot[it] = kv < kv_len ? __ldg(rtt + kv) : -1;
      }
      if (j > 0) {
        cp_async_wait<0>();
        fence_proxy_async_smem();
        mbar_arrive(&kv_full[(j - 1) & 1]);
      }
      mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
      const uint32_t kdst = smem_u32(sK + st * kTileBytes);
      const uint32_t vdst = smem_u32(sV + st * kTileBytes);
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int row = it * 8 + rsub;
        const bool valid = slot[it] >= 0;
        const int64_t s64 = valid ? slot[it] : 0;
        const uint32_t off = tile_off(row, chunk);
        cp_async16(kdst + off, kp + s64 * row_elems + col, valid);
        cp_async16(vdst + off, vp + s64 * row_elems + col, valid);
      }
      cp_async_commit();
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    mbar_arrive(&kv_full[(n_tiles - 1) & 1]);
  } else {
    // ===================== softmax / correction / epilogue (one thread per query row) ==========
    const int row = threadIdx.x;  // == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const int limit = pre + m0 + row;  // last visible key of this row (causal)
    float m_used = -INFINITY;  // max currently folded into O and l (raw score units)
    float l = 0.f;
    const float thr = kRescaleThreshold / p.scale_log2;
    for (int j = 0; j < n_tiles; ++j) {
      const int b = j & 1;
      mbar_wait(&s_full[b], (j >> 1) & 1);
      tc_fence_after_sync();
      const uint32_t ts = tmem_S + lane_base + b * kPfN;
      const bool need_mask = (j * kPfN + kPfN - 1) > (pre + m0);
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < kPfN; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(ts + c, r);
        tmem_ld_wait();
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (j * kPfN + c + i <= limit) mx = fmaxf(mx, __uint_as_float(r[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
      // lazy rescale decision is warp-uniform (tcgen05.ld/st are warp-collective)
      const bool grow = mx > m_used + thr;
      const bool do_rescale = __any_sync(0xffffffffu, grow);
      float alpha = 1.f;
      if (do_rescale) {
        const float m_new = fmaxf(m_used, mx);
        alpha = (m_used == -INFINITY) ? 0.f : exp2f((m_used - m_new) * p.scale_log2);
        l *= alpha;
        m_used = m_new;
      }
      const float m_off = (m_used == -INFINITY) ? 0.f : m_used * p.scale_log2;
      // pass 2: p = exp2(s * scale - m), packed to 16-bit in registers
      uint32_t pk[kPfN / 2];
      float psum = 0.f;
#pragma unroll
      for (int c = 0; c < kPfN; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(ts + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = fast_exp2(__uint_as_float(r[i]) * p.scale_log2 - m_off);
          float p1 = fast_exp2(__uint_as_float(r[i + 1]) * p.scale_log2 - m_off);
          if (need_mask) {
            if (j * kPfN + c + i > limit) p0 = 0.f;
            if (j * kPfN + c + i + 1 > limit) p1 = 0.f;
          }
          const uint32_t w = pack2<T>(p0, p1);
          // accumulate what the tensor core will actually see (the rounded values)
          const float2 pr = unpack2<T>(w);
          psum += pr.x + pr.y;
          pk[(c + i) >> 1] = w;
        }
      }
      l += psum;
      // S buffer b may now be overwritten by S_{j+2}
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[b]);
      // P buffer free + O stable once PV_{j-1} has completed
      if (j > 0) {
        mbar_wait(o_done, (j - 1) & 1);
        tc_fence_after_sync();
        if (do_rescale) {
          const uint32_t to = tmem_O + lane_base;
#pragma unroll 1
          for (int c = 0; c < kPfD; c += 32) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(to + c, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st_32x32b_x32(to + c, r);
          }
          tmem_st_wait();
        }
      }
      // P -> smem (K-major, 128-byte swizzle): 16 chunks of 8 values
#pragma unroll
      for (int ch = 0; ch < 16; ++ch) {
        uint4 v = make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
        *reinterpret_cast<uint4*>(sP + tile_off(row, ch)) = v;
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // epilogue: O / l
    mbar_wait(o_done, (n_tiles - 1) & 1);
    tc_fence_after_sync();
    const float inv_l = l > 0.f ? 1.f / l : 0.f;
    T* dst = reinterpret_cast<T*>(p.out) + static_cast<int64_t>(q_row0 + row) * p.out_stride +
             static_cast<int64_t>(head) * kPfD;
    const uint32_t to = tmem_O + lane_base;
#pragma unroll 1
    for (int c = 0; c < kPfD; c += 32) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(to + c, r);
      tmem_ld_wait();
      if (row < rows_valid) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 o;
          o.x = pack2<T>(__uint_as_float(r[i]) * inv_l, __uint_as_float(r[i + 1]) * inv_l);
          o.y = pack2<T>(__uint_as_float(r[i + 2]) * inv_l, __uint_as_float(r[i + 3]) * inv_l);
          o.z = pack2<T>(__uint_as_float(r[i + 4]) * inv_l, __uint_as_float(r[i + 5]) * inv_l);
          o.w = pack2<T>(__uint_as_float(r[i + 6]) * inv_l, __uint_as_float(r[i + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + c + i) = o;
        }
      }
    }
  }

  griddep_launch_dependents();
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

int check_attn_args(const char* fn, const void* out, const void* q, const void* k_pool,
                    const void* v_pool, const void* rtt, const void* rpi, int Hq, int Hk, int D,
                    int64_t q_stride, int64_t out_stride);

}  // namespace spd

using namespace spd;

extern "C" int spd_attn_prefill_ragged(void* out, const void* q, const void* k_pool,
                                       const void* v_pool, const int32_t* req_to_token,
                                       const int64_t* req_pool_indices, const int32_t* prefix_lens,
                                       const int32_t* extend_lens, const int32_t* extend_start_loc,
                                       int64_t batch, int64_t total_tokens, int64_t max_extend_len,
                                       int num_q_heads, int num_kv_heads, int head_dim,
                                       int64_t q_stride, int64_t out_stride,
                                       int64_t req_to_token_stride, float sm_scale, int dtype,
                                       spd_stream_t stream) {
  int rc = check_attn_args("spd_attn_prefill_ragged", out, q, k_pool, v_pool, req_to_token,
                           req_pool_indices, num_q_heads, num_kv_heads, head_dim, q_stride, out_stride);
  if (rc) return rc;
  SPD_REQUIRE(prefix_lens && extend_lens && extend_start_loc && batch >= 0 && total_tokens >= 0 &&
                  max_extend_len >= 0,
              SPD_E_BADARG, "spd_attn_prefill_ragged: bad arguments");
  SPD_REQUIRE(head_dim == kPfD, SPD_E_SHAPE,
              "spd_attn_prefill_ragged: head_dim must be 128 on the tcgen05 path (got %d)", head_dim);
  SPD_REQUIRE(batch <= 65535 && num_q_heads <= 65535, SPD_E_SHAPE,
              "spd_attn_prefill_ragged: batch / heads exceed the grid limits");
  if (batch == 0 || total_tokens == 0 || max_extend_len == 0) return 0;
  PrefillParams p{};
  p.out = out;
  p.k_pool = k_pool;
  p.v_pool = v_pool;
  p.req_to_token = req_to_token;
  p.req_pool_indices = req_pool_indices;
  p.prefix_lens = prefix_lens;
  p.extend_lens = extend_lens;
  p.extend_start_loc = extend_start_loc;
  p.Hq = num_q_heads;
  p.Hk = num_kv_heads;
  p.out_stride = out_stride;
  p.rtt_stride = req_to_token_stride;
  p.scale_log2 = sm_scale * 1.4426950408889634f;
  p.q_tiles = static_cast<int>((max_extend_len + kPfM - 1) / kPfM);
  CUtensorMap tq;
  if ((rc = get_tmap(&tq, q, static_cast<uint64_t>(total_tokens),
                     static_cast<uint64_t>(num_q_heads) * head_dim, static_cast<uint64_t>(q_stride), kPfM)))
    return rc;
  dim3 grid(p.q_tiles, num_q_heads, static_cast<unsigned>(batch));
  SPD_DISPATCH_DTYPE(dtype, {
    auto kern = attn_prefill_kernel<T>;
    static bool configured = false;
    if (!configured) {
      SPD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPfSmem));
      configured = true;
    }
    SPD_CUDA(launch(kern, grid, dim3(kPfThreads), kPfSmem, as_stream(stream), pdl_enabled(), tq, p));
  });
  return 0;
}
