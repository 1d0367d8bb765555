// Ragged chunked-prefill ("extend") attention on tcgen05: QK^T and PV both run on the tensor
// cores with S and O accumulators in TMEM.  One pass over [prefix pages || new tokens] with the
// causal mask applied only on the diagonal tiles — no separate ragged / paged kernels and no LSE
// merge as in the flashinfer backend (S/layers/attention/flashinfer_backend.py:425-445).
// Semantics: S/layers/attention/torch_native_backend.py:26-109.
//
// CTA = 128 query rows of one request x one query head (grid: q-tile (heaviest first), head,
// request; CTAs past a request's length exit).  KV is visited in 128-token tiles; EVEN tiles
// belong to softmax warpgroup 0, ODD tiles to warpgroup 1.  Each warpgroup runs an independent
// online softmax (own running max m, row sum l, own S buffer, own P buffer, own O accumulator
// in TMEM), so two tiles are in softmax at any time while the tensor core works on the next
// S / previous PV; the two (m, l, O) triples are merged in the epilogue like a 2-way split-KV.
//   K loaders (warps 8-9), V loaders (warps 10-11): gather rows through req_to_token with
//       16-byte cp.async into 128-byte-swizzled smem; K and V rings are released separately
//       (K_j right after S_j, V_j after PV_j), so K runs well ahead of V;
//   MMA (warp 12): S_j = Q K_j^T (K-major operands), O_g += P_j V_j (V as MN-major operand),
//       issue order S0 S1 | PV0 S2 | PV1 S3 | ...;
//   softmax (warps 0-3 / 4-7): one thread per query row; ONE pass per tile in the exp2 domain
//       against a lazily raised reference max (O in TMEM is rescaled only when a row max grows by
//       more than 2^8 — then the tile is redone, which besides the first tile is rare).
// Q arrives by TMA.  TMEM: S0 | S1 | O0 | O1 = 512 columns.
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace spd {

constexpr int kPfThreads = 416;  // 8 softmax + 2 K-loader + 2 V-loader + 1 MMA warps
constexpr int kPfM = 128;    // query rows per CTA
constexpr int kPfN = 128;    // kv tokens per tile
constexpr int kPfD = 128;    // head dim
constexpr int kTileBytes = 128 * 128 * 2;  // one [128 x 128] 16-bit tile = two 16 KB swizzle blocks
constexpr int kPfSmem = kTileBytes * 7 + 1024 + 512;  // Q, K0, K1, V0, V1, P0, P1 + align + barriers
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

// One tile row for one thread: S (TMEM) -> p = exp2(s*scale - m_off) -> P (smem, 16-bit).
// Returns the row sum of p; pmax receives max(p) (masked entries excluded) — p > 2^8 means the
// row outgrew the reference max.  kMask is a template parameter so that the 15-of-16 unmasked
// tiles run the lean loop: FFMA + MUFU.EX2 + FADD + FMNMX + half an F2FP per element.
template <typename T, bool kMask>
__device__ __forceinline__ float softmax_row_pass(uint32_t ts, uint8_t* sP, int row, float scale_log2,
                                                  float m_off, int kv0, int limit, float& pmax,
                                                  uint64_t* p_free_bar, uint32_t p_free_parity) {
  float ps0 = 0.f, ps1 = 0.f, pm0 = 0.f, pm1 = 0.f;
  const uint32_t prow = row * 128;
  const uint32_t sw = row & 7;
#pragma unroll 1
  for (int c = 0; c < kPfN; c += 32) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(ts + c, r);
    tmem_ld_wait();
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      float p0 = fast_exp2(fmaf(__uint_as_float(r[i]), scale_log2, -m_off));
      float p1 = fast_exp2(fmaf(__uint_as_float(r[i + 1]), scale_log2, -m_off));
      if constexpr (kMask) {
        p0 = (kv0 + c + i) <= limit ? p0 : 0.f;
        p1 = (kv0 + c + i + 1) <= limit ? p1 : 0.f;
      }
      ps0 += p0;
      ps1 += p1;
      pm0 = fmaxf(pm0, p0);
      pm1 = fmaxf(pm1, p1);
      pk[i >> 1] = pack2<T>(p0, p1);
    }
    // the P buffer is still being read by this warpgroup's previous PV: wait as late as possible
    // (the first 32 exponentials above overlap that MMA)
    if (c == 0 && p_free_bar != nullptr) mbar_wait(p_free_bar, p_free_parity);
    // chunks (c/8 .. c/8+3) of this row: all inside one 64-column swizzle block
    uint8_t* blk = sP + (c >> 6) * 16384 + prow;
    const uint32_t ch0 = (c >> 3) & 7;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
      *reinterpret_cast<uint4*>(blk + (((ch0 + ch) ^ sw) << 4)) =
          make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
  }
  pmax = fmaxf(pm0, pm1);
  return ps0 + ps1;
}

// max of the visible raw scores of one row of S (TMEM)
template <bool kMask>
__device__ __forceinline__ float softmax_row_max(uint32_t ts, int kv0, int limit) {
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < kPfN; c += 32) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(ts + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      float a = __uint_as_float(r[i]), b = __uint_as_float(r[i + 1]);
      if constexpr (kMask) {
        a = (kv0 + c + i) <= limit ? a : -INFINITY;
        b = (kv0 + c + i + 1) <= limit ? b : -INFINITY;
      }
      m0 = fmaxf(m0, a);
      m1 = fmaxf(m1, b);
    }
  }
  return fmaxf(m0, m1);
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
  uint8_t* sP = smem + 5 * kTileBytes;  // one per warpgroup
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * kTileBytes);
  uint64_t* q_full = bars;          // 1
  uint64_t* k_full = bars + 1;      // 2
  uint64_t* k_empty = bars + 3;     // 2
  uint64_t* v_full = bars + 5;      // 2
  uint64_t* v_empty = bars + 7;     // 2
  uint64_t* s_full = bars + 9;      // 2 (per warpgroup)
  uint64_t* s_empty = bars + 11;    // 2
  uint64_t* p_full = bars + 13;     // 2
  uint64_t* pv_done = bars + 15;    // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 12) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      mbar_init(q_full, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&k_full[i], 64);
        mbar_init(&k_empty[i], 1);
        mbar_init(&v_full[i], 64);
        mbar_init(&v_empty[i], 1);
        mbar_init(&s_full[i], 1);
        mbar_init(&s_empty[i], 4);
        mbar_init(&p_full[i], 4);
        mbar_init(&pv_done[i], 1);
      }
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
  const uint32_t tmem_S = tmem_base;        // 2 x 128 columns (warpgroup 0 / 1)
  const uint32_t tmem_O = tmem_base + 256;  // 2 x 128 columns

  griddep_wait();

  if (warp == 12) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      mbar_expect_tx(q_full, kTileBytes);
      tma_load_2d(sQ, &tmQ, q_full, head * kPfD, q_row0, kEvictFirst);
      tma_load_2d(sQ + 16384, &tmQ, q_full, head * kPfD + 64, q_row0, kEvictFirst);
      constexpr uint32_t idesc_qk = umma_idesc<T>(kPfM, kPfN, 0, 0);  // A, B K-major
      constexpr uint32_t idesc_pv = umma_idesc<T>(kPfM, kPfD, 0, 1);  // B (V) MN-major
      const uint32_t q_addr = smem_u32(sQ);
      auto issue_qk = [&](int j) {
        const int g = j & 1, u = j >> 1;
        mbar_wait(&k_full[g], u & 1);
        mbar_wait(&s_empty[g], (u & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t k_addr = smem_u32(sK + g * kTileBytes);
#pragma unroll
        for (int k = 0; k < kPfD / 16; ++k) {
          const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
          umma_ss(tmem_S + g * kPfN, umma_desc_sw128(q_addr + off, 16, 1024),
                  umma_desc_sw128(k_addr + off, 16, 1024), idesc_qk, k > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[g]);
        umma_commit(&k_empty[g]);  // K stage free as soon as S_j is done
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      if (n_tiles > 1) issue_qk(1);
      for (int j = 0; j < n_tiles; ++j) {
        const int g = j & 1, u = j >> 1;
        mbar_wait(&p_full[g], u & 1);
        // S_{j+2} first when its K tile has already landed: warpgroup g can then start its next
        // softmax while the tensor core is still busy with PV_j
        bool s_issued = false;
        if (j + 2 < n_tiles && mbar_try_wait(&k_full[g], (u + 1) & 1)) {
          issue_qk(j + 2);
          s_issued = true;
        }
        mbar_wait(&v_full[g], u & 1);
        tc_fence_after_sync();
        const uint32_t v_addr = smem_u32(sV + g * kTileBytes);
        const uint32_t p_addr = smem_u32(sP + g * kTileBytes);
#pragma unroll
        for (int k = 0; k < kPfN / 16; ++k) {
          // A = P[128 x 16 kv] (K-major), B = V[16 kv x 128 d] (MN-major: 2 blocks of 64 d,
          // LBO = 16 KB; kv rows 128 B apart, 8-row groups SBO = 1 KB)
          const uint32_t a_off = (k >> 2) * 16384 + (k & 3) * 32;
          umma_ss(tmem_O + g * kPfD, umma_desc_sw128(p_addr + a_off, 16, 1024),
                  umma_desc_sw128(v_addr + k * 2048, 16384, 1024), idesc_pv,
                  (u > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&v_empty[g]);
        umma_commit(&pv_done[g]);
        if (j + 2 < n_tiles && !s_issued) issue_qk(j + 2);
      }
    }
  } else if (warp >= 8) {
    // ===================== K loaders (warps 8-9) / V loaders (warps 10-11) =====================
    const bool is_v = warp >= 10;
    const int lt = (threadIdx.x - 256) & 63;
    const int chunk = lt & 15;   // 16-byte chunk of the 256-byte head row
    const int rsub = lt >> 4;    // 0..3
    const T* pool = reinterpret_cast<const T*>(is_v ? p.v_pool : p.k_pool);
    uint8_t* sbase = is_v ? sV : sK;
    uint64_t* full = is_v ? v_full : k_full;
    uint64_t* empty = is_v ? v_empty : k_empty;
    const int32_t* rtt = p.req_to_token + p.req_pool_indices[req] * p.rtt_stride;
    const int64_t row_elems = static_cast<int64_t>(p.Hk) * kPfD;
    const int64_t col = static_cast<int64_t>(hk) * kPfD + chunk * 8;
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j & 1;
      // page-table lookups first, all in flight together and overlapped with the waits below (the
      // cp.async asm is a compiler barrier; interleaving the lookups with it would serialise the
      // dependent global-load latencies)
      int slot[32];
#pragma unroll
      for (int it = 0; it < 32; ++it) {
        const int kv = j * kPfN + it * 4 + rsub;
        slot[it] = kv < kv_len ? __ldg(rtt + kv) : -1;
      }
      // publish the previous tile BEFORE waiting for this tile's stage (see the MMA issue order:
      // the consumer of tile j-1 must never wait on the producer waiting on the consumer)
      if (j > 0) {
        cp_async_wait<0>();
        fence_proxy_async_smem();
        mbar_arrive(&full[(j - 1) & 1]);
      }
      mbar_wait(&empty[st], ((j >> 1) & 1) ^ 1);
      const uint32_t dst = smem_u32(sbase + st * kTileBytes);
#pragma unroll
      for (int it = 0; it < 32; ++it) {
        const int row = it * 4 + rsub;
        const bool valid = slot[it] >= 0;
        const int64_t s64 = valid ? slot[it] : 0;
        cp_async16(dst + tile_off(row, chunk), pool + s64 * row_elems + col, valid);
      }
      cp_async_commit();
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    mbar_arrive(&full[(n_tiles - 1) & 1]);
  } else {
    // ===================== softmax warpgroups (one thread per query row) =====================
    const int g = warp >> 2;            // warpgroup: 0 = even tiles, 1 = odd tiles
    const int row = threadIdx.x & 127;  // == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int limit = pre + m0 + row;   // last visible key of this row (causal)
    const uint32_t ts = tmem_S + lane_base + g * kPfN;
    const uint32_t to = tmem_O + lane_base + g * kPfD;
    uint8_t* myP = sP + g * kTileBytes;
    float m_used = -INFINITY;  // max currently folded into O_g and l (raw score units)
    float l = 0.f;
    int u = 0;
    for (int j = g; j < n_tiles; j += 2, ++u) {
      const int kv0 = j * kPfN;
      const bool need_mask = (kv0 + kPfN - 1) > (pre + m0);
      mbar_wait(&s_full[g], u & 1);
      tc_fence_after_sync();
      if (u == 0) {
        // first tile of this warpgroup: establish the reference max with a max-only pre-pass
        m_used = need_mask ? softmax_row_max<true>(ts, kv0, limit) : softmax_row_max<false>(ts, kv0, limit);
      }
      // P_g / O_g are free once this warpgroup's previous PV has completed; the pass waits for
      // that right before its first P store
      uint64_t* pfree = u > 0 ? &pv_done[g] : nullptr;
      const uint32_t pfree_par = (u - 1) & 1;
      float m_off = (m_used == -INFINITY) ? 0.f : m_used * p.scale_log2;
      float pmax;
      float psum = need_mask ? softmax_row_pass<T, true>(ts, myP, row, p.scale_log2, m_off, kv0, limit,
                                                         pmax, pfree, pfree_par)
                             : softmax_row_pass<T, false>(ts, myP, row, p.scale_log2, m_off, kv0, limit,
                                                          pmax, pfree, pfree_par);
      // raise the reference max only when some row outgrew it by > 2^8 (warp-uniform decision:
      // tcgen05.ld/st are warp-collective); then redo the tile against the new max
      if (__any_sync(0xffffffffu, pmax > 256.f)) {
        const float mx = need_mask ? softmax_row_max<true>(ts, kv0, limit) : softmax_row_max<false>(ts, kv0, limit);
        const float m_new = fmaxf(m_used, mx);
        const float alpha = (m_used == -INFINITY) ? 0.f : fast_exp2((m_used - m_new) * p.scale_log2);
        l *= alpha;
        m_used = m_new;
        m_off = (m_used == -INFINITY) ? 0.f : m_used * p.scale_log2;
        if (u > 0) {
          tc_fence_after_sync();  // PV_{j-2} was observed complete inside the pass
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
        psum = need_mask ? softmax_row_pass<T, true>(ts, myP, row, p.scale_log2, m_off, kv0, limit, pmax,
                                                     nullptr, 0)
                         : softmax_row_pass<T, false>(ts, myP, row, p.scale_log2, m_off, kv0, limit, pmax,
                                                      nullptr, 0);
      }
      l += psum;
      // S_g may be overwritten by S_{j+2}; P_g is ready for PV_j
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&s_empty[g]);
        mbar_arrive(&p_full[g]);
      }
    }
    // ---- epilogue: merge the two warpgroups' (m, l, O) like a 2-way split-KV ----
    const int my_tiles = u;  // tiles this warpgroup processed
    if (my_tiles > 0) {
      mbar_wait(&pv_done[g], (my_tiles - 1) & 1);
      tc_fence_after_sync();
    }
    // exchange (m, l) through this warpgroup's own (now idle) P buffer
    float* stats = reinterpret_cast<float*>(myP);
    stats[row * 2] = m_used;
    stats[row * 2 + 1] = l;
    // after this barrier BOTH warpgroups have observed their last PV: O_0 and O_1 are final
    tc_fence_before_sync();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    tc_fence_after_sync();
    const float* st0 = reinterpret_cast<const float*>(sP);
    const float* st1 = reinterpret_cast<const float*>(sP + kTileBytes);
    const float m_e = st0[row * 2], l_e = st0[row * 2 + 1];
    const float m_o = st1[row * 2], l_o = st1[row * 2 + 1];
    const bool has_o = n_tiles > 1;
    const float M = fmaxf(m_e, has_o ? m_o : -INFINITY);
    const float f_e = (m_e == -INFINITY) ? 0.f : fast_exp2((m_e - M) * p.scale_log2);
    const float f_o = (!has_o || m_o == -INFINITY) ? 0.f : fast_exp2((m_o - M) * p.scale_log2);
    const float L = l_e * f_e + (has_o ? l_o * f_o : 0.f);
    const float inv_l = L > 0.f ? 1.f / L : 0.f;
    const float w_e = f_e * inv_l, w_o = f_o * inv_l;
    T* dst = reinterpret_cast<T*>(p.out) + static_cast<int64_t>(q_row0 + row) * p.out_stride +
             static_cast<int64_t>(head) * kPfD;
    const uint32_t to_e = tmem_O + lane_base, to_o = tmem_O + lane_base + kPfD;
    // warpgroup 0 writes head-dim columns 0..63, warpgroup 1 columns 64..127
#pragma unroll 1
    for (int c = g * 64; c < g * 64 + 64; c += 32) {
      uint32_t a[32], b[32];
      tmem_ld_32x32b_x32(to_e + c, a);
      if (has_o) tmem_ld_32x32b_x32(to_o + c, b);
      tmem_ld_wait();
      if (row < rows_valid) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          float o[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            o[k] = __uint_as_float(a[i + k]) * w_e;
            if (has_o) o[k] += __uint_as_float(b[i + k]) * w_o;
          }
          uint4 v;
          v.x = pack2<T>(o[0], o[1]);
          v.y = pack2<T>(o[2], o[3]);
          v.z = pack2<T>(o[4], o[5]);
          v.w = pack2<T>(o[6], o[7]);
          *reinterpret_cast<uint4*>(dst + c + i) = v;
        }
      }
    }
  }

  griddep_launch_dependents();
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tmem_base, 512);
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
