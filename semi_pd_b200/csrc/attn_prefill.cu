// Ragged chunked-prefill ("extend") attention on tcgen05: QK^T and PV both run on the tensor
// cores with S and O accumulators in TMEM.  One pass over [prefix pages || new tokens] with the
// causal mask applied only on the diagonal tiles — no separate ragged / paged kernels and no LSE
// merge as in the flashinfer backend (S/layers/attention/flashinfer_backend.py:425-445).
// Semantics: S/layers/attention/torch_native_backend.py:26-109.
//
// PERSISTENT CTAs, one per SM of the caller's partition, walk a static work list of ITEMS.  An
// item is TWO 128-row query tiles of one request that share every K/V tile:
//   * GQA with an even group (Llama: 4 query heads per KV head): the same 128 tokens of two query
//     heads of one KV head — identical causal range, perfectly balanced;
//   * otherwise: two consecutive 128-token tiles of one query head.
// Each query tile belongs to one softmax warpgroup (own S, P, O, running max and row sum), so a
// K/V tile staged once in shared memory feeds 2 x 128 query rows (half the L2->SM traffic per
// FLOP of a single-tile CTA) and the tensor core alternates between the two warpgroups: while
// warpgroup 0 exponentiates S0_j the tensor core runs PV1_{j-1} / S1_j, and vice versa.
// Roles (448 threads):
//   warps 0-3 / 4-7  softmax warpgroup 0 / 1: one thread per query row, ONE pass per tile in the
//                    exp2 domain against a lazily raised reference max (O in TMEM is rescaled only
//                    when a row outgrows it by 2^8 — then the tile is redone; rare after tile 0);
//                    epilogue O / l straight from TMEM to global memory;
//   warps 8-9, 10-11 K / V loaders: gather rows through req_to_token with 16-byte cp.async into
//                    128-byte-swizzled smem; one 3-slot ring K_j V_j K_{j+1} ... runs ahead of the
//                    consumer ACROSS items, so an item's first tiles are in flight while the
//                    previous item drains;
//   warp 12          MMA issuer: S_w = Q_w K_j^T (K-major), O_w += P_w V_j (V MN-major);
//   warp 13          Q loader (TMA): the next item's Q tiles are fetched as soon as the current
//                    item's last S MMA has read them.
// Items are ordered heaviest (latest query tiles) first and dealt to the CTAs in snake order.
// smem: Q0 Q1 P0 P1 + 3 K/V slots = 7 x 32 KB.  TMEM: S0 | S1 | O0 | O1 = 512 columns.
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace spd {

constexpr int kPfThreads = 448;
constexpr int kPfM = 128;    // query rows per warpgroup tile
constexpr int kPfN = 128;    // kv tokens per tile
constexpr int kPfD = 128;    // head dim
constexpr int kTileBytes = 128 * 128 * 2;  // one [128 x 128] 16-bit tile = two 16 KB swizzle blocks
constexpr int kPfSlots = 3;
constexpr int kPfSmem = kTileBytes * (4 + kPfSlots) + 1024 + 512;
constexpr float kRescaleTrigger = 256.f;  // a row sum above 2^8 may hide an element above 2^8

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
  int pair_heads;  // 1: the two tiles are two heads of one KV head; 0: two consecutive token tiles
  int tiles_x;     // item rows along the token axis
  int heads_x;     // item columns along the head axis
  int batch;
  int total_items;
};

struct PfItem {
  int req, pre, hk, n;         // n = KV tiles of the item (max over the two query tiles)
  int head[2], m0[2], rows[2], nt[2], q_row0[2];
  int kv_len;
};

__device__ __forceinline__ PfItem pf_item(const PrefillParams& p, int idx) {
  PfItem it;
  const int per_x = p.heads_x * p.batch;
  const int qx = p.tiles_x - 1 - idx / per_x;  // heaviest (latest) query tiles first
  const int r = idx - (idx / per_x) * per_x;
  it.req = r / p.heads_x;
  const int hx = r - it.req * p.heads_x;
  if (p.pair_heads) {
    it.head[0] = 2 * hx;
    it.head[1] = 2 * hx + 1;
    it.m0[0] = it.m0[1] = qx * kPfM;
  } else {
    it.head[0] = it.head[1] = hx;
    it.m0[0] = 2 * qx * kPfM;
    it.m0[1] = (2 * qx + 1) * kPfM;
  }
  it.hk = it.head[0] / (p.Hq / p.Hk);
  const int ext = p.extend_lens[it.req];
  it.pre = p.prefix_lens[it.req];
  const int start = p.extend_start_loc[it.req];
  it.kv_len = 0;
#pragma unroll
  for (int w = 0; w < 2; ++w) {
    const int rows = min(kPfM, ext - it.m0[w]);
    it.rows[w] = rows > 0 ? rows : 0;
    const int kv = it.pre + it.m0[w] + it.rows[w];  // keys visible to the tile's last valid row
    it.nt[w] = rows > 0 ? (kv + kPfN - 1) / kPfN : 0;
    it.q_row0[w] = start + it.m0[w];
    if (rows > 0) it.kv_len = max(it.kv_len, kv);
  }
  it.n = max(it.nt[0], it.nt[1]);
  return it;
}

// k-th item of CTA c out of G CTAs, snake order over the heaviest-first list
__device__ __forceinline__ int pf_rank(int k, int c, int G) {
  return k * G + ((k & 1) ? (G - 1 - c) : c);
}

// byte offset of (row, 16-byte chunk) inside a [128 x 128] 16-bit tile stored as two
// [128 rows x 128 B] blocks with the 128-byte swizzle (chunk index XOR row%8)
__device__ __forceinline__ uint32_t tile_off(int row, int chunk16) {
  return (chunk16 >> 3) * 16384 + row * 128 + (((chunk16 & 7) ^ (row & 7)) << 4);
}

// One tile row for one thread: S (TMEM) -> p = exp2(s*scale - m_off) -> P (smem, 16-bit).
// Returns the row sum of p (masked entries excluded).  kMask is a template parameter so that the
// unmasked tiles run the lean loop: FFMA + MUFU.EX2 + FADD + half an F2FP per element.
template <typename T, bool kMask>
__device__ __forceinline__ float softmax_row_pass(uint32_t ts, uint8_t* sP, int row, float scale_log2,
                                                  float m_off, int kv0, int limit,
                                                  uint64_t* p_free_bar, uint32_t p_free_parity) {
  float ps0 = 0.f, ps1 = 0.f;
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
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                     // 2 tiles (warpgroup 0 / 1)
  uint8_t* sP = smem + 2 * kTileBytes;    // 2 tiles
  uint8_t* sKV = smem + 4 * kTileBytes;   // ring: position 2t = K of tile t, 2t+1 = V of tile t
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (4 + kPfSlots) * kTileBytes);
  uint64_t* q_full = bars;            // 1
  uint64_t* q_empty = bars + 1;       // 1
  uint64_t* kv_full = bars + 2;       // kPfSlots
  uint64_t* kv_empty = bars + 5;      // kPfSlots
  uint64_t* s_full = bars + 8;        // 2 (per warpgroup)
  uint64_t* s_empty = bars + 10;      // 2
  uint64_t* p_full = bars + 12;       // 2
  uint64_t* pv_done = bars + 14;      // 2
  uint64_t* o_free = bars + 16;       // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta = blockIdx.x, ncta = gridDim.x;

  if (warp == 12) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      mbar_init(q_empty, 1);
      for (int i = 0; i < kPfSlots; ++i) {
        mbar_init(&kv_full[i], 64);
        mbar_init(&kv_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&s_empty[i], 4);
        mbar_init(&p_full[i], 4);
        mbar_init(&pv_done[i], 1);
        mbar_init(&o_free[i], 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (warp == 13 && lane == 0) tma_prefetch_desc(&tmQ);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;        // 2 x 128 columns (warpgroup 0 / 1)
  const uint32_t tmem_O = tmem_base + 256;  // 2 x 128 columns

  griddep_wait();

  if (warp == 13) {
    // ===================== Q loader =====================
    if (elect_one()) {
      uint32_t itn = 0;
      for (int k = 0;; ++k) {
        const int idx = pf_rank(k, cta, ncta);
        if (idx >= p.total_items) break;
        const PfItem it = pf_item(p, idx);
        if (it.n == 0) continue;
        mbar_wait(q_empty, (itn & 1) ^ 1);
        mbar_expect_tx(q_full, ((it.rows[0] > 0) + (it.rows[1] > 0)) * kTileBytes);
#pragma unroll
        for (int w = 0; w < 2; ++w)
          if (it.rows[w] > 0) {
            tma_load_2d(sQ + w * kTileBytes, &tmQ, q_full, it.head[w] * kPfD, it.q_row0[w], kEvictFirst);
            tma_load_2d(sQ + w * kTileBytes + 16384, &tmQ, q_full, it.head[w] * kPfD + 64, it.q_row0[w],
                        kEvictFirst);
          }
        ++itn;
      }
    }
  } else if (warp == 12) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc_qk = umma_idesc<T>(kPfM, kPfN, 0, 0);  // A, B K-major
      constexpr uint32_t idesc_pv = umma_idesc<T>(kPfM, kPfD, 0, 1);  // B (V) MN-major
      uint32_t t_base = 0;          // tiles of all previous items (ring position / 2)
      uint32_t sq[2] = {0, 0};      // S tiles issued per warpgroup (all items)
      uint32_t pq[2] = {0, 0};      // PV tiles issued per warpgroup (all items)
      uint32_t eq[2] = {0, 0};      // items per warpgroup (o_free phase)
      uint32_t itn = 0;
      for (int kk = 0;; ++kk) {
        const int idx = pf_rank(kk, cta, ncta);
        if (idx >= p.total_items) break;
        const PfItem it = pf_item(p, idx);
        if (it.n == 0) continue;
        mbar_wait(q_full, itn & 1);
        auto k_ready = [&](int j, bool block) -> bool {
          const uint32_t pos = 2 * (t_base + j);
          if (block) {
            mbar_wait(&kv_full[pos % kPfSlots], (pos / kPfSlots) & 1);
            return true;
          }
          return mbar_try_wait(&kv_full[pos % kPfSlots], (pos / kPfSlots) & 1);
        };
        auto issue_s = [&](int w, int j) {  // K_j is known to be resident
          const uint32_t pos = 2 * (t_base + j);
          mbar_wait(&s_empty[w], (sq[w] & 1) ^ 1);
          tc_fence_after_sync();
          const uint32_t q_addr = smem_u32(sQ + w * kTileBytes);
          const uint32_t k_addr = smem_u32(sKV + (pos % kPfSlots) * kTileBytes);
#pragma unroll
          for (int k = 0; k < kPfD / 16; ++k) {
            const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
            umma_ss(tmem_S + w * kPfN, umma_desc_sw128(q_addr + off, 16, 1024),
                    umma_desc_sw128(k_addr + off, 16, 1024), idesc_qk, k > 0 ? 1u : 0u);
          }
          umma_commit(&s_full[w]);
          ++sq[w];
        };
        // after the S MMAs of tile j (both warpgroups) have been issued: release K_j, and Q after
        // the item's last S
        auto s_tile_issued = [&](int j) {
          umma_commit(&kv_empty[(2 * (t_base + j)) % kPfSlots]);
          if (j == it.n - 1) umma_commit(q_empty);
        };
        k_ready(0, true);
#pragma unroll
        for (int w = 0; w < 2; ++w)
          if (it.nt[w] > 0) issue_s(w, 0);
        s_tile_issued(0);
        for (int j = 0; j < it.n; ++j) {
          const uint32_t vpos = 2 * (t_base + j) + 1;
          int s_next_issued = 0;
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            if (j >= it.nt[w]) continue;
            mbar_wait(&p_full[w], pq[w] & 1);
            // S_{j+1} first when its K tile has already landed: the warpgroup can then start its
            // next softmax while the tensor core is still busy with PV_j
            bool s_issued = false;
            if (j + 1 < it.nt[w] && k_ready(j + 1, false)) {
              issue_s(w, j + 1);
              s_issued = true;
            }
            mbar_wait(&kv_full[vpos % kPfSlots], (vpos / kPfSlots) & 1);
            if (j == 0) mbar_wait(&o_free[w], (eq[w] & 1) ^ 1);  // previous item's O has been read
            tc_fence_after_sync();
            const uint32_t v_addr = smem_u32(sKV + (vpos % kPfSlots) * kTileBytes);
            const uint32_t p_addr = smem_u32(sP + w * kTileBytes);
#pragma unroll
            for (int k = 0; k < kPfN / 16; ++k) {
              // A = P[128 x 16 kv] (K-major), B = V[16 kv x 128 d] (MN-major: 2 blocks of 64 d,
              // LBO = 16 KB; kv rows 128 B apart, 8-row groups SBO = 1 KB)
              const uint32_t a_off = (k >> 2) * 16384 + (k & 3) * 32;
              umma_ss(tmem_O + w * kPfD, umma_desc_sw128(p_addr + a_off, 16, 1024),
                      umma_desc_sw128(v_addr + k * 2048, 16384, 1024), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&pv_done[w]);
            ++pq[w];
            if (j + 1 < it.nt[w] && !s_issued) {
              k_ready(j + 1, true);
              issue_s(w, j + 1);
            }
            if (j + 1 < it.nt[w]) ++s_next_issued;
          }
          umma_commit(&kv_empty[vpos % kPfSlots]);  // V_j: both PVs issued
          if (j + 1 < it.n) s_tile_issued(j + 1);
          (void)s_next_issued;
        }
#pragma unroll
        for (int w = 0; w < 2; ++w)
          if (it.nt[w] > 0) ++eq[w];
        t_base += it.n;
        ++itn;
      }
    }
  } else if (warp >= 8) {
    // ===================== K loaders (warps 8-9) / V loaders (warps 10-11) =====================
    const int is_v = warp >= 10 ? 1 : 0;
    const int lt = (threadIdx.x - 256) & 63;
    const int chunk = lt & 15;   // 16-byte chunk of the 256-byte head row
    const int rsub = lt >> 4;    // 0..3
    const T* pool = reinterpret_cast<const T*>(is_v ? p.v_pool : p.k_pool);
    const int64_t row_elems = static_cast<int64_t>(p.Hk) * kPfD;
    uint32_t t = 0;        // tiles loaded so far (all items)
    int pending = -1;      // slot whose cp.async group has been committed but not yet published
    for (int k = 0;; ++k) {
      const int idx = pf_rank(k, cta, ncta);
      if (idx >= p.total_items) break;
      const PfItem it = pf_item(p, idx);
      if (it.n == 0) continue;
      const int32_t* rtt = p.req_to_token + p.req_pool_indices[it.req] * p.rtt_stride;
      const int64_t col = static_cast<int64_t>(it.hk) * kPfD + chunk * 8;
      for (int j = 0; j < it.n; ++j, ++t) {
        // page-table lookups first, all in flight together and overlapped with the waits below (the
        // cp.async asm is a compiler barrier; interleaving the lookups with it would serialise the
        // dependent global-load latencies)
        int slot[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int kv = j * kPfN + i * 4 + rsub;
          slot[i] = kv < it.kv_len ? __ldg(rtt + kv) : -1;
        }
        // publish the previous tile BEFORE waiting for this tile's slot (the consumer of the
        // previous tile must never wait on a producer that waits on the consumer)
        if (pending >= 0) {
          cp_async_wait<0>();
          fence_proxy_async_smem();
          mbar_arrive(&kv_full[pending]);
        }
        const uint32_t pos = 2 * t + is_v;
        const int st = pos % kPfSlots;
        mbar_wait(&kv_empty[st], ((pos / kPfSlots) & 1) ^ 1);
        const uint32_t dst = smem_u32(sKV + st * kTileBytes);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int row = i * 4 + rsub;
          const bool valid = slot[i] >= 0;
          const int64_t s64 = valid ? slot[i] : 0;
          cp_async16(dst + tile_off(row, chunk), pool + s64 * row_elems + col, valid);
        }
        cp_async_commit();
        pending = st;
      }
    }
    if (pending >= 0) {
      cp_async_wait<0>();
      fence_proxy_async_smem();
      mbar_arrive(&kv_full[pending]);
    }
  } else {
    // ===================== softmax warpgroups (one thread per query row) =====================
    const int w = warp >> 2;            // warpgroup == query tile of the item
    const int row = threadIdx.x & 127;  // == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t ts = tmem_S + lane_base + w * kPfN;
    const uint32_t to = tmem_O + lane_base + w * kPfD;
    uint8_t* myP = sP + w * kTileBytes;
    uint32_t u = 0;  // tiles processed by this warpgroup (all items)
    for (int k = 0;; ++k) {
      const int idx = pf_rank(k, cta, ncta);
      if (idx >= p.total_items) break;
      const PfItem it = pf_item(p, idx);
      if (it.nt[w] == 0) continue;
      const int limit = it.pre + it.m0[w] + row;  // last visible key of this row (causal)
      float m_used = -INFINITY;  // max currently folded into O and l (raw score units)
      float l = 0.f;
      for (int j = 0; j < it.nt[w]; ++j, ++u) {
        const int kv0 = j * kPfN;
        const bool need_mask = (kv0 + kPfN - 1) > (it.pre + it.m0[w]);
        mbar_wait(&s_full[w], u & 1);
        tc_fence_after_sync();
        if (j == 0) {
          // first tile of the item: establish the reference max with a max-only pre-pass
          m_used = need_mask ? softmax_row_max<true>(ts, kv0, limit) : softmax_row_max<false>(ts, kv0, limit);
        }
        // P / O are free once this warpgroup's previous PV has completed; the pass waits for that
        // right before its first P store (the previous item's last PV was awaited by its epilogue)
        uint64_t* pfree = j > 0 ? &pv_done[w] : nullptr;
        const uint32_t pfree_par = (u - 1) & 1;
        float m_off = (m_used == -INFINITY) ? 0.f : m_used * p.scale_log2;
        float psum = need_mask ? softmax_row_pass<T, true>(ts, myP, row, p.scale_log2, m_off, kv0, limit,
                                                           pfree, pfree_par)
                               : softmax_row_pass<T, false>(ts, myP, row, p.scale_log2, m_off, kv0, limit,
                                                            pfree, pfree_par);
        // raise the reference max only when some row may have outgrown it by > 2^8 (an element
        // above 2^8 implies a row sum above 2^8; warp-uniform decision: tcgen05.ld/st are
        // warp-collective); then redo the tile against the new max
        if (__any_sync(0xffffffffu, psum > kRescaleTrigger)) {
          const float mx = need_mask ? softmax_row_max<true>(ts, kv0, limit) : softmax_row_max<false>(ts, kv0, limit);
          const float m_new = fmaxf(m_used, mx);
          const float alpha = (m_used == -INFINITY) ? 0.f : fast_exp2((m_used - m_new) * p.scale_log2);
          l *= alpha;
          m_used = m_new;
          m_off = (m_used == -INFINITY) ? 0.f : m_used * p.scale_log2;
          if (j > 0) {
            tc_fence_after_sync();  // PV_{j-1} was observed complete inside the pass
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
          psum = need_mask ? softmax_row_pass<T, true>(ts, myP, row, p.scale_log2, m_off, kv0, limit, nullptr, 0)
                           : softmax_row_pass<T, false>(ts, myP, row, p.scale_log2, m_off, kv0, limit, nullptr, 0);
        }
        l += psum;
        // S may be overwritten by S_{j+1}; P is ready for PV_j
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&s_empty[w]);
          mbar_arrive(&p_full[w]);
        }
      }
      // ---- epilogue: O / l, TMEM -> registers -> global ----
      mbar_wait(&pv_done[w], (u - 1) & 1);
      tc_fence_after_sync();
      const float inv_l = l > 0.f ? 1.f / l : 0.f;
      T* dst = reinterpret_cast<T*>(p.out) + static_cast<int64_t>(it.q_row0[w] + row) * p.out_stride +
               static_cast<int64_t>(it.head[w]) * kPfD;
#pragma unroll 1
      for (int c = 0; c < kPfD; c += 64) {
        uint32_t a[32], b[32];
        tmem_ld_32x32b_x32(to + c, a);
        tmem_ld_32x32b_x32(to + c + 32, b);
        tmem_ld_wait();
        if (c == 64) {
          // every O column of this warp is in registers: the next item's first PV may overwrite it
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&o_free[w]);
        }
        if (row < it.rows[w]) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 v;
            v.x = pack2<T>(__uint_as_float(a[i]) * inv_l, __uint_as_float(a[i + 1]) * inv_l);
            v.y = pack2<T>(__uint_as_float(a[i + 2]) * inv_l, __uint_as_float(a[i + 3]) * inv_l);
            v.z = pack2<T>(__uint_as_float(a[i + 4]) * inv_l, __uint_as_float(a[i + 5]) * inv_l);
            v.w = pack2<T>(__uint_as_float(a[i + 6]) * inv_l, __uint_as_float(a[i + 7]) * inv_l);
            *reinterpret_cast<uint4*>(dst + c + i) = v;
          }
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 v;
            v.x = pack2<T>(__uint_as_float(b[i]) * inv_l, __uint_as_float(b[i + 1]) * inv_l);
            v.y = pack2<T>(__uint_as_float(b[i + 2]) * inv_l, __uint_as_float(b[i + 3]) * inv_l);
            v.z = pack2<T>(__uint_as_float(b[i + 4]) * inv_l, __uint_as_float(b[i + 5]) * inv_l);
            v.w = pack2<T>(__uint_as_float(b[i + 6]) * inv_l, __uint_as_float(b[i + 7]) * inv_l);
            *reinterpret_cast<uint4*>(dst + c + 32 + i) = v;
          }
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
                                       int sm_budget, spd_stream_t stream) {
  int rc = check_attn_args("spd_attn_prefill_ragged", out, q, k_pool, v_pool, req_to_token,
                           req_pool_indices, num_q_heads, num_kv_heads, head_dim, q_stride, out_stride);
  if (rc) return rc;
  SPD_REQUIRE(prefix_lens && extend_lens && extend_start_loc && batch >= 0 && total_tokens >= 0 &&
                  max_extend_len >= 0,
              SPD_E_BADARG, "spd_attn_prefill_ragged: bad arguments");
  SPD_REQUIRE(head_dim == kPfD, SPD_E_SHAPE,
              "spd_attn_prefill_ragged: head_dim must be 128 on the tcgen05 path (got %d)", head_dim);
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
  const int q_tiles = static_cast<int>((max_extend_len + kPfM - 1) / kPfM);
  p.pair_heads = ((num_q_heads / num_kv_heads) % 2 == 0) ? 1 : 0;
  p.tiles_x = p.pair_heads ? q_tiles : (q_tiles + 1) / 2;
  p.heads_x = p.pair_heads ? num_q_heads / 2 : num_q_heads;
  p.batch = static_cast<int>(batch);
  const int64_t total = static_cast<int64_t>(p.tiles_x) * p.heads_x * batch;
  SPD_REQUIRE(total < (1ll << 30), SPD_E_SHAPE, "spd_attn_prefill_ragged: too many work items");
  p.total_items = static_cast<int>(total);
  CUtensorMap tq;
  if ((rc = get_tmap(&tq, q, static_cast<uint64_t>(total_tokens),
                     static_cast<uint64_t>(num_q_heads) * head_dim, static_cast<uint64_t>(q_stride), kPfM)))
    return rc;
  const int sms = sm_budget > 0 ? sm_budget : device_sm_count_cached();
  const int grid = p.total_items < sms ? p.total_items : sms;
  SPD_DISPATCH_DTYPE(dtype, {
    auto kern = attn_prefill_kernel<T>;
    static bool configured = false;
    if (!configured) {
      SPD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPfSmem));
      configured = true;
    }
    SPD_CUDA(launch(kern, dim3(grid), dim3(kPfThreads), kPfSmem, as_stream(stream), pdl_enabled(), tq, p));
  });
  return 0;
}
