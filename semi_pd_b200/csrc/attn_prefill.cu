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
// FLOP of a single-tile CTA).
// HALF-TILE UNITS.  A 128-token K/V tile is consumed as two independent 64-column halves a / b of
// the warpgroup's S tile (S_a = Q K[0:64]^T in TMEM columns 0-63, S_b in 64-127, each with its own
// full / P-ready barriers).  While a warpgroup exponentiates half b of tile j, the tensor core
// runs PV_a(j) and S_a(j+1) for it, so its softmax never waits for the S -> softmax -> PV -> S
// chain of the whole tile (ncu on the whole-tile version: softmax warps 50 % of their time on
// s_full, tensor pipe 44 %); the second warpgroup fills the remaining tensor time.
// Roles (384 threads):
//   warps 0-3 / 4-7  softmax warpgroup 0 / 1: one thread per query row, ONE pass per tile in the
//                    exp2 domain against a lazily raised reference max (O in TMEM is rescaled only
//                    when a row outgrows it by 2^8 — then the tile is redone; rare after tile 0).
//                    P never touches shared memory: the 16-bit probabilities are written back
//                    into TMEM over the first 32 columns of the S half they came from
//                    (tcgen05.st) and feed the PV MMA as its A operand from TMEM;
//                    epilogue O / l straight from TMEM to global memory;
//   warp 8, warp 9   K / V loader into 128-byte-swizzled smem; one 5-slot ring K_j V_j K_{j+1} ...
//                    runs ahead of the consumer ACROSS items, so an item's first tiles are in
//                    flight while the previous item drains.  A tile whose 128 page-table entries
//                    are consecutive slots (what the run allocator produces, memory_pool.py) is a
//                    strided 2-D box of the pool and is fetched with two TMA tensor copies by one
//                    lane; any other tile (scattered slots, the ragged tail of a sequence) is
//                    gathered row by row through req_to_token with 16-byte cp.async;
//   warp 10          MMA issuer: S_w,h = Q_w K_j[h]^T (both K-major in smem), O_w += P_w,h V_j[h] (P
//                    from TMEM, V MN-major in smem); per warpgroup and half the order is PV(j,h)
//                    then S(j+1,h) (which overwrites the columns P(j,h) lives in; tcgen05.mma
//                    executes in order);
//   warp 11          Q loader (TMA): the next item's Q tiles are fetched as soon as the current
//                    item's last S MMA has read them.
// Items are ordered heaviest (latest query tiles) first and dealt to the CTAs in snake order.
// smem: Q0 Q1 + 5 K/V slots = 7 x 32 KB.  TMEM: S0a/P0a S0b/P0b | S1a/P1a S1b/P1b | O0 | O1 = 512 columns.
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace spd {

constexpr int kPfThreads = 384;  // 12 warps -> 168 registers per thread for the softmax warps
constexpr int kPfM = 128;    // query rows per warpgroup tile
constexpr int kPfN = 128;    // kv tokens per tile
constexpr int kPfH = 64;     // kv tokens per half-tile unit (columns of one S half)
constexpr int kPfD = 128;    // head dim
constexpr int kTileBytes = 128 * 128 * 2;  // one [128 x 128] 16-bit tile = two 16 KB swizzle blocks
constexpr int kPfSlots = 4;
constexpr int kPfStage = 2 * 16384;  // epilogue staging: one [128 x 64] 16-bit block per warpgroup
constexpr int kPfSmem = kTileBytes * (2 + kPfSlots) + kPfStage + 1024 + 512;
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
  int use_tma;  // 0: always gather (SPD_ATTN_TMA=0)
  // n / d for the three launch-constant divisors of the item decode (per_x = heads_x * batch,
  // heads_x, GQA group) as multiply-high + shift: four dependent integer divisions sat on every
  // item boundary of every role
  uint32_t per_x_m, per_x_s, heads_x_m, heads_x_s, group_m, group_s;
  long long* trace;  // debug timeline of CTA 0 (scripts/debug_attn_prefill_timeline.py); kTrace builds only
  int trace_cap;
};

// timeline region r: (tag, clock64) pairs written by one thread of CTA 0
template <bool kTrace>
__device__ __forceinline__ void pf_trace(const PrefillParams& p, int region, uint32_t& n, long long tag) {
  if constexpr (kTrace) {
    if (blockIdx.x == 0 && p.trace && static_cast<int>(n) < p.trace_cap) {
      long long* e = p.trace + (static_cast<size_t>(region) * p.trace_cap + n) * 2;
      e[0] = tag;
      e[1] = clock64();
      ++n;
    }
  }
}

struct PfItem {
  int req, pre, hk, n;         // n = KV tiles of the item (max over the two query tiles)
  int head[2], m0[2], rows[2], nt[2], q_row0[2];
  int kv_len;
};

// floor(n / d), n < 2^31, with (m, s) from fastdiv_magic(d)
__device__ __forceinline__ int fastdiv(int n, uint32_t m, uint32_t s) {
  return static_cast<int>((static_cast<uint64_t>(__umulhi(static_cast<uint32_t>(n), m)) + static_cast<uint32_t>(n)) >> s);
}
// Granlund-Montgomery round-up magic: s = ceil(log2 d), m = floor(2^32 * (2^s - d) / d) + 1
inline void fastdiv_magic(uint32_t d, uint32_t& m, uint32_t& s) {
  s = 0;
  while ((1ull << s) < d) ++s;
  m = static_cast<uint32_t>(((1ull << 32) * ((1ull << s) - d)) / d + 1);
}

// the three per-request values an item needs from global memory; loaded ONE ITEM AHEAD by every
// role (the dependent loads used to sit between two items: ~1000 clk per item boundary)
struct PfRaw {
  int ext, pre, start;
};
__device__ __forceinline__ PfRaw pf_raw(const PrefillParams& p, int idx) {
  PfRaw r{0, 0, 0};
  if (idx < p.total_items) {
    const int per_x = p.heads_x * p.batch;
    const int req = fastdiv(idx - fastdiv(idx, p.per_x_m, p.per_x_s) * per_x, p.heads_x_m, p.heads_x_s);
    r.ext = __ldg(p.extend_lens + req);
    r.pre = __ldg(p.prefix_lens + req);
    r.start = __ldg(p.extend_start_loc + req);
  }
  return r;
}

__device__ __forceinline__ PfItem pf_item(const PrefillParams& p, int idx, const PfRaw& raw) {
  PfItem it;
  const int per_x = p.heads_x * p.batch;
  const int ix = fastdiv(idx, p.per_x_m, p.per_x_s);
  const int qx = p.tiles_x - 1 - ix;  // heaviest (latest) query tiles first
  const int r = idx - ix * per_x;
  it.req = fastdiv(r, p.heads_x_m, p.heads_x_s);
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
  it.hk = fastdiv(it.head[0], p.group_m, p.group_s);
  const int ext = raw.ext;
  it.pre = raw.pre;
  const int start = raw.start;
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

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (here the 16-bit probabilities, two per 32-bit
// column, row = TMEM lane) is read from tensor memory
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Packed fp32 pairs (sm_100 FFMA2 / FADD2): one instruction per two elements.  The softmax warps
// are issue-bound around the MUFU stream (one warp per scheduler and warpgroup), so halving the
// FFMA / FADD count is what shortens a row pass.
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// 32 columns of one tile row: p = exp2(s*scale - m_off), packed to 16 bits into pk[0..16)
template <typename T, bool kMask>
__device__ __forceinline__ void softmax_chunk(const uint32_t* r, int c, uint32_t* pk, uint64_t scale2,
                                              uint64_t neg_m2, int kv0, int limit, uint64_t& ps) {
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float x0, x1;
    unpack_f32x2(ffma2(pack_f32x2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), scale2, neg_m2), x0, x1);
    float p0 = fast_exp2(x0);
    float p1 = fast_exp2(x1);
    if constexpr (kMask) {
      p0 = (kv0 + c + i) <= limit ? p0 : 0.f;
      p1 = (kv0 + c + i + 1) <= limit ? p1 : 0.f;
    }
    ps = fadd2(ps, pack_f32x2(p0, p1));
    pk[i >> 1] = pack2<T>(p0, p1);
  }
}

// One row of one S unit (kU = 64 or 128 columns) for one thread: S (TMEM) -> kU/2 packed 16-bit
// pairs in registers.  Returns the row sum of p (masked entries excluded).  Two 32-column TMEM
// loads are in flight while the previous 32 columns are exponentiated.  kMask is a template
// parameter so that the unmasked units run the lean loop: FFMA + MUFU.EX2 + FADD + half an F2FP
// per element.
template <typename T, bool kMask, int kU>
__device__ __forceinline__ float softmax_row_pass(uint32_t ts, uint32_t* pk, float scale_log2, float m_off,
                                                  int kv0, int limit) {
  const uint64_t scale2 = pack_f32x2(scale_log2, scale_log2);
  const uint64_t neg_m2 = pack_f32x2(-m_off, -m_off);
  uint64_t ps = pack_f32x2(0.f, 0.f);
  uint32_t ra[32], rb[32];
  tmem_ld_32x32b_x32(ts, ra);
  tmem_ld_32x32b_x32(ts + 32, rb);
  tmem_ld_wait();
  softmax_chunk<T, kMask>(ra, 0, pk, scale2, neg_m2, kv0, limit, ps);
  if constexpr (kU == 128) tmem_ld_32x32b_x32(ts + 64, ra);  // in flight under the next chunk
  softmax_chunk<T, kMask>(rb, 32, pk + 16, scale2, neg_m2, kv0, limit, ps);
  if constexpr (kU == 128) {
    tmem_ld_32x32b_x32(ts + 96, rb);
    tmem_ld_wait();
    softmax_chunk<T, kMask>(ra, 64, pk + 32, scale2, neg_m2, kv0, limit, ps);
    softmax_chunk<T, kMask>(rb, 96, pk + 48, scale2, neg_m2, kv0, limit, ps);
  }
  float ps0, ps1;
  unpack_f32x2(ps, ps0, ps1);
  return ps0 + ps1;
}

// max of the visible raw scores of one row of an S unit (TMEM)
template <bool kMask>
__device__ __forceinline__ void rowmax_chunk(const uint32_t* r, int c, int kv0, int limit, float& m0, float& m1) {
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
template <bool kMask, int kU>
__device__ __forceinline__ float softmax_row_max(uint32_t ts, int kv0, int limit) {
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int c = 0; c < kU; c += 64) {
    uint32_t ra[32], rb[32];
    tmem_ld_32x32b_x32(ts + c, ra);
    tmem_ld_32x32b_x32(ts + c + 32, rb);
    tmem_ld_wait();
    rowmax_chunk<kMask>(ra, c, kv0, limit, m0, m1);
    rowmax_chunk<kMask>(rb, c + 32, kv0, limit, m0, m1);
  }
  return fmaxf(m0, m1);
}

// kU: columns of one softmax / MMA unit.  64: the half-tile pipeline described above.  128: one
// unit per K/V tile — S = Q K^T goes out as ONE N = 128 group (8 instead of 16 SS-form MMAs per
// tile, the form whose fixed cost per instruction dominates at N = 64) and the two warpgroups
// alternate on the tensor core tile by tile.
template <typename T, int kU, bool kTrace>
__global__ void __launch_bounds__(kPfThreads, 1)
attn_prefill_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                    const PrefillParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                     // 2 tiles (warpgroup 0 / 1)
  uint8_t* sKV = smem + 2 * kTileBytes;   // ring: position 2t = K of tile t, 2t+1 = V of tile t
  uint8_t* sO = smem + (2 + kPfSlots) * kTileBytes;  // epilogue staging, 16 KB per warpgroup
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO + kPfStage);
  uint64_t* q_full = bars;                     // 1
  uint64_t* q_empty = bars + 1;                // 1
  uint64_t* kv_full = bars + 2;                // kPfSlots
  uint64_t* kv_empty = kv_full + kPfSlots;     // kPfSlots
  uint64_t* s_full = kv_empty + kPfSlots;      // 4: [warpgroup][half]
  uint64_t* p_full = s_full + 4;               // 4: [warpgroup][half]
  uint64_t* pv_done = p_full + 4;              // 4: [warpgroup][half] (see the parity note in the softmax)
  uint64_t* o_free = pv_done + 4;              // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta = blockIdx.x, ncta = gridDim.x;
  constexpr int kNU = kPfN / kU;  // units per K/V tile

  if (warp == 10) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      mbar_init(q_empty, 1);
      for (int i = 0; i < kPfSlots; ++i) {
        mbar_init(&kv_full[i], 32);
        mbar_init(&kv_empty[i], 1);
      }
      for (int i = 0; i < 4; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&p_full[i], 4);
        mbar_init(&pv_done[i], 1);
      }
      for (int i = 0; i < 2; ++i) mbar_init(&o_free[i], 4);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (warp == 11 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmO);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;        // 2 x (64 + 64) columns (warpgroup 0 / 1, half a / b); P_h aliases
                                            // the first 32 columns of S_h
  const uint32_t tmem_O = tmem_base + 256;  // 2 x 128 columns

  griddep_wait();

  if (warp == 11) {
    // ===================== Q loader =====================
    if (elect_one()) {
      uint32_t itn = 0;
      int idx = pf_rank(0, cta, ncta);
      PfRaw raw = pf_raw(p, idx);
      for (int k = 0; idx < p.total_items; ++k) {
        const PfItem it = pf_item(p, idx, raw);
        idx = pf_rank(k + 1, cta, ncta);
        raw = pf_raw(p, idx);
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
  } else if (warp == 10) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc_qk = umma_idesc<T>(kPfM, kU, 0, 0);  // A, B K-major; N = one unit
      constexpr uint32_t idesc_pv = umma_idesc<T>(kPfM, kPfD, 0, 1);  // A = P (TMEM), B = V MN-major
      uint32_t t_base = 0;          // tiles of all previous items (ring position / 2)
      uint32_t pq[2] = {0, 0};      // PV units issued per warpgroup (all items); unit of the tile = pq % kNU
      uint32_t eq[2] = {0, 0};      // items per warpgroup (o_free phase)
      uint32_t itn = 0;
      uint32_t tn = 0;
      int idx = pf_rank(0, cta, ncta);
      PfRaw raw = pf_raw(p, idx);
      for (int kk = 0; idx < p.total_items; ++kk) {
        const PfItem it = pf_item(p, idx, raw);
        idx = pf_rank(kk + 1, cta, ncta);
        raw = pf_raw(p, idx);
        if (it.n == 0) continue;
        mbar_wait(q_full, itn & 1);
        // S_w(j, h): the warpgroup's previous P in these columns has been consumed by PV(j-1, h),
        // issued earlier by this thread — tcgen05.mma executes in issue order.
        // (Tried in round 2: interleaving the two warpgroups' MMAs instruction by instruction —
        // back-to-back tcgen05.mma on ONE accumulator cost ~43 clk (SS) / ~10 clk (TS) more than
        // N/2 each, alternating accumulators runs at the shared-memory operand rate,
        // profiles/r02_tcgen05_2cta_probe.log.  The issuer's time per unit pair did drop from
        // 2 x 795 to 1293 clk, but pairing forces both warpgroups' softmax passes into the same
        // phase (1020 -> 1500 clk per pass, XU contention) and the kernel got 2-5 % slower.)
        auto issue_s = [&](int w, int j, int h) {
          const uint32_t pos = 2 * (t_base + j);
          mbar_wait(&kv_full[pos % kPfSlots], (pos / kPfSlots) & 1);
          tc_fence_after_sync();
          const uint32_t q_addr = smem_u32(sQ + w * kTileBytes);
          const uint32_t k_addr = smem_u32(sKV + (pos % kPfSlots) * kTileBytes) + h * (kU * 128);
#pragma unroll
          for (int k = 0; k < kPfD / 16; ++k) {
            const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
            umma_ss(tmem_S + w * kPfN + h * kU, umma_desc_sw128(q_addr + off, 16, 1024),
                    umma_desc_sw128(k_addr + off, 16, 1024), idesc_qk, k > 0 ? 1u : 0u);
          }
          umma_commit(&s_full[w * 2 + h]);
        };
        // after every unit of S(j) of both warpgroups has been issued: release K_j, and Q after
        // the item's last S
        auto s_tile_issued = [&](int j) {
          umma_commit(&kv_empty[(2 * (t_base + j)) % kPfSlots]);
          if (j == it.n - 1) umma_commit(q_empty);
        };
        // the epilogue of the previous item reads O, not S — and its last PVs (reading P in the S
        // columns) were issued before these, so S(0, .) may go out immediately
#pragma unroll
        for (int h = 0; h < kNU; ++h)
#pragma unroll
          for (int w = 0; w < 2; ++w)
            if (it.nt[w] > 0) issue_s(w, 0, h);
        s_tile_issued(0);
        for (int j = 0; j < it.n; ++j) {
          const uint32_t vpos = 2 * (t_base + j) + 1;
#pragma unroll
          for (int h = 0; h < kNU; ++h) {
#pragma unroll
            for (int w = 0; w < 2; ++w) {
              if (j >= it.nt[w]) continue;
              pf_trace<kTrace>(p, 2, tn, 0x100 + w * 2 + h);  // waiting for P
              mbar_wait(&p_full[w * 2 + h], (pq[w] / kNU) & 1);
              pf_trace<kTrace>(p, 2, tn, 0x200 + w * 2 + h);  // P ready
              mbar_wait(&kv_full[vpos % kPfSlots], (vpos / kPfSlots) & 1);
              if (j == 0 && h == 0) mbar_wait(&o_free[w], (eq[w] & 1) ^ 1);  // previous item's O has been read
              tc_fence_after_sync();
              const uint32_t v_addr = smem_u32(sKV + (vpos % kPfSlots) * kTileBytes) + h * (kU * 128);
#pragma unroll
              for (int k = 0; k < kU / 16; ++k) {
                // A = P[128 x 16 kv]: 8 TMEM columns per k-step; B = V[16 kv x 128 d] (MN-major: 2
                // blocks of 64 d, LBO = 16 KB; kv rows 128 B apart, 8-row groups SBO = 1 KB)
                umma_ts(tmem_O + w * kPfD, tmem_S + w * kPfN + h * kU + k * 8,
                        umma_desc_sw128(v_addr + k * 2048, 16384, 1024), idesc_pv,
                        (j > 0 || h > 0 || k > 0) ? 1u : 0u);
              }
              umma_commit(&pv_done[w * 2 + h]);
              ++pq[w];
              if (j + 1 < it.nt[w]) issue_s(w, j + 1, h);
              pf_trace<kTrace>(p, 2, tn, 0x300 + w * 2 + h);  // PV (+ next S) issued
            }
          }
          umma_commit(&kv_empty[vpos % kPfSlots]);  // V_j: all PVs issued
          if (j + 1 < it.n) s_tile_issued(j + 1);
        }
#pragma unroll
        for (int w = 0; w < 2; ++w)
          if (it.nt[w] > 0) ++eq[w];
        t_base += it.n;
        ++itn;
      }
    }
  } else if (warp >= 8) {
    // ===================== K loader (warp 8) / V loader (warp 9) =====================
    const int is_v = warp == 9 ? 1 : 0;
    const int chunk = lane & 15;   // 16-byte chunk of the 256-byte head row
    const int rsub = lane >> 4;    // 0..1
    const T* pool = reinterpret_cast<const T*>(is_v ? p.v_pool : p.k_pool);
    const int64_t row_elems = static_cast<int64_t>(p.Hk) * kPfD;
    uint32_t t = 0;        // tiles loaded so far (all items)
    int pending = -1;      // slot whose cp.async group is committed but not yet published
    int idx = pf_rank(0, cta, ncta);
    PfRaw raw = pf_raw(p, idx);
    for (int k = 0; idx < p.total_items; ++k) {
      const PfItem it = pf_item(p, idx, raw);
      idx = pf_rank(k + 1, cta, ncta);
      raw = pf_raw(p, idx);
      if (it.n == 0) continue;
      const int32_t* rtt = p.req_to_token + p.req_pool_indices[it.req] * p.rtt_stride;
      const int64_t col = static_cast<int64_t>(it.hk) * kPfD + chunk * 8;
      for (int j = 0; j < it.n; ++j, ++t) {
        // page-table lookups first, all in flight together (the cp.async asm is a compiler
        // barrier; interleaving the lookups with it would serialise the dependent load latencies)
        int slot[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const int kv = j * kPfN + i * 2 + rsub;
          slot[i] = kv < it.kv_len ? __ldg(rtt + kv) : -1;
        }
        // consecutive slots?  (full tile whose 128 entries are base, base+1, ...; warp vote)
        const int base = (p.use_tma && (j + 1) * kPfN <= it.kv_len) ? __ldg(rtt + j * kPfN) : -1;
        bool contig = base > 0;
#pragma unroll
        for (int i = 0; i < 64; ++i) contig &= (slot[i] == base + i * 2 + rsub);
        contig = __all_sync(0xffffffffu, contig);
        const uint32_t pos = 2 * t + is_v;
        const int st = pos % kPfSlots;
        const uint32_t epar = ((pos / kPfSlots) & 1) ^ 1;
        // Keep two tiles in flight when the ring has room.  If this tile's slot is still in use,
        // publish the previous tile BEFORE blocking on the slot: the consumer of the previous
        // tile must never wait on a producer that waits on the consumer.
        bool publish_after = true;
        if (pending >= 0 && !mbar_try_wait(&kv_empty[st], epar)) {
          cp_async_wait<0>();
          fence_proxy_async_smem();
          mbar_arrive(&kv_full[pending]);
          pending = -1;
          publish_after = false;
        }
        mbar_wait(&kv_empty[st], epar);
        if (contig) {
          // TMA: rows base .. base+127 of the pool, the 128 columns of kv head hk, as the slot's
          // two 16 KB swizzle blocks; the 32 KB are counted on the slot's barrier next to the
          // warp's 32 arrivals
          if (lane == 0) {
            uint8_t* dst = sKV + st * kTileBytes;
            fence_proxy_async_smem();  // the slot may last have been written by the gather path
            mbar_expect_tx_noarrive(&kv_full[st], kTileBytes);
            tma_load_2d(dst, is_v ? &tmV : &tmK, &kv_full[st], it.hk * kPfD, base, kEvictNormal);
            tma_load_2d(dst + 16384, is_v ? &tmV : &tmK, &kv_full[st], it.hk * kPfD + 64, base, kEvictNormal);
          }
          __syncwarp();
          mbar_arrive(&kv_full[st]);
          if (pending >= 0) {  // a gathered tile is still unpublished: no later group will cover it
            cp_async_wait<0>();
            fence_proxy_async_smem();
            mbar_arrive(&kv_full[pending]);
            pending = -1;
          }
          continue;
        }
        const uint32_t dst = smem_u32(sKV + st * kTileBytes);
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const int row = i * 2 + rsub;
          const bool valid = slot[i] >= 0;
          const int64_t s64 = valid ? slot[i] : 0;
          cp_async16(dst + tile_off(row, chunk), pool + s64 * row_elems + col, valid);
        }
        cp_async_commit();
        if (publish_after && pending >= 0) {
          cp_async_wait<1>();  // everything but the group just committed has landed
          fence_proxy_async_smem();
          mbar_arrive(&kv_full[pending]);
        }
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
    uint32_t u = 0;  // units processed by this warpgroup (all items); unit of the tile = u % kNU
    uint32_t tn = 0;
    const bool tracer = kTrace && (threadIdx.x & 127) == 0;
    // ---- epilogue: O / l, TMEM -> registers -> global ----
    // (Tried in round 2: DEFERRING it until the next item's first unit has been exponentiated, so
    // that the ~1500-1900 clk drain of the item's last PV is spent on useful work.  The drain wait
    // did shrink to ~700 clk, but o_free — and with it the next item's first PV — moved later by
    // a whole softmax pass, and the kernel got 5 % slower at 8 x 1024; not kept.)
    float e_l = 0.f;  // state of the item being written out
    int e_rows = 0, e_head = 0, e_q_row0 = 0;
    uint32_t e_u = 0;
    auto epilogue = [&]() {
      // last unit = half b of the last tile; S of that unit complete => PV(b) of the tile before is
      // complete, so only this one phase of pv_done[w][b] can be outstanding; PV(b) complete => PV(a)
      mbar_wait(&pv_done[w * 2 + (kNU - 1)], ((e_u - 1) / kNU) & 1);
      tc_fence_after_sync();
      if (tracer) pf_trace<kTrace>(p, w, tn, 0x500);  // last PV complete
      const float inv_l = e_l > 0.f ? 1.f / e_l : 0.f;
      // A full tile leaves through shared memory and one TMA store per 64-column block: with one
      // thread per row, direct 16-byte stores touch 32 different lines per warp instruction and
      // the 128 rows x 256 B took ~5900 clk per item (in-kernel timeline, round 2) — a quarter of
      // the kernel at 1024-token prompts.  A ragged last tile (fewer than 128 rows) must not write
      // the rows of the next request and keeps the direct path.
      const bool full_tile = e_rows == kPfM;
      uint8_t* stage = sO + w * (kPfStage / 2);
      T* dst = reinterpret_cast<T*>(p.out) + static_cast<int64_t>(e_q_row0 + row) * p.out_stride +
               static_cast<int64_t>(e_head) * kPfD;
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
        uint4 v[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[i].x = pack2<T>(__uint_as_float(a[8 * i]) * inv_l, __uint_as_float(a[8 * i + 1]) * inv_l);
          v[i].y = pack2<T>(__uint_as_float(a[8 * i + 2]) * inv_l, __uint_as_float(a[8 * i + 3]) * inv_l);
          v[i].z = pack2<T>(__uint_as_float(a[8 * i + 4]) * inv_l, __uint_as_float(a[8 * i + 5]) * inv_l);
          v[i].w = pack2<T>(__uint_as_float(a[8 * i + 6]) * inv_l, __uint_as_float(a[8 * i + 7]) * inv_l);
          v[4 + i].x = pack2<T>(__uint_as_float(b[8 * i]) * inv_l, __uint_as_float(b[8 * i + 1]) * inv_l);
          v[4 + i].y = pack2<T>(__uint_as_float(b[8 * i + 2]) * inv_l, __uint_as_float(b[8 * i + 3]) * inv_l);
          v[4 + i].z = pack2<T>(__uint_as_float(b[8 * i + 4]) * inv_l, __uint_as_float(b[8 * i + 5]) * inv_l);
          v[4 + i].w = pack2<T>(__uint_as_float(b[8 * i + 6]) * inv_l, __uint_as_float(b[8 * i + 7]) * inv_l);
        }
        if (full_tile) {
          // the block's previous TMA store has finished READING the staging memory
          if (row == 0) bulk_wait_group_read0();
          asm volatile("bar.sync %0, 128;" ::"r"(2 + w) : "memory");
#pragma unroll
          for (int i = 0; i < 8; ++i)  // [128 rows x 128 B], 128-byte swizzle: chunk ^ (row % 8)
            *reinterpret_cast<uint4*>(stage + row * 128 + ((i ^ (row & 7)) << 4)) = v[i];
          fence_proxy_async_smem();
          asm volatile("bar.sync %0, 128;" ::"r"(2 + w) : "memory");
          if (row == 0) {
            tma_store_2d(&tmO, stage, e_head * kPfD + c, e_q_row0);
            bulk_commit_group();
          }
        } else if (row < e_rows) {
#pragma unroll
          for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(dst + c + 8 * i) = v[i];
        }
      }
      if (tracer) pf_trace<kTrace>(p, w, tn, 0x600);  // O stored
    };
    int idx = pf_rank(0, cta, ncta);
    PfRaw raw = pf_raw(p, idx);
    for (int k = 0; idx < p.total_items; ++k) {
      const PfItem it = pf_item(p, idx, raw);
      idx = pf_rank(k + 1, cta, ncta);
      raw = pf_raw(p, idx);
      const int nt = w == 0 ? it.nt[0] : it.nt[1];
      if (nt == 0) continue;
      const int m0 = w == 0 ? it.m0[0] : it.m0[1];
      const int rows_valid = w == 0 ? it.rows[0] : it.rows[1];
      const int head = w == 0 ? it.head[0] : it.head[1];
      const int q_row0 = w == 0 ? it.q_row0[0] : it.q_row0[1];
      const int limit = it.pre + m0 + row;  // last visible key of this row (causal)
      float m_used = -INFINITY;  // max currently folded into O and l (raw score units)
      float l = 0.f;
      for (int jh = 0; jh < kNU * nt; ++jh, ++u) {
        const int h = jh % kNU;
        const int kv0 = (jh / kNU) * kPfN + h * kU;
        const bool need_mask = (kv0 + kU - 1) > (it.pre + m0);
        const uint32_t tsh = ts + h * kU;
        if (tracer) pf_trace<kTrace>(p, w, tn, 0x100 + h);  // waiting for S
        mbar_wait(&s_full[w * 2 + h], (u / kNU) & 1);
        tc_fence_after_sync();
        if (tracer) pf_trace<kTrace>(p, w, tn, 0x200 + h);  // S ready
        if (jh == 0) {
          // first unit of the item: establish the reference max with a max-only pre-pass
          m_used = need_mask ? softmax_row_max<true, kU>(tsh, kv0, limit) : softmax_row_max<false, kU>(tsh, kv0, limit);
        }
        float m_off = (m_used == -INFINITY) ? 0.f : m_used * p.scale_log2;
        uint32_t pk[kU / 2];
        float psum = need_mask ? softmax_row_pass<T, true, kU>(tsh, pk, p.scale_log2, m_off, kv0, limit)
                               : softmax_row_pass<T, false, kU>(tsh, pk, p.scale_log2, m_off, kv0, limit);
        // raise the reference max only when some row may have outgrown it by > 2^8 (an element
        // above 2^8 implies a row sum above 2^8; warp-uniform decision: tcgen05.ld/st are
        // warp-collective); S is still intact (P is stored below), so the unit is simply redone
        if (__any_sync(0xffffffffu, psum > kRescaleTrigger)) {
          const float mx = need_mask ? softmax_row_max<true, kU>(tsh, kv0, limit)
                                     : softmax_row_max<false, kU>(tsh, kv0, limit);
          const float m_new = fmaxf(m_used, mx);
          const float alpha = (m_used == -INFINITY) ? 0.f : fast_exp2((m_used - m_new) * p.scale_log2);
          l *= alpha;
          m_used = m_new;
          m_off = (m_used == -INFINITY) ? 0.f : m_used * p.scale_log2;
          if (jh > 0) {
            // O must be quiescent: S of this unit was issued after PV of unit u-kNU (so units
            // <= u-kNU are complete), but PV of the units in between (the other half) may still
            // be running.  pv_done is per unit-of-tile so that this wait cannot alias: on that
            // barrier at most the one phase we wait for is outstanding (a parity wait only tells
            // the current phase from the one before it).
            mbar_wait(&pv_done[w * 2 + ((u - 1) % kNU)], ((u - 1) / kNU) & 1);
            tc_fence_after_sync();
#pragma unroll 1
            for (int c = 0; c < kPfD; c += 32) {
              uint32_t r[32];
              tmem_ld_32x32b_x32(to + c, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
              tmem_st_32x32b_x32(to + c, r);
            }
          }
          psum = need_mask ? softmax_row_pass<T, true, kU>(tsh, pk, p.scale_log2, m_off, kv0, limit)
                           : softmax_row_pass<T, false, kU>(tsh, pk, p.scale_log2, m_off, kv0, limit);
        }
        l += psum;
        // P over the first kU/2 columns of this S unit (every column of it is in registers by now)
        tmem_st_32x32b_x32(tsh, pk);
        if constexpr (kU == 128) tmem_st_32x32b_x32(tsh + 32, pk + 32);
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[w * 2 + h]);
        if (tracer) pf_trace<kTrace>(p, w, tn, 0x300 + h);  // P published
      }
      if (tracer) pf_trace<kTrace>(p, w, tn, 0x400);  // item done
      e_l = l;
      e_rows = rows_valid;
      e_head = head;
      e_q_row0 = q_row0;
      e_u = u;
      epilogue();
    }
  }

  if (warp < 8 && (threadIdx.x & 127) == 0) bulk_wait_group0();  // this thread's TMA stores have landed
  griddep_launch_dependents();
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 10) tmem_dealloc(tmem_base, 512);
}

int check_attn_args(const char* fn, const void* out, const void* q, const void* k_pool,
                    const void* v_pool, const void* rtt, const void* rpi, int Hq, int Hk, int D,
                    int64_t q_stride, int64_t out_stride);

}  // namespace spd

using namespace spd;

// debug only (not part of include/spd.h): subsequent launches run the traced build and CTA 0 writes
// its timeline — 3 regions (softmax warpgroup 0, 1, MMA issuer) of `cap` (tag, clock) pairs
static long long* g_pf_trace = nullptr;
static int g_pf_trace_cap = 0;
extern "C" int spd_debug_attn_prefill_trace(void* buf, int64_t cap) {
  g_pf_trace = static_cast<long long*>(buf);
  g_pf_trace_cap = static_cast<int>(cap);
  return 0;
}

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
  fastdiv_magic(static_cast<uint32_t>(p.heads_x) * static_cast<uint32_t>(p.batch), p.per_x_m, p.per_x_s);
  fastdiv_magic(static_cast<uint32_t>(p.heads_x), p.heads_x_m, p.heads_x_s);
  fastdiv_magic(static_cast<uint32_t>(num_q_heads / num_kv_heads), p.group_m, p.group_s);
  CUtensorMap tq, tk, tv, to;
  if ((rc = get_tmap(&tq, q, static_cast<uint64_t>(total_tokens),
                     static_cast<uint64_t>(num_q_heads) * head_dim, static_cast<uint64_t>(q_stride), kPfM)))
    return rc;
  // the output as a TMA-store target: same geometry as Q (rows beyond total_tokens are clipped)
  if ((rc = get_tmap(&to, out, static_cast<uint64_t>(total_tokens),
                     static_cast<uint64_t>(num_q_heads) * head_dim, static_cast<uint64_t>(out_stride), kPfM)))
    return rc;
  // K/V pools as [slots, Hk*D] tensors, box 128 rows x 64 columns (the row extent only bounds the
  // coordinates; they are slot ids taken from req_to_token)
  static const bool tma_on = []() { const char* e = getenv("SPD_ATTN_TMA"); return !(e && e[0] == '0'); }();
  p.use_tma = tma_on ? 1 : 0;
  const uint64_t row_elems = static_cast<uint64_t>(num_kv_heads) * head_dim;
  if ((rc = get_tmap(&tk, k_pool, 1ull << 31, row_elems, row_elems, kPfN))) return rc;
  if ((rc = get_tmap(&tv, v_pool, 1ull << 31, row_elems, row_elems, kPfN))) return rc;
  const int sms = sm_budget > 0 ? sm_budget : device_sm_count_cached();
  const int grid = p.total_items < sms ? p.total_items : sms;
  // SPD_ATTN_PF_UNIT = 64 | 128: width of a softmax / MMA unit (see the kernel's kU)
  static const int unit = []() { const char* e = getenv("SPD_ATTN_PF_UNIT"); return (e && atoi(e) == 128) ? 128 : 64; }();
#define SPD_PF_LAUNCH(U, TR)                                                                           \
  do {                                                                                                 \
    auto kern = attn_prefill_kernel<T, U, TR>;                                                         \
    static bool configured = false;                                                                    \
    if (!configured) {                                                                                 \
      SPD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPfSmem));      \
      configured = true;                                                                               \
    }                                                                                                  \
    SPD_CUDA(launch(kern, dim3(grid), dim3(kPfThreads), kPfSmem, as_stream(stream), pdl_enabled(), tq, tk, tv, to, p)); \
  } while (0)
  SPD_DISPATCH_DTYPE(dtype, {
    if (g_pf_trace) {
      p.trace = g_pf_trace;
      p.trace_cap = g_pf_trace_cap;
      if (unit == 64) SPD_PF_LAUNCH(64, true);
      else SPD_PF_LAUNCH(128, true);
    } else if (unit == 64) {
      SPD_PF_LAUNCH(64, false);
    } else {
      SPD_PF_LAUNCH(128, false);
    }
  });
#undef SPD_PF_LAUNCH
  return 0;
}
