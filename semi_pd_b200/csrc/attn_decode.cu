// Batched paged-decode attention on tcgen05 — a PERSISTENT, HBM-streaming kernel.
// One query row per request over its seq_len cached rows (torch_native_backend.py:111-179);
// replaces flashinfer's BatchDecodeWithPagedKVCacheWrapper + its host-side plan()
// (S/layers/attention/flashinfer_backend.py:454-488,627-674).
//
// Swap-AB formulation (tcgen05's minimum M is 64, a decode (request, kv-head) pair has only
// G = Hq/Hk <= 8 query rows):
//     S^T[128 kv x 16] = K_tile[128 kv x D] * Q_g^T        (kv rows feed UMMA-M, heads feed N)
//     O^T[D x 16]     += V_tile^T[D x 128 kv] * P^T         (V is the MN-major A operand)
// so every K/V byte is read from HBM exactly once, straight into swizzled smem, and consumed by
// the tensor core; the 16-wide N makes the MMAs ~60 cycles per 64 KB tile — the kernel's only
// real cost is the stream itself.
//
// grid = #SMs (or the partition's SM budget); each CTA walks work items (request, kv head,
// kv split) with a static stride and keeps its TMEM, barriers and the 3-stage K/V ring alive
// across items, so the gather of the next item overlaps the epilogue of the current one.
// Splits (only used when request x kv-head pairs do not fill the SMs) are merged by the
// last-arriving CTA of a pair — no host plan, no second launch, CUDA-graph safe.
//
// Softmax without per-tile reductions: every thread owns one kv token of the tile (a TMEM lane of
// S^T).  All threads exponentiate against a shared REFERENCE max m_ref that is only raised when
// some score exceeds it by more than 2^8 (one barrier-with-OR per tile detects that); row sums are
// kept per thread and reduced once per item.  Raising m_ref (always on an item's first tile,
// rarely afterwards) takes the slow path: true tile max via shuffles, O^T rescaled in TMEM.
//
// warps 0-3: softmax, O rescale, epilogue     warps 4-5 / 6-7: K / V loaders (separate 3-slot
// rings) + Q staging     warp 8: MMA issuer + TMEM allocator.
//
// K/V staging.  The pool contract is token-granular (page_size 1), but the run allocator
// (semi_pd_b200/memory_pool.py) places positions [k*128, k*128+128) of a request in 128 CONSECUTIVE
// slots.  A tile of one kv head is then a strided 2-D box of the pool — 128 rows x 256 B, row
// stride Hk*D*2 B — and the loader fetches it with TWO TMA tensor copies (cp.async.bulk.tensor,
// 128-byte swizzle, 16 KB each) issued by one thread, completion counted on the slot's mbarrier.
// The loader verifies on the device, per tile, that the 128 page-table entries really are
// consecutive (and the tile is full); anything else — a foreign req_to_token, the reference
// allocator's scattered slots, the ragged last tile of a sequence — takes the gather path:
// 16-byte cp.async per row chunk through req_to_token into the same swizzled layout.
// (tile::gather4, 4 rows x 128 B per instruction, was built and measured in round 1: correct,
// but ~1 instruction per 27 clk per SM -> 36 GB/s per SM, below the cp.async gather's 47;
// profiles/r01_decode_attention_loader_experiments.md.)
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace spd {

constexpr int kDcThreads = 288;
constexpr int kDcTile = 128;   // kv tokens per tile
constexpr int kDcD = 128;      // head dim
constexpr int kDcN = 16;       // UMMA N (>= GQA group)
constexpr int kDcStages = 3;
constexpr int kDcTileBytes = 128 * 128 * 2;
constexpr int kDcQBytes = 16 * 128 * 2;   // [16 rows x 128 d] as two 2 KB swizzle blocks
constexpr int kDcLenCache = 1024;         // seq_lens cached in smem
constexpr int kDcQBufs = 4;  // Q staging buffers: item i uses i % 4 (see the loader)
constexpr int kDcMisc = 1024 + kDcLenCache * 4;
constexpr int kDcSmem = kDcStages * 2 * kDcTileBytes + (kDcQBufs + 2) * kDcQBytes + 1024 + kDcMisc;
constexpr int kDcMaxSplits = 32;
constexpr int kDcMaxGroup = 8;
constexpr float kDcRescale = 8.0f;  // log2 units a score may exceed m_ref before it is raised

struct DecodeParams {
  void* out;
  const void* q;
  const void* k_pool;
  const void* v_pool;
  const int32_t* req_to_token;
  const int64_t* req_pool_indices;
  const int64_t* seq_lens;
  int batch, Hq, Hk, splits;
  int64_t q_stride, out_stride, rtt_stride;
  float scale_log2;
  float* o_part;   // [b][Hq][splits][D]
  float* ml_part;  // [b][Hq][splits][2]
  int* counters;   // [b][Hk]
  int use_tma;     // 0: always gather (SPD_ATTN_TMA=0)
};

struct DcItem {
  int b, hk, split, t0, t1, n_tiles, n_ne;  // n_ne: non-empty splits of this (request, kv head)
};

__device__ __forceinline__ DcItem dc_item(const DecodeParams& p, const int* s_len, int w) {
  DcItem it;
  it.split = w % p.splits;
  const int r = w / p.splits;
  it.hk = r % p.Hk;
  it.b = r / p.Hk;
  const int L = it.b < kDcLenCache ? s_len[it.b] : static_cast<int>(p.seq_lens[it.b]);
  int per = (L + p.splits - 1) / p.splits;
  per = (per + kDcTile - 1) / kDcTile * kDcTile;
  // a split shorter than two tiles costs more in merge traffic than it saves in streaming time:
  // short sequences therefore use fewer (non-empty) splits than the launch provides
  if (p.splits > 1) per = max(per, 2 * kDcTile);
  it.n_ne = (L + per - 1) / per;
  it.t0 = it.split * per;
  it.t1 = min(L, it.t0 + per);
  it.n_tiles = it.t1 > it.t0 ? (it.t1 - it.t0 + kDcTile - 1) / kDcTile : 0;
  return it;
}

// walks the CTA's global tile sequence: (item, tile within item).  `iter` counts NON-EMPTY items
// only: it selects the Q / O double buffers, and consecutive non-empty items must alternate.
struct DcCursor {
  int w, iter, j;
  DcItem item;
  bool valid;
  __device__ __forceinline__ void init(const DecodeParams& p, const int* s_len, int total) {
    w = blockIdx.x;
    iter = 0;
    j = 0;
    settle(p, s_len, total);
  }
  __device__ __forceinline__ void settle(const DecodeParams& p, const int* s_len, int total) {
    valid = false;
    while (w < total) {
      item = dc_item(p, s_len, w);
      if (item.n_tiles > 0) { valid = true; return; }
      w += gridDim.x;  // empty items carry no tiles (the softmax warps still publish them)
    }
  }
  __device__ __forceinline__ void next(const DecodeParams& p, const int* s_len, int total) {
    if (++j >= item.n_tiles) {
      j = 0;
      w += gridDim.x;
      ++iter;
      settle(p, s_len, total);
    }
  }
};

__device__ __forceinline__ uint32_t dc_tile_off(int row, int chunk16) {
  return (chunk16 >> 3) * 16384 + row * 128 + (((chunk16 & 7) ^ (row & 7)) << 4);
}
// [16 rows x 128 cols] 16-bit, two [16 x 128 B] swizzle blocks of 2 KB
__device__ __forceinline__ uint32_t dc_small_off(int row, int col) {
  const int chunk = col >> 3;
  return (chunk >> 3) * 2048 + row * 128 + (((chunk & 7) ^ (row & 7)) << 4) + (col & 7) * 2;
}
// barrier + OR-reduction over the 128 softmax threads (named barrier 1)
__device__ __forceinline__ bool softmax_bar_or(bool pred) {
  uint32_t out;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.u32 q, %1, 0;\n\t"
      "barrier.cta.red.or.pred p, 1, 128, q;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(out)
      : "r"(static_cast<uint32_t>(pred))
      : "memory");
  return out != 0;
}
__device__ __forceinline__ void softmax_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
// barrier + AND-reduction over the 64 threads of one loader group (named barrier 2: K, 3: V)
__device__ __forceinline__ bool loader_bar_and(bool pred, int is_v) {
  uint32_t out;
  if (is_v) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.u32 q, %1, 0;\n\t"
        "barrier.cta.red.and.pred p, 3, 64, q;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(out)
        : "r"(static_cast<uint32_t>(pred))
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.u32 q, %1, 0;\n\t"
        "barrier.cta.red.and.pred p, 2, 64, q;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(out)
        : "r"(static_cast<uint32_t>(pred))
        : "memory");
  }
  return out != 0;
}

template <typename T>
__global__ void __launch_bounds__(kDcThreads, 1)
attn_decode_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const DecodeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sK = smem;                                    // 3 x 32 KB
  uint8_t* sV = smem + kDcStages * kDcTileBytes;         // 3 x 32 KB
  uint8_t* sQ = smem + 2 * kDcStages * kDcTileBytes;     // kDcQBufs x 4 KB
  uint8_t* sP = sQ + kDcQBufs * kDcQBytes;               // 2 x 4 KB
  uint8_t* misc = sP + 2 * kDcQBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);
  uint64_t* k_full = bars;         // 3
  uint64_t* k_empty = bars + 3;    // 3
  uint64_t* v_full = bars + 6;     // 3
  uint64_t* v_empty = bars + 9;    // 3
  uint64_t* s_full = bars + 12;    // 2
  uint64_t* s_empty = bars + 14;   // 2
  uint64_t* p_full = bars + 16;    // 2
  uint64_t* p_empty = bars + 18;   // 2
  uint64_t* o_done = bars + 20;    // 1
  uint64_t* o_empty = bars + 21;   // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 23);
  int* flag_slot = reinterpret_cast<int*>(tmem_slot + 1);
  float* red = reinterpret_cast<float*>(misc + 256);   // [4 warps][8 heads] scratch (+ 8 totals)
  int* s_len = reinterpret_cast<int*>(misc + 1024);    // seq_lens[0 .. kDcLenCache)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int G = p.Hq / p.Hk;
  const int total = p.batch * p.Hk * p.splits;

  if (warp == 8) {
    if (lane == 0) {
      for (int i = 0; i < kDcStages; ++i) {
        mbar_init(&k_full[i], 64);
        mbar_init(&k_empty[i], 1);
        mbar_init(&v_full[i], 64);
        mbar_init(&v_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&s_empty[i], 4);
        mbar_init(&p_full[i], 128);
        mbar_init(&p_empty[i], 1);
        mbar_init(&o_empty[i], 4);
      }
      mbar_init(o_done, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  // zero P / Q staging once: rows >= G are never written and only feed ignored output columns
  for (int i = threadIdx.x; i < ((kDcQBufs + 2) * kDcQBytes) / 16; i += kDcThreads)
    reinterpret_cast<uint4*>(sQ)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();

  // everything above overlaps the previous kernel's tail under PDL; seq_lens is its output
  // PDL: release the dependents first (they block in their own griddepcontrol.wait until THIS grid
  // has completed, and a grid completes only after its own wait — so order is transitive), then
  // wait for the predecessor.  A decode GEMM two kernels downstream can then take over an SM the
  // moment the previous GEMM's CTA leaves it and stream its weights under this kernel.
  griddep_launch_dependents();
  griddep_wait();
  for (int i = threadIdx.x; i < min(p.batch, kDcLenCache); i += kDcThreads)
    s_len[i] = static_cast<int>(p.seq_lens[i]);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;        // 2 x 16 columns
  const uint32_t tmem_O = tmem_base + 32;   // 2 x 16 columns

  if (warp == 8) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc<T>(128, kDcN, 0, 0);  // A = K (K-major), B = Q (K-major)
      constexpr uint32_t idesc_o = umma_idesc<T>(128, kDcN, 1, 0);  // A = V (MN-major), B = P^T (K-major)
      DcCursor cs, cp;  // cs runs one tile ahead of cp
      cs.init(p, s_len, total);
      cp.init(p, s_len, total);
      int ts = 0;  // global tile counter of the S stream
      auto issue_s = [&]() {
        const int st = ts % kDcStages, sb = ts & 1;
        mbar_wait(&k_full[st], (ts / kDcStages) & 1);
        fence_proxy_async_smem();  // cp.async (generic proxy) writes -> UMMA operand reads
        mbar_wait(&s_empty[sb], ((ts >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t k_addr = smem_u32(sK + st * kDcTileBytes);
        const uint32_t q_addr = smem_u32(sQ + (cs.iter % kDcQBufs) * kDcQBytes);
#pragma unroll
        for (int k = 0; k < kDcD / 16; ++k) {
          umma_ss(tmem_S + sb * kDcN,
                  umma_desc_sw128(k_addr + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                  umma_desc_sw128(q_addr + (k >> 2) * 2048 + (k & 3) * 32, 16, 1024), idesc_s,
                  k > 0 ? 1u : 0u);
        }
        umma_commit(&k_empty[st]);  // K_t: consumed once these MMAs have read it
        umma_commit(&s_full[sb]);
        ++ts;
        cs.next(p, s_len, total);
      };
      if (cs.valid) issue_s();
      int tp = 0;  // global tile counter of the PV stream
      int o_uses[2] = {0, 0};
      while (cp.valid) {
        if (cs.valid) issue_s();
        const int st = tp % kDcStages, pb = tp & 1, ob = cp.iter & 1;
        mbar_wait(&p_full[pb], (tp >> 1) & 1);
        mbar_wait(&v_full[st], (tp / kDcStages) & 1);
        fence_proxy_async_smem();
        if (cp.j == 0) {
          mbar_wait(&o_empty[ob], (o_uses[ob] & 1) ^ 1);
          ++o_uses[ob];
        }
        tc_fence_after_sync();
        const uint32_t v_addr = smem_u32(sV + st * kDcTileBytes);
        const uint32_t p_addr = smem_u32(sP + pb * kDcQBytes);
#pragma unroll
        for (int k = 0; k < kDcTile / 16; ++k) {
          umma_ss(tmem_O + ob * kDcN, umma_desc_sw128(v_addr + k * 2048, 16384, 1024),
                  umma_desc_sw128(p_addr + (k >> 2) * 2048 + (k & 3) * 32, 16, 1024), idesc_o,
                  (cp.j > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&v_empty[st]);
        umma_commit(&p_empty[pb]);
        umma_commit(o_done);
        ++tp;
        cp.next(p, s_len, total);
      }
    }
  } else if (warp >= 4) {
    // ===================== K loader (warps 4-5, + Q staging) / V loader (warps 6-7) =====================
    // Independent rings: a K slot is free again as soon as S_t has been issued, a V slot only after
    // PV_t (one softmax later), so the K gather runs ahead of the V gather instead of both being
    // held back by the slower of the two.
    const int is_v = warp >= 6 ? 1 : 0;
    const int lt = threadIdx.x - (is_v ? 192 : 128);  // 0..63
    const int chunk = lt & 15;   // 16-byte chunk of the 256-byte head row
    const int rsub = lt >> 4;    // 0..3
    const T* pool = reinterpret_cast<const T*>(is_v ? p.v_pool : p.k_pool);
    uint8_t* ring = is_v ? sV : sK;
    uint64_t* full = is_v ? v_full : k_full;
    uint64_t* empty = is_v ? v_empty : k_empty;
    const int64_t row_elems = static_cast<int64_t>(p.Hk) * kDcD;
    DcCursor c;
    c.init(p, s_len, total);
    int t = 0;
    // Page-table lookups run ONE TILE AHEAD of the gather (32 slots per thread), so the dependent
    // chain req_pool_indices -> req_to_token -> pool row is off the issue path.
    // `base` = slot of the tile's first row when the tile is FULL (else -1): the candidate TMA row.
    auto lookup = [&](const DcCursor& cc, int* sl, int& base) {
      const DcItem& it = cc.item;
      const int32_t* rtt0 = p.req_to_token + p.req_pool_indices[it.b] * p.rtt_stride;
      const int tile0 = it.t0 + cc.j * kDcTile;
      const int tokb = tile0 + rsub;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int tok = tokb + i * 4;
        sl[i] = tok < it.t1 ? __ldg(rtt0 + tok) : -1;
      }
      base = (p.use_tma && tile0 + kDcTile <= it.t1) ? __ldg(rtt0 + tile0) : -1;
    };
    const CUtensorMap* tmap = is_v ? &tmV : &tmK;
    if (lt == 0) tma_prefetch_desc(tmap);
    int slot[32], base = -1;
    if (c.valid) lookup(c, slot, base);
    while (c.valid) {
      const int st = t % kDcStages;
      const int b = c.item.b, hk = c.item.hk, cj = c.j, citer = c.iter;
      c.next(p, s_len, total);
      int nslot[32], nbase = -1;
      if (c.valid) lookup(c, nslot, nbase);
      // Are the tile's 128 page-table entries consecutive slots?  Each thread checks its 32 rows
      // against the first row's slot; the group agrees through one barrier-with-AND.
      bool contig = base > 0;
#pragma unroll
      for (int i = 0; i < 32; ++i) contig &= (slot[i] == base + i * 4 + rsub);
      if (p.use_tma) contig = loader_bar_and(contig, is_v);
      mbar_wait(&empty[st], ((t / kDcStages) & 1) ^ 1);
      if (!is_v && cj == 0) {
        // stage the G query rows of this (request, kv head) as the K-major B operand, in the same
        // cp.async group as the K tile.  Buffer citer % 4 was last read by the S MMAs of item
        // citer - 4, whose tiles are all older than tile t-3 — complete once k_empty[st] fired.
        for (int r = rsub; r < G; r += 4) {
          const T* src = reinterpret_cast<const T*>(p.q) + b * p.q_stride +
                         (static_cast<int64_t>(hk) * G + r) * kDcD + chunk * 8;
          cp_async16(smem_u32(sQ + (citer % kDcQBufs) * kDcQBytes + dc_small_off(r, chunk * 8)), src, true);
        }
      }
      if (contig) {
        // one strided 2-D box of the pool: rows base .. base+127, the 128 columns of kv head hk,
        // as two [128 x 64] boxes = the two 16 KB swizzle blocks of the slot
        if (lt == 0) {
          uint8_t* dst = ring + st * kDcTileBytes;
          fence_proxy_async_smem();  // the slot may last have been written by the gather path
          mbar_expect_tx_noarrive(&full[st], kDcTileBytes);
          tma_load_2d(dst, tmap, &full[st], hk * kDcD, base, kEvictFirst);
          tma_load_2d(dst + 16384, tmap, &full[st], hk * kDcD + 64, base, kEvictFirst);
        }
      } else {
        const int64_t col = static_cast<int64_t>(hk) * kDcD + chunk * 8;
        const uint32_t dst = smem_u32(ring + st * kDcTileBytes);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int row = i * 4 + rsub;
          const bool valid = slot[i] >= 0;
          const int64_t s64 = valid ? slot[i] : 0;
          cp_async16(dst + dc_tile_off(row, chunk), pool + s64 * row_elems + col, valid);
        }
      }
      // completion is asynchronous: the barrier receives this thread's arrival when all of its
      // cp.asyncs above have landed — the loader never waits for data, only for a free slot, so
      // every slot of the ring can be in flight.  (The generic->async proxy fence is on the
      // consumer side, after its wait on this barrier.)
      // (TMA tiles: the arrival below fires at once — or with the Q rows staged above — and the
      // phase completes when the 32 KB the issuing thread announced have landed as well.)
      cp_async_mbar_arrive_noinc(&full[st]);
      ++t;
      base = nbase;
#pragma unroll
      for (int i = 0; i < 32; ++i) slot[i] = nslot[i];
    }
  } else {
    // ===================== softmax / correction / epilogue =====================
    const int tid = threadIdx.x;  // TMEM lane: kv token within tile (S^T) or head-dim index (O^T)
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    int t = 0;     // global tile counter
    int iter = 0;  // non-empty items seen so far (same convention as DcCursor::iter)
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const DcItem it = dc_item(p, s_len, w);
      float m_ref[kDcMaxGroup], l_thr[kDcMaxGroup];
#pragma unroll
      for (int g = 0; g < kDcMaxGroup; ++g) { m_ref[g] = -INFINITY; l_thr[g] = 0.f; }
      const int ob = iter & 1;
      for (int j = 0; j < it.n_tiles; ++j, ++t) {
        const int sb = t & 1;
        mbar_wait(&s_full[sb], (t >> 1) & 1);
        tc_fence_after_sync();
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem_S + lane_base + sb * kDcN, r);
        tmem_ld_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[sb]);
        const bool valid = (it.t0 + j * kDcTile + tid) < it.t1;
        float s[kDcMaxGroup];
        bool exceed = false;
#pragma unroll
        for (int g = 0; g < kDcMaxGroup; ++g) {
          s[g] = valid ? __uint_as_float(r[g]) * p.scale_log2 : -INFINITY;
          exceed |= (g < G) && (s[g] > m_ref[g] + kDcRescale);
        }
        // one barrier per tile: does ANY score need a higher reference max?
        if (softmax_bar_or(exceed)) {
          // slow path (first tile of an item, rarely later): true tile max, raise m_ref, rescale
#pragma unroll
          for (int g = 0; g < kDcMaxGroup; ++g) {
            const float wm = warp_max(s[g]);
            if (lane == 0) red[warp * 8 + g] = wm;
          }
          softmax_bar();
          float alpha[kDcMaxGroup];
#pragma unroll
          for (int g = 0; g < kDcMaxGroup; ++g) {
            const float mt = fmaxf(fmaxf(red[g], red[8 + g]), fmaxf(red[16 + g], red[24 + g]));
            const float m_new = fmaxf(m_ref[g], mt);  // finite: every tile has >= 1 valid token
            alpha[g] = (m_ref[g] == -INFINITY) ? 0.f : fast_exp2(m_ref[g] - m_new);
            l_thr[g] *= alpha[g];
            m_ref[g] = m_new;
          }
          if (j > 0) {
            // O^T stable once PV_{t-1} is complete.  That completion is observed on p_empty (committed
            // with every PV): this thread waited for the PREVIOUS phase of that barrier at tile t-1,
            // so the parity wait cannot alias — a per-PV barrier that is only waited on now and then
            // (the softmax may run up to two PVs ahead) would.
            mbar_wait(&p_empty[(t - 1) & 1], ((t - 1) >> 1) & 1);
            tc_fence_after_sync();
            uint32_t o[16];
            tmem_ld_32x32b_x16(tmem_O + lane_base + ob * kDcN, o);
            tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < kDcMaxGroup; ++g)
              o[g] = __float_as_uint(__uint_as_float(o[g]) * alpha[g]);
            tmem_st_32x32b_x16(tmem_O + lane_base + ob * kDcN, o);
            tmem_st_wait();
            tc_fence_before_sync();
          }
          softmax_bar();  // `red` may be reused by the next slow path
        }
        // P^T -> smem (B operand, K-major: row = head, col = kv token); row sums stay per thread
        mbar_wait(&p_empty[sb], ((t >> 1) & 1) ^ 1);
        uint8_t* pdst = sP + sb * kDcQBytes;
#pragma unroll
        for (int g = 0; g < kDcMaxGroup; ++g) {
          if (g < G) {
            const T pt = from_f32<T>(fast_exp2(s[g] - m_ref[g]));  // <= 2^8; 0 for invalid rows
            l_thr[g] += to_f32<T>(pt);  // sum what the tensor core multiplies
            *reinterpret_cast<T*>(pdst + dc_small_off(g, tid)) = pt;
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(&p_full[sb]);  // 128 arrivals
      }
      // ---- epilogue of this item: reduce the per-thread row sums, then O^T / l ----
#pragma unroll
      for (int g = 0; g < kDcMaxGroup; ++g) {
        const float ws = warp_sum(l_thr[g]);
        if (lane == 0) red[warp * 8 + g] = ws;
      }
      softmax_bar();
      float l_tot[kDcMaxGroup], o[kDcMaxGroup];
#pragma unroll
      for (int g = 0; g < kDcMaxGroup; ++g)
        l_tot[g] = (red[g] + red[8 + g]) + (red[16 + g] + red[24 + g]);
      if (it.n_tiles > 0) {
        mbar_wait(&p_empty[(t - 1) & 1], ((t - 1) >> 1) & 1);  // last PV of the item complete
        tc_fence_after_sync();
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem_O + lane_base + ob * kDcN, r);
        tmem_ld_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[ob]);
#pragma unroll
        for (int g = 0; g < kDcMaxGroup; ++g) o[g] = __uint_as_float(r[g]);
        ++iter;
      } else {
#pragma unroll
        for (int g = 0; g < kDcMaxGroup; ++g) o[g] = 0.f;
      }
      if (p.splits == 1) {
        T* dst = reinterpret_cast<T*>(p.out) + it.b * p.out_stride +
                 static_cast<int64_t>(it.hk) * G * kDcD + tid;
#pragma unroll
        for (int g = 0; g < kDcMaxGroup; ++g)
          if (g < G) dst[g * kDcD] = from_f32<T>(l_tot[g] > 0.f ? o[g] / l_tot[g] : 0.f);
        softmax_bar();  // `red` is rewritten by the next item
      } else {
#pragma unroll
        for (int g = 0; g < kDcMaxGroup; ++g)
          if (g < G && it.n_tiles > 0) {
            const int64_t idx = (static_cast<int64_t>(it.b) * p.Hq + it.hk * G + g) * p.splits + it.split;
            p.o_part[idx * kDcD + tid] = o[g];
            if (tid == 0) { p.ml_part[idx * 2] = m_ref[g]; p.ml_part[idx * 2 + 1] = l_tot[g]; }
          }
        __threadfence();
        softmax_bar();
        if (tid == 0) *flag_slot = atomicAdd(&p.counters[it.b * p.Hk + it.hk], 1);
        softmax_bar();
        if (*flag_slot == p.splits - 1) {
          // last split of this (request, kv head): merge the non-empty partials in split order,
          // all heads of the group at once so the loads of one split are independent
          __threadfence();
          const int64_t base0 = (static_cast<int64_t>(it.b) * p.Hq + it.hk * G) * p.splits;
          float M[kDcMaxGroup], num[kDcMaxGroup], den[kDcMaxGroup];
#pragma unroll
          for (int g = 0; g < kDcMaxGroup; ++g) { M[g] = -INFINITY; num[g] = 0.f; den[g] = 0.f; }
#pragma unroll 2
          for (int s2 = 0; s2 < it.n_ne; ++s2)
#pragma unroll
            for (int g = 0; g < kDcMaxGroup; ++g)
              if (g < G) M[g] = fmaxf(M[g], __ldcg(&p.ml_part[(base0 + g * p.splits + s2) * 2]));
#pragma unroll 2
          for (int s2 = 0; s2 < it.n_ne; ++s2) {
#pragma unroll
            for (int g = 0; g < kDcMaxGroup; ++g)
              if (g < G) {
                const int64_t idx = base0 + g * p.splits + s2;
                const float2 ml = __ldcg(reinterpret_cast<const float2*>(&p.ml_part[idx * 2]));
                const float ov = __ldcg(&p.o_part[idx * kDcD + tid]);
                const float wgt = ml.y > 0.f ? fast_exp2(ml.x - M[g]) : 0.f;
                num[g] += wgt * ov;
                den[g] += wgt * ml.y;
              }
          }
          T* dst = reinterpret_cast<T*>(p.out) + it.b * p.out_stride +
                   static_cast<int64_t>(it.hk) * G * kDcD + tid;
#pragma unroll
          for (int g = 0; g < kDcMaxGroup; ++g)
            if (g < G) dst[g * kDcD] = from_f32<T>(den[g] > 0.f ? num[g] / den[g] : 0.f);
          if (tid == 0) p.counters[it.b * p.Hk + it.hk] = 0;  // re-arm for the next launch
        }
        softmax_bar();  // flag_slot / red are reused by the next item
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 64);
}

int check_attn_args(const char* fn, const void* out, const void* q, const void* k_pool,
                    const void* v_pool, const void* rtt, const void* rpi, int Hq, int Hk, int D,
                    int64_t q_stride, int64_t out_stride);

}  // namespace spd

using namespace spd;

extern "C" int spd_attn_decode_paged(void* out, const void* q, const void* k_pool,
                                     const void* v_pool, const int32_t* req_to_token,
                                     const int64_t* req_pool_indices, const int64_t* seq_lens,
                                     int64_t batch, int num_q_heads, int num_kv_heads, int head_dim,
                                     int64_t q_stride, int64_t out_stride,
                                     int64_t req_to_token_stride, float sm_scale, int dtype,
                                     void* workspace, int64_t workspace_bytes, int sm_budget,
                                     spd_stream_t stream) {
  int rc = check_attn_args("spd_attn_decode_paged", out, q, k_pool, v_pool, req_to_token,
                           req_pool_indices, num_q_heads, num_kv_heads, head_dim, q_stride, out_stride);
  if (rc) return rc;
  SPD_REQUIRE(seq_lens && batch >= 0, SPD_E_BADARG, "spd_attn_decode_paged: bad seq_lens/batch");
  SPD_REQUIRE(head_dim == kDcD, SPD_E_SHAPE,
              "spd_attn_decode_paged: head_dim must be 128 on the tcgen05 path (got %d)", head_dim);
  if (batch == 0) return 0;
  const int sms = sm_budget > 0 ? sm_budget : device_sm_count_cached();
  DecodeParams p{};
  p.out = out;
  p.q = q;
  p.k_pool = k_pool;
  p.v_pool = v_pool;
  p.req_to_token = req_to_token;
  p.req_pool_indices = req_pool_indices;
  p.seq_lens = seq_lens;
  p.batch = static_cast<int>(batch);
  p.Hq = num_q_heads;
  p.Hk = num_kv_heads;
  p.q_stride = q_stride;
  p.out_stride = out_stride;
  p.rtt_stride = req_to_token_stride;
  p.scale_log2 = sm_scale * 1.4426950408889634f;
  // KV splits only when the (request, kv head) pairs cannot fill the SMs; the choice depends on
  // the launch shape alone (lengths are device-side) — nothing to plan on the host per step.
  const int64_t pairs = batch * num_kv_heads;
  int splits = 1;
  if (pairs < sms) {
    splits = static_cast<int>((sms + pairs - 1) / pairs);
    if (splits > kDcMaxSplits) splits = kDcMaxSplits;
  }
  p.splits = splits;
  if (splits > 1) {
    SPD_REQUIRE(workspace && workspace_bytes >= spd_attn_decode_workspace_bytes(batch, num_q_heads, head_dim),
                SPD_E_WORKSPACE, "spd_attn_decode_paged: workspace too small (%lld bytes)",
                (long long)workspace_bytes);
    SPD_REQUIRE(pairs <= 16384, SPD_E_SHAPE, "spd_attn_decode_paged: batch x kv heads exceeds 16384");
    // layout (attn_simt.cu: spd_attn_decode_workspace_bytes): counters at a fixed offset first
    p.counters = reinterpret_cast<int*>(workspace);
    p.o_part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + 65536);
    p.ml_part = p.o_part + batch * num_q_heads * kDcMaxSplits * head_dim;
  }
  // tensor maps over the pools: [slots, Hk*D] 16-bit, box 128 rows x 64 columns.  The C-ABI does
  // not carry the pool's slot count; the row extent only bounds the coordinates the kernel may
  // pass, and those are slot ids read from req_to_token, so the maximum is declared.
  static const bool tma_on = []() { const char* e = getenv("SPD_ATTN_TMA"); return !(e && e[0] == '0'); }();
  p.use_tma = tma_on ? 1 : 0;
  CUtensorMap tk, tv;
  const uint64_t row_elems = static_cast<uint64_t>(num_kv_heads) * head_dim;
  if ((rc = get_tmap(&tk, k_pool, 1ull << 31, row_elems, row_elems, kDcTile))) return rc;
  if ((rc = get_tmap(&tv, v_pool, 1ull << 31, row_elems, row_elems, kDcTile))) return rc;
  const int64_t total = pairs * splits;
  const int grid = static_cast<int>(total < sms ? total : sms);
  SPD_DISPATCH_DTYPE(dtype, {
    auto kern = attn_decode_kernel<T>;
    static bool configured = false;
    if (!configured) {
      SPD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kDcSmem));
      configured = true;
    }
    SPD_CUDA(launch(kern, dim3(grid), dim3(kDcThreads), kDcSmem, as_stream(stream), pdl_enabled(), tk, tv, p));
  });
  return 0;
}
