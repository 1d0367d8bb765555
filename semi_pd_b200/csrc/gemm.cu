// C[M,N] = A[M,K] * W[N,K]^T on the 5th-generation tensor cores (tcgen05.mma, accumulators in
// TMEM), operands staged by TMA into 128-byte-swizzled shared memory.  Replaces F.linear /
// cuBLAS for W_qkv, W_o, W_gate_up, W_down and lm_head (S/layers/linear.py:165-172,
// S/layers/logits_processor.py:416-419).
//
// One kernel, two regimes — both operands are K-major, so the same main loop serves either:
//   * prefill regime (M > 256): D[128 x BN] tile = A_tile * W_tile^T, persistent CTAs walk the
//     tile list M-fastest so the W tile of a column block is shared through L2.
//   * decode regime (M <= 256), "swap-AB": D[128 x BN] = W_tile[128 rows of N] * A[BN >= M]^T,
//     i.e. the weight dimension feeds UMMA-M (128 TMEM lanes) and the few tokens feed UMMA-N,
//     so streaming the weights at HBM speed is the only cost.  Work is scheduled STREAM-K style:
//     the (tile x k-block) unit space is cut into one equal contiguous range per CTA, so every SM
//     streams the same number of weight bytes whatever N is (32 tiles or 1002) and whatever SM
//     budget the caller's partition has.  A tile shared by several CTAs is finished by its OWNER,
//     the CTA holding the tile's last k-blocks: that segment is always the owner's first, so its
//     accumulator stays parked in TMEM while the CTA works through the rest of its range; the
//     other contributors dump fp32 partials (coalesced, L2-resident) as their last piece of work
//     and raise a flag; at the very end the owner adds the partials in ascending-k order to its
//     parked accumulator and stores C.  Nobody waits before having published its own partial
//     (no deadlock: the grid never exceeds the caller's SM budget, so all CTAs are co-resident on
//     a green-context partition; and an owner only ever waits for LOWER-indexed CTAs, which the
//     block scheduler dispatches first, so even on SMs shared with another kernel the CTAs it
//     waits for are running or done), the wait is short because every CTA has the same amount of
//     work, and the sum order is fixed (bit-deterministic, no atomics on data).
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane),
// warp 2 = TMEM allocator, warps 4-7 = epilogue (TMEM -> registers -> global).
// Pipelines: smem ring full/empty mbarriers (TMA <-> MMA), TMEM double buffer full/empty
// (MMA <-> epilogue).
#include <cuda.h>
#include <stdlib.h>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace spd {

constexpr int kGemmThreads = 256;
constexpr int kBM = 128;  // UMMA M (TMEM lanes)
constexpr int kBK = 64;   // one 128-byte swizzle atom of 16-bit elements
constexpr int kMaxStages = 12;
constexpr int kSmemBudget = 220 * 1024;

struct GemmParams {
  void* C;
  int64_t ldc;
  int64_t x_rows;    // rows of the 128-lane operand (M normal, N swapped)
  int64_t y_rows;    // rows of the BN operand
  int64_t y_rows_w;  // SILU_MUL: offset (rows) of the "up" half inside W
  int x_tiles, y_tiles, kblocks;
  int streamk;       // 0: whole tiles, strided over CTAs; 1: equal unit ranges (decode regime)
  int panel;         // tile rasterisation: y tiles per panel (see tile_coords)
  int out_f32;
  float* ws;        // stream-K partials: one [col/4][lane][4] fp32 tile per CTA
  int* counters;    // stream-K: one 'partial published' flag per CTA
  // grouped forms (spd_gemm_grouped; normal regime, whole tiles only)
  const int* x_tile_group;  // rows mode: group (expert) of every 128-row x tile, < 0 = skip the tile;
                            // the Y (weight) operand's rows start at group * y_group_rows
  int64_t y_group_rows;
  int64_t y_group_cols;     // rows mode, K-concatenated weights: the Y operand's COLUMNS start at group * y_group_cols
  int y_tiles_per_group;    // cols mode: group = y tile / y_tiles_per_group; the X (activation)
  int64_t x_group_cols;     // operand's columns start at group * x_group_cols,
  int64_t c_group_cols;     // and the group's output columns at group * c_group_cols (0: contiguous)
};

// (x column offset, y row offset) of a tile under the grouped forms; false = skip the tile
__device__ __forceinline__ bool group_offsets(const GemmParams& p, int xt, int yt, int& xcol0, int& yrow0,
                                              int& ycol0) {
  xcol0 = 0;
  yrow0 = 0;
  ycol0 = 0;
  if (p.x_tile_group) {
    const int g = p.x_tile_group[xt];
    if (g < 0) return false;
    yrow0 = static_cast<int>(g * p.y_group_rows);
    ycol0 = static_cast<int>(g * p.y_group_cols);
  }
  if (p.y_tiles_per_group > 0) xcol0 = static_cast<int>((yt / p.y_tiles_per_group) * p.x_group_cols);
  return true;
}

template <int BN>
__host__ __device__ constexpr int gemm_stage_bytes() { return (kBM + BN) * kBK * 2; }
template <int BN>
__host__ __device__ constexpr int gemm_stages() {
  return (kSmemBudget / gemm_stage_bytes<BN>()) < kMaxStages ? (kSmemBudget / gemm_stage_bytes<BN>())
                                                             : kMaxStages;
}
template <int BN>
__host__ __device__ constexpr int gemm_smem_bytes() {
  return gemm_stages<BN>() * gemm_stage_bytes<BN>() + 1024 /*align*/ + 256 /*barriers*/;
}

// Tile index -> (x tile, y tile).  Tiles are walked panel by panel: a panel is `panel` consecutive
// y tiles (rows of W in the prefill regime) times ALL x tiles, x outer / y inner.  The ~150
// concurrently running CTAs then share a W panel sized to stay in L2 (~40 MB) while the A tiles
// stream through once per panel; with panel = 1 (every A tile re-read for every W tile) ncu showed
// 5.5x the algorithmic DRAM reads for the 8192-token W_gate_up GEMM.
__device__ __forceinline__ void tile_coords(const GemmParams& p, int tile, int& xt, int& yt) {
  if (p.panel <= 1) {
    xt = tile % p.x_tiles;
    yt = tile / p.x_tiles;
    return;
  }
  const int per_panel = p.x_tiles * p.panel;
  const int pn = tile / per_panel;
  const int r = tile - pn * per_panel;
  const int y0 = pn * p.panel;
  const int q = min(p.panel, p.y_tiles - y0);
  xt = r / q;
  yt = y0 + r - xt * q;
}

// A CTA's work list: segments (tile, [kb0, kb1)).  All three warp roles walk it identically.
struct SegIter {
  int kblocks, num_tiles, streamk;
  long long u, u_end;  // stream-K: current unit, end of this CTA's range
  int w;               // whole-tile mode: next tile
  int nseg;            // segments handed out so far
  __device__ __forceinline__ static long long range_start(long long c, long long units, long long g) {
    return c * units / g;
  }
  __device__ __forceinline__ void init(const GemmParams& p) {
    kblocks = p.kblocks;
    num_tiles = p.x_tiles * p.y_tiles;
    streamk = p.streamk;
    nseg = 0;
    if (streamk) {
      const long long units = static_cast<long long>(num_tiles) * kblocks;
      u = range_start(blockIdx.x, units, gridDim.x);
      u_end = range_start(blockIdx.x + 1, units, gridDim.x);
    } else {
      w = blockIdx.x;
    }
  }
  // slot: which of the CTA's two partial buffers a partial segment uses (0 = first segment)
  __device__ __forceinline__ bool next(int& tile, int& kb0, int& kb1, int& slot) {
    if (streamk) {
      if (u >= u_end) return false;
      tile = static_cast<int>(u / kblocks);
      kb0 = static_cast<int>(u - static_cast<long long>(tile) * kblocks);
      const long long left = u_end - u;
      kb1 = (kblocks - kb0) < left ? kblocks : kb0 + static_cast<int>(left);
      u += kb1 - kb0;
    } else {
      if (w >= num_tiles) return false;
      tile = w;
      kb0 = 0;
      kb1 = kblocks;
      w += gridDim.x;
    }
    slot = nseg == 0 ? 0 : 1;
    ++nseg;
    return true;
  }
};

// kSwap: D[i][j] -> C[j][n0+i]   else D[i][j] -> C[m0+i][n0+j]
// kEpi == SPD_EPI_SILU_MUL (normal mode only): Y tile = [BN/2 gate rows | BN/2 up rows]
template <typename T, int BN, bool kSwap, int kEpi>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
            const GemmParams p) {
  constexpr int kStages = gemm_stages<BN>();
  constexpr int kXBytes = kBM * kBK * 2;
  constexpr int kYBytes = BN * kBK * 2;
  constexpr int kTmemCols = (2 * BN) <= 32 ? 32 : (2 * BN) <= 64 ? 64 : (2 * BN) <= 128 ? 128 : (2 * BN) <= 256 ? 256 : 512;
  constexpr int kBNOut = (kEpi == SPD_EPI_SILU_MUL) ? BN / 2 : BN;  // output columns per tile
  constexpr int kChunk = BN >= 32 ? 32 : 16;
  static_assert(!(kSwap && kEpi == SPD_EPI_SILU_MUL), "fused SiLU*mul only in the non-swapped regime");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sX = smem;
  uint8_t* sY = smem + kStages * kXBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * (kXBytes + kYBytes));
  uint64_t* full = bars;                     // [kStages]
  uint64_t* empty = bars + kStages;          // [kStages]
  uint64_t* tfull = bars + 2 * kStages;      // [2]
  uint64_t* tempty = bars + 2 * kStages + 2; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // everything above overlaps the previous kernel's tail under PDL.  In the decode regime
  // (swap-AB: X = the weights) the producer also fills the ring with WEIGHT tiles before it waits
  // for the predecessor — see gemm_pair_decode.cu — and the dependents are released at once.
  if constexpr (kSwap) griddep_launch_dependents();
  if (!(kSwap && warp == 0)) griddep_wait();

  if (warp == 0) {
    if (elect_one()) {
      // operand cache policy: in the decode regime the weights (X) stream through once
      const uint64_t polX = kSwap ? kEvictFirst : kEvictNormal;
      const uint64_t polY = kSwap ? kEvictLast : kEvictNormal;
      int stage = 0, phase = 0;
      int ahead = 0;
      if constexpr (kSwap) {
        SegIter pre;
        pre.init(p);
        int tile, kb0, kb1, slot;
        while (ahead < kStages && pre.next(tile, kb0, kb1, slot)) {
          int xt, yt;
          tile_coords(p, tile, xt, yt);
          for (int kb = kb0; kb < kb1 && ahead < kStages; ++kb, ++ahead) {
            mbar_expect_tx(&full[ahead], kXBytes + kYBytes);
            tma_load_2d(sX + ahead * kXBytes, &tmX, &full[ahead], kb * kBK, xt * kBM, polX);
          }
        }
        griddep_wait();
      }
      int n = 0;
      SegIter it;
      it.init(p);
      int tile, kb0, kb1, slot;
      while (it.next(tile, kb0, kb1, slot)) {
        int xt, yt, xcol0, yrow0, ycol0;
        tile_coords(p, tile, xt, yt);
        if (!group_offsets(p, xt, yt, xcol0, yrow0, ycol0)) continue;
        for (int kb = kb0; kb < kb1; ++kb, ++n) {
          if (n < ahead) {  // the weight tile of this stage is already in flight (or landed)
            tma_load_2d(sY + stage * kYBytes, &tmY, &full[stage], ycol0 + kb * kBK, yrow0 + yt * BN, polY);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], kXBytes + kYBytes);
          tma_load_2d(sX + stage * kXBytes, &tmX, &full[stage], xcol0 + kb * kBK, xt * kBM, polX);
          if constexpr (kEpi == SPD_EPI_SILU_MUL) {
            tma_load_2d(sY + stage * kYBytes, &tmY, &full[stage], ycol0 + kb * kBK, yrow0 + yt * kBNOut, polY);
            tma_load_2d(sY + stage * kYBytes + kYBytes / 2, &tmY, &full[stage], ycol0 + kb * kBK,
                        yrow0 + static_cast<int>(p.y_rows_w) + yt * kBNOut, polY);
          } else {
            tma_load_2d(sY + stage * kYBytes, &tmY, &full[stage], ycol0 + kb * kBK, yrow0 + yt * BN, polY);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc<T>(kBM, BN, 0, 0);
      int stage = 0, phase = 0, n = 0;
      SegIter it;
      it.init(p);
      int tile, kb0, kb1, slot;
      while (it.next(tile, kb0, kb1, slot)) {
        if (p.x_tile_group) {  // grouped rows mode: all roles skip the same (unused) tiles
          int xt, yt, a, b, c;
          tile_coords(p, tile, xt, yt);
          if (!group_offsets(p, xt, yt, a, b, c)) continue;
        }
        // stream-K: slot 0 is reserved for the first segment (parked until the end if this CTA
        // owns that tile), every later segment goes through slot 1
        const int as = p.streamk ? (n == 0 ? 0 : 1) : (n & 1);
        const int use = p.streamk ? (n == 0 ? 0 : n - 1) : (n >> 1);
        ++n;
        mbar_wait(&tempty[as], (use & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after_sync();
          const uint32_t xa = smem_u32(sX + stage * kXBytes);
          const uint32_t ya = smem_u32(sY + stage * kYBytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t ad = umma_desc_sw128(xa + k * 32, 16, 1024);
            const uint64_t bd = umma_desc_sw128(ya + k * 32, 16, 1024);
            umma_ss(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);  // smem slot free once these MMAs have read it
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[as]);  // accumulator complete
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;     // D row == TMEM lane
    const int et = threadIdx.x - 128;  // 0..127 among the epilogue threads
    const long long units = static_cast<long long>(p.x_tiles) * p.y_tiles * p.kblocks;
    // columns that carry data (swap-AB: tokens), rounded up to the 32-column chunk
    const int n_cols = kSwap ? static_cast<int>(((p.y_rows + kChunk - 1) / kChunk) * kChunk) : BN;
    const int cols_used = n_cols < BN ? n_cols : BN;

    // v[0..kChunk) of output chunk c of tile (xt, yt) -> C
    auto store_chunk = [&](int xt, int yt, int c, const float* v) {
      if constexpr (!kSwap) {
        const int64_t m = static_cast<int64_t>(xt) * kBM + row;
        const int64_t nl = static_cast<int64_t>(yt) * kBNOut + c;  // logical column: bounds; n0: where it lands
        int64_t n0 = nl;
        if (p.c_group_cols > 0) {  // cols mode with strided output groups
          const int g = yt / p.y_tiles_per_group;
          n0 = g * p.c_group_cols + static_cast<int64_t>(yt - g * p.y_tiles_per_group) * kBNOut + c;
        }
        if (m < p.x_rows) {
          if (p.out_f32) {
            float* dst = reinterpret_cast<float*>(p.C) + m * p.ldc + n0;
#pragma unroll
            for (int j = 0; j < kChunk; j += 4) {
              if (nl + j < p.y_rows) {
                float4 o;
                o.x = to_f32<T>(from_f32<T>(v[j]));
                o.y = to_f32<T>(from_f32<T>(v[j + 1]));
                o.z = to_f32<T>(from_f32<T>(v[j + 2]));
                o.w = to_f32<T>(from_f32<T>(v[j + 3]));
                *reinterpret_cast<float4*>(dst + j) = o;
              }
            }
          } else {
            T* dst = reinterpret_cast<T*>(p.C) + m * p.ldc + n0;
#pragma unroll
            for (int j = 0; j < kChunk; j += 8) {
              if (nl + j < p.y_rows) {
                uint4 o;
                o.x = pack2<T>(v[j], v[j + 1]);
                o.y = pack2<T>(v[j + 2], v[j + 3]);
                o.z = pack2<T>(v[j + 4], v[j + 5]);
                o.w = pack2<T>(v[j + 6], v[j + 7]);
                *reinterpret_cast<uint4*>(dst + j) = o;
              }
            }
          }
        }
      } else {
        // swapped: lane index runs along N (contiguous in C), column j is the token.  The type
        // branch and the address arithmetic are hoisted: one predicated store per element (the
        // per-element form compiled to ~24 instructions per 2-byte store).
        const int64_t nn = static_cast<int64_t>(xt) * kBM + row;
        if (nn < p.x_rows) {
          const int m0 = yt * BN + c;
          const int m_left = static_cast<int>(p.y_rows) - m0;  // tokens of this chunk that exist
          if (p.out_f32) {
            float* dst = reinterpret_cast<float*>(p.C) + static_cast<int64_t>(m0) * p.ldc + nn;
#pragma unroll
            for (int j = 0; j < kChunk; ++j)
              if (j < m_left) dst[static_cast<int64_t>(j) * p.ldc] = to_f32<T>(from_f32<T>(v[j]));
          } else {
            T* dst = reinterpret_cast<T*>(p.C) + static_cast<int64_t>(m0) * p.ldc + nn;
#pragma unroll
            for (int j = 0; j < kChunk; ++j)
              if (j < m_left) dst[static_cast<int64_t>(j) * p.ldc] = from_f32<T>(v[j]);
          }
        }
      }
    };

    int n = 0;
    int own_tile = -1;  // stream-K: tile whose accumulator is parked in TMEM slot 0
    SegIter it;
    it.init(p);
    int tile, kb0, kb1, slot;
    while (it.next(tile, kb0, kb1, slot)) {
      int xt, yt;
      tile_coords(p, tile, xt, yt);
      if (p.x_tile_group && p.x_tile_group[xt] < 0) continue;
      const int as = p.streamk ? (n == 0 ? 0 : 1) : (n & 1);
      const int use = p.streamk ? (n == 0 ? 0 : n - 1) : (n >> 1);
      ++n;
      const bool complete = (kb0 == 0 && kb1 == p.kblocks);
      if (!complete && kb1 == p.kblocks) {
        // this CTA holds the tile's last k-blocks: it owns the tile.  Always the CTA's first
        // segment -> TMEM slot 0, which stays parked until the fix-up after the loop.
        own_tile = tile;
        continue;
      }
      mbar_wait(&tfull[as], use & 1);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;

      if (!complete) {
        // contributor (always this CTA's last segment): publish the fp32 partial.  Layout
        // [col/4][lane][4]: one warp store covers 512 contiguous bytes.
        float* mine = p.ws + static_cast<size_t>(blockIdx.x) * (kBM * BN) + row * 4;
#pragma unroll 1
        for (int c = 0; c < cols_used; c += kChunk) {
          uint32_t r[kChunk];
          if constexpr (kChunk == 32) tmem_ld_32x32b_x32(taddr + c, r);
          else tmem_ld_32x32b_x16(taddr + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < kChunk; j += 4)
            *reinterpret_cast<uint4*>(mine + static_cast<size_t>((c + j) >> 2) * (kBM * 4)) =
                make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[as]);
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.counters + blockIdx.x), "r"(1) : "memory");
        }
        continue;
      }

#pragma unroll 1
      for (int c = 0; c < (kSwap ? cols_used : kBNOut); c += kChunk) {
        float v[kChunk];
        uint32_t r[kChunk];
        if constexpr (kChunk == 32) tmem_ld_32x32b_x32(taddr + c, r);
        else tmem_ld_32x32b_x16(taddr + c, r);
        if constexpr (kEpi == SPD_EPI_SILU_MUL) {
          uint32_t u[kChunk];
          tmem_ld_32x32b_x32(taddr + kBNOut + c, u);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < kChunk; ++j) {
            // gate_up_proj output is rounded to T before SiluAndMul reads it
            const float g = to_f32<T>(from_f32<T>(__uint_as_float(r[j])));
            const float up = to_f32<T>(from_f32<T>(__uint_as_float(u[j])));
            v[j] = silu_f32(g) * up;
          }
        } else {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < kChunk; ++j) v[j] = __uint_as_float(r[j]);
        }
        store_chunk(xt, yt, c, v);
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
    }

    if (own_tile >= 0) {
      // ---- deferred fix-up of the owned tile: partials of the other contributors (ascending CTA
      // = ascending k) + the parked accumulator, in that fixed order ----
      int xt, yt;
      tile_coords(p, own_tile, xt, yt);
      const long long t0 = static_cast<long long>(own_tile) * p.kblocks;
      const int c_first = static_cast<int>(((t0 + 1) * gridDim.x - 1) / units);
      const int c_last = static_cast<int>(blockIdx.x);  // == ((t0 + kblocks) * grid - 1) / units
      mbar_wait(&tfull[0], 0);
      tc_fence_after_sync();
      for (int cc = c_first + et; cc < c_last; cc += 128) {
        int f;
        do {
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(p.counters + cc) : "memory");
        } while (f == 0);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < cols_used; c += kChunk) {
        float v[kChunk];
#pragma unroll
        for (int j = 0; j < kChunk; ++j) v[j] = 0.f;
        // the partials of up to 4 contributors are requested together (independent L2 round
        // trips), then added in ascending-k order
        for (int cc0 = c_first; cc0 < c_last; cc0 += 4) {
          float4 f[4][kChunk / 4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (cc0 + i < c_last) {
              const float* ps = p.ws + static_cast<size_t>(cc0 + i) * (kBM * BN) + row * 4;
#pragma unroll
              for (int j = 0; j < kChunk; j += 4)
                f[i][j >> 2] = __ldcg(reinterpret_cast<const float4*>(ps + static_cast<size_t>((c + j) >> 2) * (kBM * 4)));
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (cc0 + i < c_last) {
#pragma unroll
              for (int j = 0; j < kChunk; j += 4) {
                v[j] += f[i][j >> 2].x; v[j + 1] += f[i][j >> 2].y; v[j + 2] += f[i][j >> 2].z; v[j + 3] += f[i][j >> 2].w;
              }
            }
          }
        }
        uint32_t r[kChunk];
        if constexpr (kChunk == 32) tmem_ld_32x32b_x32(taddr + c, r);
        else tmem_ld_32x32b_x16(taddr + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < kChunk; ++j) v[j] += __uint_as_float(r[j]);
        store_chunk(xt, yt, c, v);
      }
      // re-arm the flags for the next launch / graph replay (each flag has exactly one reader)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int cc = c_first + et; cc < c_last; cc += 128) p.counters[cc] = 0;
    }
  }

  if constexpr (!kSwap) griddep_launch_dependents();
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct TmapKey {
  const void* base;
  uint64_t rows, cols, stride;
  uint32_t box_rows;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && stride == o.stride &&
           box_rows == o.box_rows;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.base);
    h ^= k.rows * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= k.cols * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= k.stride * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
    h ^= k.box_rows + (h << 6) + (h >> 2);
    return h;
  }
};

// Descriptors are pure functions of (pointer, shape, box); cache them so steady-state launches
// do no driver work.  Thread-safe: the prefill and decode workers call concurrently.
int get_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t stride,
             uint32_t box_rows) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{base, rows, cols, stride, box_rows};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  CUtensorMap m;
  int rc = make_tmap_2d_sw128(&m, base, rows, cols, stride, box_rows, kBK);
  if (rc) return rc;
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 8192) cache.clear();
    cache[key] = m;
  }
  *out = m;
  return 0;
}

int gemm_cluster_try(void* C, const void* A, const void* W, int64_t M, int64_t N, int64_t K, int64_t lda,
                     int64_t ldw, int64_t ldc, int dtype, int out_f32, int sms, cudaStream_t stream);

int gemm_pair_try(void* C, const void* A, const void* W, int64_t M, int64_t N, int64_t K, int64_t lda,
                  int64_t ldw, int64_t ldc, int dtype, int out_f32, int epilogue, int sms, int panel_mb,
                  cudaStream_t stream);

int gemm_pair_decode_try(void* C, const void* A, const void* W, int64_t M, int64_t N, int64_t K, int64_t lda,
                         int64_t ldw, int64_t ldc, int dtype, int out_f32, int sms, float* ws, int64_t ws_bytes,
                         int* flags, int n_flags, int max_split, cudaStream_t stream);

constexpr int64_t kGemmWsBytes = 48ll << 20;  // stream-K partials: one [128 x BN] fp32 tile per CTA
constexpr int kGemmCounters = 4096;

template <typename T, int BN, bool kSwap, int kEpi>
static int launch_gemm(const CUtensorMap& tx, const CUtensorMap& ty, const GemmParams& p, int grid,
                       cudaStream_t stream) {
  auto kern = gemm_kernel<T, BN, kSwap, kEpi>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    SPD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  gemm_smem_bytes<BN>()));
    configured = true;
  }
  SPD_CUDA(launch(kern, dim3(grid), dim3(kGemmThreads), gemm_smem_bytes<BN>(), stream,
                  pdl_enabled(), tx, ty, p));
  return 0;
}

}  // namespace spd

using namespace spd;

extern "C" {

int64_t spd_gemm_workspace_bytes(void) { return kGemmWsBytes + kGemmCounters * 4; }

int spd_gemm(void* C, const void* A, const void* W, int64_t M, int64_t N, int64_t K, int64_t lda,
             int64_t ldw, int64_t ldc, int dtype, int out_f32, int epilogue, void* workspace,
             int64_t workspace_bytes, int sm_budget, spd_stream_t stream) {
  SPD_REQUIRE(C && A && W, SPD_E_BADARG, "spd_gemm: null pointer");
  SPD_REQUIRE(M >= 0 && N > 0 && K > 0, SPD_E_BADARG, "spd_gemm: bad sizes M=%lld N=%lld K=%lld",
              (long long)M, (long long)N, (long long)K);
  SPD_REQUIRE(dtype == SPD_BF16 || dtype == SPD_F16, SPD_E_BADARG, "spd_gemm: bad dtype %d", dtype);
  SPD_REQUIRE(epilogue == SPD_EPI_NONE || epilogue == SPD_EPI_SILU_MUL, SPD_E_BADARG,
              "spd_gemm: bad epilogue %d", epilogue);
  SPD_REQUIRE(K % 8 == 0 && N % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldc % (out_f32 ? 4 : 8) == 0,
              SPD_E_ALIGN, "spd_gemm: K, N and the leading dimensions must be multiples of 8");
  SPD_REQUIRE(aligned16(A) && aligned16(W) && aligned16(C), SPD_E_ALIGN,
              "spd_gemm: pointers must be 16-byte aligned");
  SPD_REQUIRE(lda >= K && ldw >= K && ldc >= N, SPD_E_BADARG, "spd_gemm: leading dimension too small");
  if (M == 0) return 0;
  cudaStream_t st = as_stream(stream);
  int sms = sm_budget > 0 ? sm_budget : device_sm_count_cached();
  const int kblocks = static_cast<int>((K + kBK - 1) / kBK);

  GemmParams p{};
  p.C = C;
  p.ldc = ldc;
  p.kblocks = kblocks;
  p.out_f32 = out_f32;
  p.streamk = 0;
  p.ws = reinterpret_cast<float*>(workspace);
  p.counters = workspace ? reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(workspace) + kGemmWsBytes)
                         : nullptr;
  CUtensorMap tx, ty;
  int rc;

  // stream-K grid: one equal unit range per SM, but never so thin that a tile is shared by more
  // than ~SPD_GEMM_SPLIT CTAs (the owner reads one partial per contributor) or a CTA gets < 4
  // k-blocks.  Tuning knobs (experiments only): SPD_GEMM_STREAMK=0 disables stream-K, =2 forces it;
  // SPD_GEMM_SPLIT=d caps the contributors per tile at ~d; SPD_GEMM_SWAPMAX=m is the largest M
  // routed to swap-AB; SPD_GEMM_CLUSTER=1 routes few-tile layers at M <= 128 to the cluster
  // split-K kernel (gemm_cluster.cu) instead.
  static const int env_streamk = getenv("SPD_GEMM_STREAMK") ? atoi(getenv("SPD_GEMM_STREAMK")) : 1;
  static const int env_split = getenv("SPD_GEMM_SPLIT") ? atoi(getenv("SPD_GEMM_SPLIT")) : 4;
  static const int env_swapmax = getenv("SPD_GEMM_SWAPMAX") ? atoi(getenv("SPD_GEMM_SWAPMAX")) : 256;
  const bool ws_ok = env_streamk && workspace && workspace_bytes >= spd_gemm_workspace_bytes();
  auto streamk_grid = [&](long long tiles, int BN) -> int {
    const long long units = tiles * kblocks;
    long long g = sms;
    const long long min_units = kblocks / env_split > 4 ? kblocks / env_split : 4;
    if (units / g < min_units) g = units / min_units > 0 ? units / min_units : 1;
    if (g * kBM * BN * 4 > kGemmWsBytes) g = kGemmWsBytes / (1ll * kBM * BN * 4);
    if (g > kGemmCounters) g = kGemmCounters;
    return static_cast<int>(g);
  };
  // whole tiles dealt round-robin leave SMs idle in the last wave; stream-K when that costs > 10 %
  auto wants_streamk = [&](long long tiles) -> bool {
    if (!ws_ok || kblocks < 8) return false;
    if (env_streamk == 2) return true;
    const long long waves = (tiles + sms - 1) / sms;
    return tiles * 10 < waves * sms * 9;
  };

  // y tiles per rasterisation panel: keep ~40 MB of W rows hot in L2 (tile_coords)
  static const int env_panel_mb = getenv("SPD_GEMM_PANEL_MB") ? atoi(getenv("SPD_GEMM_PANEL_MB")) : 40;
  auto panel_for = [&](int BN) -> int {
    const long long tile_bytes = static_cast<long long>(BN) * K * 2;
    long long q = (static_cast<long long>(env_panel_mb) << 20) / tile_bytes;
    return static_cast<int>(q < 1 ? 1 : q);
  };

  // Decode regime, M <= 256: swap-AB.  The token dimension feeds UMMA-N (16, or M rounded up to a
  // multiple of 32), each 128-row weight tile is read once, and no MMA work is spent on padding
  // rows (the un-swapped form would pad 160 tokens to 2 x 128 and read W twice from L2).
  const bool swap = (M <= env_swapmax) && epilogue == SPD_EPI_NONE;
  if (swap) {
    // experiment knob: SPD_GEMM_PREFER_CLUSTER=1 routes few-tile layers at M <= 128 to the DSMEM
    // split-K kernel before the pair kernel is considered
    static const int env_prefer_cluster = getenv("SPD_GEMM_PREFER_CLUSTER") ? atoi(getenv("SPD_GEMM_PREFER_CLUSTER")) : 0;
    if (env_prefer_cluster && M > 32 && (N + kBM - 1) / kBM <= sms / 2) {
      rc = gemm_cluster_try(C, A, W, M, N, K, lda, ldw, ldc, dtype, out_f32, sms, st);
      if (rc != 1) return rc;
    }
    // 32 < M <= 256: CTA pairs share one activation tile between 256 weight rows (gemm_pair_decode.cu)
    rc = gemm_pair_decode_try(C, A, W, M, N, K, lda, ldw, ldc, dtype, out_f32, sms, ws_ok ? p.ws : nullptr,
                              kGemmWsBytes, ws_ok ? p.counters : nullptr, kGemmCounters, env_split, st);
    if (rc != 1) return rc;
    rc = gemm_cluster_try(C, A, W, M, N, K, lda, ldw, ldc, dtype, out_f32, sms, st);
    if (rc != 1) return rc;
    const int BN = M <= 16 ? 16 : static_cast<int>((M + 31) / 32 * 32);
    p.x_rows = N;
    p.y_rows = M;
    p.x_tiles = static_cast<int>((N + kBM - 1) / kBM);
    p.y_tiles = 1;
    int grid = static_cast<int>(p.x_tiles < sms ? p.x_tiles : sms);
    if (wants_streamk(p.x_tiles)) {
      grid = streamk_grid(p.x_tiles, BN);
      p.streamk = 1;
    }
    if ((rc = get_tmap(&tx, W, N, K, ldw, kBM))) return rc;
    if ((rc = get_tmap(&ty, A, M, K, lda, BN))) return rc;
#define SPD_GEMM_SWAP(BN_)                                                               \
  case BN_:                                                                              \
    SPD_DISPATCH_DTYPE(dtype, { return launch_gemm<T, BN_, true, SPD_EPI_NONE>(tx, ty, p, grid, st); }); \
    break;
    switch (BN) {
      SPD_GEMM_SWAP(16)
      SPD_GEMM_SWAP(32)
      SPD_GEMM_SWAP(64)
      SPD_GEMM_SWAP(96)
      SPD_GEMM_SWAP(128)
      SPD_GEMM_SWAP(160)
      SPD_GEMM_SWAP(192)
      SPD_GEMM_SWAP(224)
      SPD_GEMM_SWAP(256)
    }
#undef SPD_GEMM_SWAP
    return set_error(SPD_E_SHAPE, "spd_gemm: unreachable BN=%d", BN);
  }

  // prefill regime proper: 256 x 256 tiles on CTA pairs (tcgen05.mma.cta_group::2, gemm_pair.cu)
  rc = gemm_pair_try(C, A, W, M, N, K, lda, ldw, ldc, dtype, out_f32, epilogue, sms, env_panel_mb, st);
  if (rc != 1) return rc;

  p.x_rows = M;
  p.y_rows = N;
  p.x_tiles = static_cast<int>((M + kBM - 1) / kBM);
  if (epilogue == SPD_EPI_SILU_MUL) {
    constexpr int BN = 256;  // 128 gate + 128 up columns per tile
    p.y_rows_w = N;
    p.y_tiles = static_cast<int>((N + BN / 2 - 1) / (BN / 2));
    p.panel = panel_for(BN);
    const int total = p.x_tiles * p.y_tiles;
    const int grid = total < sms ? total : sms;
    if ((rc = get_tmap(&tx, A, M, K, lda, kBM))) return rc;
    if ((rc = get_tmap(&ty, W, 2 * N, K, ldw, BN / 2))) return rc;
    SPD_DISPATCH_DTYPE(dtype, { return launch_gemm<T, BN, false, SPD_EPI_SILU_MUL>(tx, ty, p, grid, st); });
  }
  {
    // 128 x 256 tiles unless that leaves most SMs idle, then 128 x 128
    const int tiles256 = p.x_tiles * static_cast<int>((N + 255) / 256);
    const bool use256 = tiles256 >= sms || N % 256 == 0;
    const int BN = use256 ? 256 : 128;
    p.y_tiles = static_cast<int>((N + BN - 1) / BN);
    p.panel = panel_for(BN);
    const int total = p.x_tiles * p.y_tiles;
    int grid = total < sms ? total : sms;
    if (M <= 256 && use256 && wants_streamk(total)) {
      grid = streamk_grid(total, BN);  // decode regime, un-swapped (fused-epilogue callers only)
      p.streamk = 1;
    }
    if ((rc = get_tmap(&tx, A, M, K, lda, kBM))) return rc;
    if ((rc = get_tmap(&ty, W, N, K, ldw, BN))) return rc;
    if (use256) {
      SPD_DISPATCH_DTYPE(dtype, { return launch_gemm<T, 256, false, SPD_EPI_NONE>(tx, ty, p, grid, st); });
    } else {
      SPD_DISPATCH_DTYPE(dtype, { return launch_gemm<T, 128, false, SPD_EPI_NONE>(tx, ty, p, grid, st); });
    }
  }
  return 0;
}


int spd_gemm_grouped(void* C, const void* A, const void* W, int64_t M, int64_t N, int64_t K, int64_t lda,
                     int64_t ldw, int64_t ldc, int num_groups, const int32_t* tile_group,
                     int64_t a_group_stride, int64_t a_cols, int64_t c_group_stride, int64_t w_group_col_stride,
                     int dtype, int epilogue, int sm_budget, spd_stream_t stream) {
  SPD_REQUIRE(C && A && W, SPD_E_BADARG, "spd_gemm_grouped: null pointer");
  SPD_REQUIRE(M >= 0 && N > 0 && K > 0 && num_groups > 0, SPD_E_BADARG, "spd_gemm_grouped: bad sizes");
  SPD_REQUIRE(dtype == SPD_BF16 || dtype == SPD_F16, SPD_E_BADARG, "spd_gemm_grouped: bad dtype %d", dtype);
  SPD_REQUIRE(epilogue == SPD_EPI_NONE || epilogue == SPD_EPI_SILU_MUL, SPD_E_BADARG,
              "spd_gemm_grouped: bad epilogue %d", epilogue);
  SPD_REQUIRE(K % 8 == 0 && N % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0 && a_group_stride % 8 == 0,
              SPD_E_ALIGN, "spd_gemm_grouped: K, N, strides and leading dimensions must be multiples of 8");
  SPD_REQUIRE(aligned16(A) && aligned16(W) && aligned16(C), SPD_E_ALIGN,
              "spd_gemm_grouped: pointers must be 16-byte aligned");
  if (M == 0) return 0;
  const bool rows_mode = tile_group != nullptr;
  cudaStream_t st = as_stream(stream);
  const int sms = sm_budget > 0 ? sm_budget : device_sm_count_cached();
  GemmParams p{};
  p.C = C;
  p.ldc = ldc;
  p.kblocks = static_cast<int>((K + kBK - 1) / kBK);
  p.x_rows = M;
  p.x_tiles = static_cast<int>((M + kBM - 1) / kBM);
  p.panel = 0;
  CUtensorMap tx, ty;
  int rc;
  if (rows_mode) {
    // MoE: rows of A sorted by expert in 128-row tiles; W = the experts' weights stacked
    SPD_REQUIRE(a_group_stride == 0 && c_group_stride == 0, SPD_E_BADARG,
                "spd_gemm_grouped: a_group_stride / c_group_stride are cols-mode arguments");
    SPD_REQUIRE(w_group_col_stride % 8 == 0 && (w_group_col_stride == 0 || epilogue == SPD_EPI_NONE),
                SPD_E_BADARG, "spd_gemm_grouped: bad w_group_col_stride");
    p.x_tile_group = tile_group;
    p.y_rows = N;
    if ((rc = get_tmap(&tx, A, M, K, lda, kBM))) return rc;
    if (epilogue == SPD_EPI_SILU_MUL) {
      constexpr int BN = 256;
      SPD_REQUIRE(N % 8 == 0, SPD_E_SHAPE, "spd_gemm_grouped: bad N");
      p.y_group_rows = 2 * N;
      p.y_rows_w = N;
      p.y_tiles = static_cast<int>((N + BN / 2 - 1) / (BN / 2));
      const int total = p.x_tiles * p.y_tiles;
      const int grid = total < sms ? total : sms;
      if ((rc = get_tmap(&ty, W, static_cast<uint64_t>(num_groups) * 2 * N, K, ldw, BN / 2))) return rc;
      SPD_DISPATCH_DTYPE(dtype, { return launch_gemm<T, BN, false, SPD_EPI_SILU_MUL>(tx, ty, p, grid, st); });
    }
    constexpr int BN = 128;
    p.y_tiles = static_cast<int>((N + BN - 1) / BN);
    const int total = p.x_tiles * p.y_tiles;
    const int grid = total < sms ? total : sms;
    if (w_group_col_stride > 0) {
      // weights concatenated along K: W [N, num_groups * w_group_col_stride], expert g = columns
      // [g * stride, g * stride + K).  K must be a whole number of 64-column blocks, or the last
      // block would read into the next expert.
      SPD_REQUIRE(K % kBK == 0, SPD_E_SHAPE, "spd_gemm_grouped: K %% 64 != 0 with K-concatenated weights");
      p.y_group_cols = w_group_col_stride;
      if ((rc = get_tmap(&ty, W, N, static_cast<uint64_t>(num_groups) * w_group_col_stride, ldw, BN))) return rc;
    } else {
      p.y_group_rows = N;
      if ((rc = get_tmap(&ty, W, static_cast<uint64_t>(num_groups) * N, K, ldw, BN))) return rc;
    }
    SPD_DISPATCH_DTYPE(dtype, { return launch_gemm<T, BN, false, SPD_EPI_NONE>(tx, ty, p, grid, st); });
  }
  // cols mode (block diagonal): C[:, g*N:(g+1)*N] = A[:, g*a_group_stride : +K] * W[g*N:(g+1)*N, :K]^T
  SPD_REQUIRE(epilogue == SPD_EPI_NONE && w_group_col_stride == 0, SPD_E_BADARG,
              "spd_gemm_grouped: no epilogue / K-concatenated weights in cols mode");
  SPD_REQUIRE(c_group_stride == 0 || (c_group_stride >= N && c_group_stride % 8 == 0), SPD_E_BADARG,
              "spd_gemm_grouped: bad c_group_stride");
  SPD_REQUIRE(N % 128 == 0 && K % kBK == 0, SPD_E_SHAPE,
              "spd_gemm_grouped: cols mode needs N %% 128 == 0 and K %% 64 == 0 (got N=%lld K=%lld)",
              (long long)N, (long long)K);
  SPD_REQUIRE(a_cols >= (num_groups - 1) * a_group_stride + K && lda >= a_cols, SPD_E_BADARG,
              "spd_gemm_grouped: a_cols / lda too small");
  constexpr int BN = 128;
  p.y_rows = static_cast<int64_t>(num_groups) * N;
  p.y_tiles = static_cast<int>(p.y_rows / BN);
  p.y_tiles_per_group = static_cast<int>(N / BN);
  p.x_group_cols = a_group_stride;
  p.c_group_cols = c_group_stride;
  const int total = p.x_tiles * p.y_tiles;
  const int grid = total < sms ? total : sms;
  if ((rc = get_tmap(&tx, A, M, a_cols, lda, kBM))) return rc;
  if ((rc = get_tmap(&ty, W, p.y_rows, K, ldw, BN))) return rc;
  SPD_DISPATCH_DTYPE(dtype, { return launch_gemm<T, BN, false, SPD_EPI_NONE>(tx, ty, p, grid, st); });
  return 0;
}

}  // extern "C"
