// Decode-regime GEMM on CTA PAIRS: C[M,N] = A[M,K] * W[N,K]^T for 32 < M <= 256 tokens, swap-AB
// with tcgen05.mma.cta_group::2.
//
// Why: measured on B200 (profiles/r01_gemm_decode_sweep.log, scripts/bench_kernels.py gd) the
// single-CTA swap-AB kernel (gemm.cu) delivers a constant ~8-9 TB/s of operand bytes into the SMs
// whatever M is — the chip-wide L2 -> SM limit — and every 128-row weight tile re-reads the whole
// activation matrix, so the weight stream runs at HBM speed only while M is tiny:
// bytes into the SMs = weights x (1 + M/128).  A CTA pair shares ONE activation tile between 256
// weight rows (each CTA stages half of the token rows, the 2-SM MMA reads both halves), which
// makes it weights x (1 + M/256).
//
//   D[256 weight rows x BN tokens] = W_tile[256 x K] * A[BN x K]^T        (BN = M rounded up to 32)
//   CTA r of the pair: A-operand rows [r*128, r*128+128) of the weight tile, B-operand rows
//   [r*BN/2, (r+1)*BN/2) of the activations, accumulator rows r*128.. in its own TMEM.
//
// Work is cut STREAM-K style over the pairs (equal (tile x k-block) unit ranges, see gemm.cu): the
// pair holding a tile's last k-blocks owns it and parks that accumulator in TMEM slot 0; the other
// contributors publish fp32 partials as their last piece of work; the owner adds them in
// ascending-k order at the very end (bit-deterministic, no atomics on data, no waiting before
// one's own partial is published).  Barriers as in gemm_pair.cu (leader's full / tempty, multicast
// commits for empty / tfull).
#include <cuda.h>
#include <stdlib.h>
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace spd {

__device__ __forceinline__ long long pd_now() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define PD_STAMP(i) do { if (p.dbg && et == 0) p.dbg[blockIdx.x * 8 + (i)] = pd_now(); } while (0)

constexpr int kPdThreads = 256;
constexpr int kPdBM = 128;  // weight rows per CTA (pair tile = 256)
constexpr int kPdBK = 64;
constexpr int kPdWBytes = kPdBM * kPdBK * 2;  // 16 KB
constexpr int kPdSmemBudget = 208 * 1024;
constexpr int kPdStageOutBytes = 32 * kPdBM * 4;  // epilogue transpose staging: 32 tokens x 128 rows x fp32

template <int BN>
__host__ __device__ constexpr int pd_stage_bytes() { return kPdWBytes + (BN / 2) * kPdBK * 2; }
template <int BN>
__host__ __device__ constexpr int pd_stages() {
  return (kPdSmemBudget / pd_stage_bytes<BN>()) < 12 ? (kPdSmemBudget / pd_stage_bytes<BN>()) : 12;
}
template <int BN>
__host__ __device__ constexpr int pd_smem_bytes() {
  return pd_stages<BN>() * pd_stage_bytes<BN>() + 1024 + 256 + kPdStageOutBytes;
}

struct PairDecodeParams {
  void* C;
  int64_t ldc;
  int64_t n_rows;  // N (weight rows)
  int64_t m_rows;  // M (tokens)
  int tiles, kblocks;
  int streamk;
  int out_f32;
  float* ws;   // one [BN/4][128][4] fp32 partial per CTA
  int* flags;  // one "partial published" flag per CTA
  long long* dbg;  // optional timeline stamps [cta][8] (globaltimer ns); nullptr in production
};

// A pair's work list: segments (tile, [kb0, kb1)); identical in every warp role of both CTAs.
struct PdSegIter {
  int kblocks, tiles, streamk, npairs;
  long long u, u_end;
  int w;
  __device__ __forceinline__ static long long range_start(long long c, long long units, long long g) {
    return c * units / g;
  }
  __device__ __forceinline__ void init(const PairDecodeParams& p, int pair, int npairs_) {
    kblocks = p.kblocks;
    tiles = p.tiles;
    streamk = p.streamk;
    npairs = npairs_;
    if (streamk) {
      const long long units = static_cast<long long>(tiles) * kblocks;
      u = range_start(pair, units, npairs);
      u_end = range_start(pair + 1, units, npairs);
    } else {
      w = pair;
    }
  }
  __device__ __forceinline__ bool next(int& tile, int& kb0, int& kb1) {
    if (streamk) {
      if (u >= u_end) return false;
      tile = static_cast<int>(u / kblocks);
      kb0 = static_cast<int>(u - static_cast<long long>(tile) * kblocks);
      const long long left = u_end - u;
      kb1 = (kblocks - kb0) < left ? kblocks : kb0 + static_cast<int>(left);
      u += kb1 - kb0;
    } else {
      if (w >= tiles) return false;
      tile = w;
      kb0 = 0;
      kb1 = kblocks;
      w += npairs;
    }
    return true;
  }
};

// cta_group::2 instruction forms (as in gemm_pair.cu)
__device__ __forceinline__ void pd_tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void pd_tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void pd_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void pd_tma_load(void* smem_dst, const void* tmap, uint32_t bar_cluster_addr,
                                            int32_t c_inner, int32_t c_outer, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(bar_cluster_addr), "r"(c_inner), "r"(c_outer), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void pd_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void pd_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

template <typename T, int BN>
__global__ void __launch_bounds__(kPdThreads, 1)
gemm_pair_decode_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA,
                        const PairDecodeParams p) {
  constexpr int kStages = pd_stages<BN>();
  constexpr int kABytes = (BN / 2) * kPdBK * 2;
  constexpr int kTmemCols = (2 * BN) <= 64 ? 64 : (2 * BN) <= 128 ? 128 : (2 * BN) <= 256 ? 256 : 512;
  constexpr int kChunk = 32;
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN: multiple of 32 in [32, 256]");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sW = smem;
  uint8_t* sA = smem + kStages * kPdWBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * (kPdWBytes + kABytes));
  uint64_t* full = bars;                      // [kStages]  (leader's are the live ones)
  uint64_t* empty = bars + kStages;           // [kStages]
  uint64_t* tfull = bars + 2 * kStages;       // [2]
  uint64_t* tempty = bars + 2 * kStages + 2;  // [2]        (leader's are the live ones)
  uint64_t* fixbar = bars + 2 * kStages + 4;  // owner fix-up: bulk copies of the partials have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 5);
  uint8_t* sOut = smem + kStages * (kPdWBytes + kABytes) + 256;  // 16 KB, 256-byte aligned

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmA);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    mbar_init(fixbar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    pd_tmem_alloc(tmem_slot, kTmemCols);
    pd_tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers exist before any cross-CTA signal
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // Programmatic dependent launch, both ends.  The dependents of this grid may be scheduled as
  // soon as every CTA is running (they block in their own griddepcontrol.wait until this grid has
  // completed); and this grid was itself launched while its predecessor was still running: the
  // WEIGHTS do not depend on the predecessor, so the producer fills the whole TMA ring with
  // weight tiles BEFORE it waits — a decode GEMM that follows RMSNorm / RoPE / SiLU·mul / the
  // attention streams its first ~200 KB per SM (x 148 = 29 MB: all of W_o, 60 % of W_qkv) in the
  // shadow of that kernel.  Only the activation tiles and every global write wait.
  griddep_launch_dependents();

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0, phase = 0;
      const int a_row = static_cast<int>(rank) * (BN / 2);
      // pass 0: weights of the first kStages k-blocks (no predecessor data); then wait; pass 1:
      // their activation tiles, and everything after them as usual
      int ahead = 0;
      {
        PdSegIter it;
        it.init(p, pair, npairs);
        int tile, kb0, kb1;
        while (ahead < kStages && it.next(tile, kb0, kb1)) {
          const int w_row = tile * (2 * kPdBM) + static_cast<int>(rank) * kPdBM;
          for (int kb = kb0; kb < kb1 && ahead < kStages; ++kb, ++ahead) {
            if (rank == 0) mbar_expect_tx(&full[ahead], 2 * (kPdWBytes + kABytes));
            pd_tma_load(sW + ahead * kPdWBytes, &tmW, map_to_rank(smem_u32(&full[ahead]), 0), kb * kPdBK, w_row,
                        kEvictFirst);
          }
        }
      }
      griddep_wait();
      PdSegIter it;
      it.init(p, pair, npairs);
      int tile, kb0, kb1, n = 0;
      while (it.next(tile, kb0, kb1)) {
        const int w_row = tile * (2 * kPdBM) + static_cast<int>(rank) * kPdBM;
        for (int kb = kb0; kb < kb1; ++kb, ++n) {
          const uint32_t lead_full = map_to_rank(smem_u32(&full[stage]), 0);
          if (n >= ahead) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (rank == 0) mbar_expect_tx(&full[stage], 2 * (kPdWBytes + kABytes));
            // weights stream through once; the activations are re-read by every pair
            pd_tma_load(sW + stage * kPdWBytes, &tmW, lead_full, kb * kPdBK, w_row, kEvictFirst);
          }
          pd_tma_load(sA + stage * kABytes, &tmA, lead_full, kb * kPdBK, a_row, kEvictLast);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = umma_idesc<T>(2 * kPdBM, BN, 0, 0);
      int stage = 0, phase = 0, n = 0;
      PdSegIter it;
      it.init(p, pair, npairs);
      int tile, kb0, kb1;
      while (it.next(tile, kb0, kb1)) {
        // stream-K: slot 0 is the first segment's (parked until the end if this pair owns that
        // tile), every later segment goes through slot 1
        const int as = p.streamk ? (n == 0 ? 0 : 1) : (n & 1);
        const int use = p.streamk ? (n == 0 ? 0 : n - 1) : (n >> 1);
        ++n;
        mbar_wait_cluster_acquire(&tempty[as], (use & 1) ^ 1);  // both CTAs drained this accumulator
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after_sync();
          const uint32_t wa = smem_u32(sW + stage * kPdWBytes);
          const uint32_t aa = smem_u32(sA + stage * kABytes);
#pragma unroll
          for (int k = 0; k < kPdBK / 16; ++k)
            pd_umma(d_tmem, umma_desc_sw128(wa + k * 32, 16, 1024), umma_desc_sw128(aa + k * 32, 16, 1024), idesc,
                    (kb > kb0 || k > 0) ? 1u : 0u);
          pd_commit_mc(&empty[stage], 0b11);  // both CTAs' slots are free once these MMAs are done
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        pd_commit_mc(&tfull[as], 0b11);
      }
    }
  } else if (warp >= 4) {
    griddep_wait();  // C, the split-K workspace and its flags are global memory
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 128;
    const uint32_t lead_tempty0 = map_to_rank(smem_u32(&tempty[0]), 0);
    const long long units = static_cast<long long>(p.tiles) * p.kblocks;
    const int n_cols = static_cast<int>(((p.m_rows + kChunk - 1) / kChunk) * kChunk);
    const int cols_used = n_cols < BN ? n_cols : BN;

    // D[row][c + j] -> C[token c + j][nn].  The accumulator has the weight row on the TMEM lane,
    // C has it on the contiguous axis: the 32-token chunk is transposed through shared memory so
    // that C is written with 16-byte stores, 256-512 contiguous bytes per token row (the direct
    // form, one 2-byte store per element behind per-element bounds checks, kept the owners of
    // split tiles busy for ~10 us after everybody else at M = 128).
    auto store_chunk_as = [&](auto* out_ptr, int tile, int c, const float* v) {
      using OutT = std::remove_pointer_t<decltype(out_ptr)>;
      constexpr int E = 16 / static_cast<int>(sizeof(OutT));  // elements per 16-byte vector
      constexpr int kVecPerRow = kPdBM / E;
      OutT* st = reinterpret_cast<OutT*>(sOut);
#pragma unroll
      for (int j = 0; j < kChunk; ++j) {
        if constexpr (sizeof(OutT) == 4) st[j * kPdBM + row] = to_f32<T>(from_f32<T>(v[j]));
        else st[j * kPdBM + row] = from_f32<T>(v[j]);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int64_t nn_base = static_cast<int64_t>(tile) * (2 * kPdBM) + rank * kPdBM;
#pragma unroll
      for (int i = 0; i < kChunk * kVecPerRow / 128; ++i) {
        const int idx = i * 128 + et;
        const int tok = idx / kVecPerRow, vec = idx % kVecPerRow;
        const int64_t m = c + tok, nn = nn_base + vec * E;
        if (m < p.m_rows && nn < p.n_rows)
          *reinterpret_cast<uint4*>(out_ptr + m * p.ldc + nn) =
              *reinterpret_cast<const uint4*>(st + tok * kPdBM + vec * E);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // staging is rewritten by the next chunk
    };
    auto store_chunk = [&](int tile, int c, const float* v) {
      if (p.out_f32) store_chunk_as(reinterpret_cast<float*>(p.C), tile, c, v);
      else store_chunk_as(reinterpret_cast<T*>(p.C), tile, c, v);
    };

    PD_STAMP(0);  // roles start
    int n = 0;
    int own_tile = -1;
    PdSegIter it;
    it.init(p, pair, npairs);
    int tile, kb0, kb1;
    while (it.next(tile, kb0, kb1)) {
      const int as = p.streamk ? (n == 0 ? 0 : 1) : (n & 1);
      const int use = p.streamk ? (n == 0 ? 0 : n - 1) : (n >> 1);
      ++n;
      const bool complete = (kb0 == 0 && kb1 == p.kblocks);
      if (!complete && kb1 == p.kblocks) {
        own_tile = tile;  // tail of the tile: this pair owns it (always its first segment, slot 0)
        continue;
      }
      mbar_wait(&tfull[as], use & 1);
      tc_fence_after_sync();
      PD_STAMP(1);  // last non-owner segment's MMAs complete
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
      if (!complete) {
        // contributor (always the pair's last segment): publish this CTA's fp32 partial,
        // [col/4][lane][4] so that one warp store covers 512 contiguous bytes
        float* mine = p.ws + static_cast<size_t>(blockIdx.x) * (kPdBM * BN) + row * 4;
#pragma unroll 1
        for (int c = 0; c < cols_used; c += kChunk) {
          uint32_t r[kChunk];
          tmem_ld_32x32b_x32(taddr + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < kChunk; j += 4)
            *reinterpret_cast<uint4*>(mine + static_cast<size_t>((c + j) >> 2) * (kPdBM * 4)) =
                make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote_release(lead_tempty0 + as * 8);
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0)
          asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.flags + blockIdx.x), "r"(1) : "memory");
        PD_STAMP(2);  // partial published
        continue;
      }
#pragma unroll 1
      for (int c = 0; c < cols_used; c += kChunk) {
        uint32_t r[kChunk];
        float v[kChunk];
        tmem_ld_32x32b_x32(taddr + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < kChunk; ++j) v[j] = __uint_as_float(r[j]);
        store_chunk(tile, c, v);
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote_release(lead_tempty0 + as * 8);
    }

    if (own_tile >= 0) {
      // deferred fix-up: partials of the other contributing pairs (ascending pair = ascending k),
      // same rank as this CTA, + the accumulator parked in slot 0
      const long long t0 = static_cast<long long>(own_tile) * p.kblocks;
      const int c_first = static_cast<int>(((t0 + 1) * npairs - 1) / units);
      const int c_last = pair;
      PD_STAMP(3);  // owner: loop done
      mbar_wait(&tfull[0], 0);
      tc_fence_after_sync();
      PD_STAMP(4);  // owner: parked accumulator complete
      for (int cc = c_first + et; cc < c_last; cc += 128) {
        int f;
        do {
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(p.flags + 2 * cc + rank) : "memory");
        } while (f == 0);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      PD_STAMP(5);  // owner: all contributor flags seen
      // The partials are pulled into the (now idle) operand ring by bulk copies, all in flight at
      // once, instead of per-thread L2 round trips (a timeline of the per-thread version showed
      // the owners finishing up to 10 us after everybody else at M = 128).  G contributors fit the
      // ring per round; a running sum between rounds lives in TMEM slot 1 (free by now), so the
      // summation order stays ascending-k whatever G is.
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      const uint32_t part_bytes = static_cast<uint32_t>(cols_used) * (kPdBM * 4);
      constexpr int kRingBytes = kStages * (kPdWBytes + kABytes);
      const int G = max(1, kRingBytes / static_cast<int>(part_bytes));
      int round = 0;
      for (int cc0 = c_first; cc0 < c_last; cc0 += G, ++round) {
        const int g_n = min(G, c_last - cc0);
        if (et == 0) {
          asm volatile("fence.proxy.async;" ::: "memory");  // partials were written / observed through the generic proxy
          mbar_expect_tx(fixbar, g_n * part_bytes);
          for (int i = 0; i < g_n; ++i) {
            const float* src = p.ws + static_cast<size_t>(2 * (cc0 + i) + rank) * (kPdBM * BN);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(smem + static_cast<size_t>(i) * part_bytes)),
                         "l"(src), "r"(part_bytes), "r"(smem_u32(fixbar))
                         : "memory");
          }
        }
        mbar_wait(fixbar, round & 1);
        const bool first_round = (cc0 == c_first), last_round = (cc0 + G >= c_last);
#pragma unroll 1
        for (int c = 0; c < cols_used; c += kChunk) {
          float v[kChunk];
          if (first_round) {
#pragma unroll
            for (int j = 0; j < kChunk; ++j) v[j] = 0.f;
          } else {
            uint32_t r[kChunk];
            tmem_ld_32x32b_x32(taddr + BN + c, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < kChunk; ++j) v[j] = __uint_as_float(r[j]);
          }
          for (int i = 0; i < g_n; ++i) {
            const float* ps = reinterpret_cast<const float*>(smem + static_cast<size_t>(i) * part_bytes) + row * 4;
#pragma unroll
            for (int j = 0; j < kChunk; j += 4) {
              const float4 f = *reinterpret_cast<const float4*>(ps + static_cast<size_t>((c + j) >> 2) * (kPdBM * 4));
              v[j] += f.x; v[j + 1] += f.y; v[j + 2] += f.z; v[j + 3] += f.w;
            }
          }
          if (last_round) {
            uint32_t r[kChunk];
            tmem_ld_32x32b_x32(taddr + c, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < kChunk; ++j) v[j] += __uint_as_float(r[j]);
            store_chunk(own_tile, c, v);
          } else {
            uint32_t r[kChunk];
#pragma unroll
            for (int j = 0; j < kChunk; ++j) r[j] = __float_as_uint(v[j]);
            tmem_st_32x32b_x32(taddr + BN + c, r);
            tmem_st_wait();
          }
        }
        // the next round's copies overwrite the staging area: everybody must be done reading it
        if (!last_round) asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      // re-arm the flags for the next launch / graph replay (each flag has exactly one reader)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int cc = c_first + et; cc < c_last; cc += 128) p.flags[2 * cc + rank] = 0;
    }
    PD_STAMP(6);  // epilogue role done
  }

  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // the peer may still read this CTA's shared memory / signal its barriers
  if (warp == 2) pd_tmem_dealloc(tmem_base, kTmemCols);
}

template <typename T, int BN>
static int launch_pair_decode(const CUtensorMap& tw, const CUtensorMap& ta, const PairDecodeParams& p, int pairs,
                              cudaStream_t stream) {
  auto kern = gemm_pair_decode_kernel<T, BN>;
  static bool configured = false;
  if (!configured) {
    SPD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pd_smem_bytes<BN>()));
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kPdThreads);
  cfg.dynamicSmemBytes = pd_smem_bytes<BN>();
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  SPD_CUDA(cudaLaunchKernelEx(&cfg, kern, tw, ta, p));
  return 0;
}

// Called by spd_gemm for the swap-AB regime.  Returns 1 if this path does not apply (the caller
// falls through to the single-CTA kernel), 0 on launch, else an error code.
// ws: fp32 partials (ws_bytes), flags: ints (n_flags), both zero-initialised once by the caller.
int gemm_pair_decode_try(void* C, const void* A, const void* W, int64_t M, int64_t N, int64_t K, int64_t lda,
                         int64_t ldw, int64_t ldc, int dtype, int out_f32, int sms, float* ws, int64_t ws_bytes,
                         int* flags, int n_flags, int max_split, cudaStream_t stream) {
  static const int enabled = getenv("SPD_GEMM_PAIR_DECODE") ? atoi(getenv("SPD_GEMM_PAIR_DECODE")) : 1;
  static const int min_m = getenv("SPD_GEMM_PAIR_MIN_M") ? atoi(getenv("SPD_GEMM_PAIR_MIN_M")) : 33;
  if (!enabled || M < min_m || M > 256 || sms < 2 || N < 256) return 1;
  const int BN = static_cast<int>((M + 31) / 32 * 32);
  PairDecodeParams p{};
  p.C = C;
  p.ldc = ldc;
  p.n_rows = N;
  p.m_rows = M;
  p.out_f32 = out_f32;
  p.kblocks = static_cast<int>((K + kPdBK - 1) / kPdBK);
  p.tiles = static_cast<int>((N + 2 * kPdBM - 1) / (2 * kPdBM));
  p.ws = ws;
  p.flags = flags;
  // debugging aid: SPD_GEMM_DBG_PTR=<device address of a zeroed int64[grid*8] buffer> records a timeline
  static const char* dbg_env = getenv("SPD_GEMM_DBG_PTR");
  p.dbg = dbg_env ? reinterpret_cast<long long*>(strtoull(dbg_env, nullptr, 0)) : nullptr;
  const int max_pairs = sms / 2;
  int pairs = p.tiles < max_pairs ? p.tiles : max_pairs;
  // whole tiles dealt round-robin leave pairs idle in the last wave; stream-K when that costs > 10 %
  const long long waves = (p.tiles + max_pairs - 1) / max_pairs;
  const bool ws_ok = ws && flags && p.kblocks >= 8;
  if (ws_ok && static_cast<long long>(p.tiles) * 10 < waves * max_pairs * 9) {
    const long long units = static_cast<long long>(p.tiles) * p.kblocks;
    long long g = max_pairs;
    const long long min_units = p.kblocks / max_split > 4 ? p.kblocks / max_split : 4;
    if (units / g < min_units) g = units / min_units > 0 ? units / min_units : 1;
    const long long per_cta = static_cast<long long>(kPdBM) * BN * 4;
    if (2 * g * per_cta > ws_bytes) g = ws_bytes / (2 * per_cta);
    if (2 * g > n_flags) g = n_flags / 2;
    if (g >= 1) {
      pairs = static_cast<int>(g);
      p.streamk = 1;
    }
  }
  CUtensorMap tw, ta;
  int rc;
  if ((rc = get_tmap(&tw, W, N, K, ldw, kPdBM))) return rc;
  if ((rc = get_tmap(&ta, A, M, K, lda, BN / 2))) return rc;
#define SPD_PD(BN_)                                                                            \
  case BN_:                                                                                    \
    SPD_DISPATCH_DTYPE(dtype, { return launch_pair_decode<T, BN_>(tw, ta, p, pairs, stream); }); \
    break;
  switch (BN) {
    SPD_PD(64)
    SPD_PD(96)
    SPD_PD(128)
    SPD_PD(160)
    SPD_PD(192)
    SPD_PD(224)
    SPD_PD(256)
  }
#undef SPD_PD
  return 1;
}

}  // namespace spd
