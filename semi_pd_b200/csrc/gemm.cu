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
//     streams the same number of weight bytes whatever N is (32 tiles or 1002); a tile that is
//     shared by several CTAs is summed from fp32 partials by its last-arriving CTA in a fixed
//     order (bit-deterministic, no atomics on data).
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
constexpr int kMaxStages = 8;
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
  float* ws;        // stream-K partials: [cta][2][lane][col]
  int* counters;    // one per tile
};

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
  constexpr int kTmemCols = (2 * BN) < 32 ? 32 : (2 * BN);
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
  int* flag_slot = reinterpret_cast<int*>(tmem_slot + 1);

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

  // everything above overlaps the previous kernel's tail under PDL
  griddep_wait();

  if (warp == 0) {
    if (elect_one()) {
      // operand cache policy: in the decode regime the weights (X) stream through once
      const uint64_t polX = kSwap ? kEvictFirst : kEvictNormal;
      const uint64_t polY = kSwap ? kEvictLast : kEvictNormal;
      int stage = 0, phase = 0;
      SegIter it;
      it.init(p);
      int tile, kb0, kb1, slot;
      while (it.next(tile, kb0, kb1, slot)) {
        int xt, yt;
        tile_coords(p, tile, xt, yt);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], kXBytes + kYBytes);
          tma_load_2d(sX + stage * kXBytes, &tmX, &full[stage], kb * kBK, xt * kBM, polX);
          if constexpr (kEpi == SPD_EPI_SILU_MUL) {
            tma_load_2d(sY + stage * kYBytes, &tmY, &full[stage], kb * kBK, yt * kBNOut, polY);
            tma_load_2d(sY + stage * kYBytes + kYBytes / 2, &tmY, &full[stage], kb * kBK,
                        static_cast<int>(p.y_rows_w) + yt * kBNOut, polY);
          } else {
            tma_load_2d(sY + stage * kYBytes, &tmY, &full[stage], kb * kBK, yt * BN, polY);
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
        const int as = n & 1, aphase = (n >> 1) & 1;
        ++n;
        mbar_wait(&tempty[as], aphase ^ 1);
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
    int n = 0;
    SegIter it;
    it.init(p);
    int tile, kb0, kb1, slot;
    while (it.next(tile, kb0, kb1, slot)) {
      int xt, yt;
      tile_coords(p, tile, xt, yt);
      const int as = n & 1, aphase = (n >> 1) & 1;
      ++n;
      mbar_wait(&tfull[as], aphase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
      const bool complete = (kb0 == 0 && kb1 == p.kblocks);
      bool do_store = true;
      int c_first = 0, c_last = 0;

      if (!complete) {
        // dump this CTA's fp32 partial ([lane][col], each thread owns one contiguous row), then
        // the last-arriving contributor of the tile sums all of them
        float* mine = p.ws + (static_cast<size_t>(blockIdx.x) * 2 + slot) * (kBM * BN) + row * BN;
#pragma unroll 1
        for (int c = 0; c < BN; c += kChunk) {
          uint32_t r[kChunk];
          if constexpr (kChunk == 32) tmem_ld_32x32b_x32(taddr + c, r);
          else tmem_ld_32x32b_x16(taddr + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < kChunk; j += 4)
            *reinterpret_cast<uint4*>(mine + c + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[as]);
        // contributors of this tile = the CTAs whose unit ranges intersect it
        const long long t0 = static_cast<long long>(tile) * p.kblocks;
        c_first = static_cast<int>(((t0 + 1) * gridDim.x - 1) / units);
        c_last = static_cast<int>(((t0 + p.kblocks) * gridDim.x - 1) / units);
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) *flag_slot = atomicAdd(&p.counters[tile], 1);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        do_store = (*flag_slot == c_last - c_first);
        if (do_store) __threadfence();
      }

      if (do_store) {
#pragma unroll 1
        for (int c = 0; c < kBNOut; c += kChunk) {
          float v[kChunk];
          if (!complete) {
#pragma unroll
            for (int j = 0; j < kChunk; ++j) v[j] = 0.f;
            // fixed summation order (ascending CTA = ascending k) -> deterministic
            for (int cc = c_first; cc <= c_last; ++cc) {
              // a contributor's piece of this tile is its FIRST segment iff its range starts in it
              const long long st = SegIter::range_start(cc, units, gridDim.x);
              const int sl = (st / p.kblocks == tile) ? 0 : 1;
              const float* ps = p.ws + (static_cast<size_t>(cc) * 2 + sl) * (kBM * BN) + row * BN + c;
#pragma unroll
              for (int j = 0; j < kChunk; j += 4) {
                const float4 f = __ldcg(reinterpret_cast<const float4*>(ps + j));
                v[j] += f.x; v[j + 1] += f.y; v[j + 2] += f.z; v[j + 3] += f.w;
              }
            }
          } else {
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
                v[j] = g / (1.0f + __expf(-g)) * up;
              }
            } else {
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < kChunk; ++j) v[j] = __uint_as_float(r[j]);
            }
          }
          if constexpr (!kSwap) {
            const int64_t m = static_cast<int64_t>(xt) * kBM + row;
            const int64_t n0 = static_cast<int64_t>(yt) * kBNOut + c;
            if (m < p.x_rows) {
              if (p.out_f32) {
                float* dst = reinterpret_cast<float*>(p.C) + m * p.ldc + n0;
#pragma unroll
                for (int j = 0; j < kChunk; j += 4) {
                  if (n0 + j < p.y_rows) {
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
                  if (n0 + j < p.y_rows) {
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
            // swapped: lane index runs along N (contiguous in C), column j is the token
            const int64_t nn = static_cast<int64_t>(xt) * kBM + row;
            if (nn < p.x_rows) {
#pragma unroll
              for (int j = 0; j < kChunk; ++j) {
                const int64_t m = static_cast<int64_t>(yt) * BN + c + j;
                if (m < p.y_rows) {
                  if (p.out_f32)
                    reinterpret_cast<float*>(p.C)[m * p.ldc + nn] = to_f32<T>(from_f32<T>(v[j]));
                  else
                    reinterpret_cast<T*>(p.C)[m * p.ldc + nn] = from_f32<T>(v[j]);
                }
              }
            }
          }
        }
        if (!complete && et == 0) p.counters[tile] = 0;  // re-arm for the next launch / graph replay
      }
      if (complete) {
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[as]);
      }
    }
  }

  griddep_launch_dependents();
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

constexpr int64_t kGemmWsBytes = 48ll << 20;  // stream-K partials: 2 per CTA
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
  // than ~8 CTAs (the fix-up reads one partial per contributor) or a CTA gets < 4 k-blocks
  // tuning knobs (experiments only): SPD_GEMM_STREAMK=0 disables stream-K, SPD_GEMM_SPLIT=d caps
  // the contributors per tile at ~d, SPD_GEMM_SWAPMAX=m is the largest M routed to swap-AB
  // Measured on B200 (scripts/bench_gemm_decode.py, profiles/r01_gemm_decode_sweep.log): one SM
  // streams ~90-100 GB/s, the global-memory fix-up of a shared tile costs ~8 us, so K is only
  // split when few tiles would each stream a long K (down_proj: 32 tiles x 224 k-blocks).
  static const int env_streamk = getenv("SPD_GEMM_STREAMK") ? atoi(getenv("SPD_GEMM_STREAMK")) : 1;
  static const int env_split = getenv("SPD_GEMM_SPLIT") ? atoi(getenv("SPD_GEMM_SPLIT")) : 4;
  static const int env_swapmax = getenv("SPD_GEMM_SWAPMAX") ? atoi(getenv("SPD_GEMM_SWAPMAX")) : 128;
  const bool ws_ok = env_streamk && workspace && workspace_bytes >= spd_gemm_workspace_bytes();
  const bool can_streamk = ws_ok && (env_streamk == 2 || kblocks >= 128);
  auto streamk_grid = [&](long long tiles, int BN) -> int {
    const long long units = tiles * kblocks;
    long long g = sms;
    const long long min_units = kblocks / env_split > 4 ? kblocks / env_split : 4;
    if (units / g < min_units) g = units / min_units > 0 ? units / min_units : 1;
    if (g * 2 * kBM * BN * 4 > kGemmWsBytes) g = kGemmWsBytes / (2ll * kBM * BN * 4);
    return static_cast<int>(g);
  };

  // y tiles per rasterisation panel: keep ~40 MB of W rows hot in L2 (tile_coords)
  static const int env_panel_mb = getenv("SPD_GEMM_PANEL_MB") ? atoi(getenv("SPD_GEMM_PANEL_MB")) : 40;
  auto panel_for = [&](int BN) -> int {
    const long long tile_bytes = static_cast<long long>(BN) * K * 2;
    long long q = (static_cast<long long>(env_panel_mb) << 20) / tile_bytes;
    return static_cast<int>(q < 1 ? 1 : q);
  };

  // M <= 128: swap-AB (the token dimension feeds UMMA-N, padded to 16..128).  Beyond that the
  // re-read of the activation tile by every weight tile dominates, so 128 < M <= 256 runs
  // un-swapped 128 x 256 tiles (activation traffic = half the weight traffic).
  const bool swap = (M <= env_swapmax) && epilogue == SPD_EPI_NONE;
  if (swap) {
    // few weight tiles (W_qkv, W_o, W_down): split K over a thread-block cluster, reduce in DSMEM
    rc = gemm_cluster_try(C, A, W, M, N, K, lda, ldw, ldc, dtype, out_f32, sms, st);
    if (rc != 1) return rc;
    int BN = 16;
    while (BN < M) BN <<= 1;  // 16,32,64 (128, 256 only when SPD_GEMM_SWAPMAX raises the limit)
    p.x_rows = N;
    p.y_rows = M;
    p.x_tiles = static_cast<int>((N + kBM - 1) / kBM);
    p.y_tiles = 1;
    int grid = static_cast<int>(p.x_tiles < sms ? p.x_tiles : sms);
    if (can_streamk && p.x_tiles <= kGemmCounters && (env_streamk == 2 || p.x_tiles * 2 <= sms)) {
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
      SPD_GEMM_SWAP(128)
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
    if (M <= 256 && use256 && can_streamk && total <= kGemmCounters &&
        (env_streamk == 2 || total * 2 <= sms)) {
      grid = streamk_grid(total, BN);  // decode regime, un-swapped
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

}  // extern "C"
