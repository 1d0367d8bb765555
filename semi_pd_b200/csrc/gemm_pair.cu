// Prefill-regime GEMM on CTA PAIRS: C[M,N] = A[M,K] * W[N,K]^T with tcgen05.mma.cta_group::2.
//
// Two CTAs on the two SMs of one TPC (cluster of 2) own one 256 x 256 output tile.  CTA r holds
// the A rows [r*128, r*128+128) of the tile and HALF of the W tile (128 of its 256 rows); one
// UMMA of shape M=256, N=256, K=16 issued by the leader CTA reads both CTAs' shared memory and
// writes each CTA's 128 x 256 fp32 accumulator into that CTA's own TMEM.  Against the single-CTA
// 128 x 256 tile (gemm.cu) every CTA stages 32 KB instead of 48 KB per k-block and the tensor
// core reads a third fewer operand bytes from shared memory per FLOP — on a power-capped B200
// (scripts/bench_gemm_prefill.py: ~1 kW, SM clock 1.15-1.45 GHz under sustained GEMM load) the
// joules per FLOP are what sets the sustained rate.
//
// Pipeline (per CTA: 256 threads; warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA only),
// warp 2 = TMEM allocator, warps 4-7 = epilogue):
//   full[s]   leader's barrier: 1 arrive (leader producer, expect_tx of BOTH CTAs' 64 KB) + the
//             transaction bytes of the four TMA loads (two per CTA, .cta_group::2 so the peer's
//             loads may signal the leader's barrier)
//   empty[s]  one per CTA, count 1: tcgen05.commit.cta_group::2 ... multicast to both CTAs
//   tfull[a]  one per CTA, count 1: multicast commit when an accumulator stage is complete
//   tempty[a] leader's barrier, count 8: the four epilogue warps of each CTA (the peer's arrive
//             remotely with release.cluster)
// The fused SiLU*mul epilogue maps naturally: CTA 0 stages the 128 gate rows, CTA 1 the 128 up
// rows of the tile, so accumulator columns [0,128) are gate and [128,256) are up in both CTAs.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace spd {

constexpr int kPgThreads = 256;
constexpr int kPgBM = 128;   // rows per CTA (pair tile = 256)
constexpr int kPgBN = 256;
constexpr int kPgBK = 64;
constexpr int kPgStages = 6;
constexpr int kPgABytes = kPgBM * kPgBK * 2;        // 16 KB
constexpr int kPgBBytes = (kPgBN / 2) * kPgBK * 2;  // 16 KB: this CTA's half of the W tile
constexpr int kPgSmemBytes = kPgStages * (kPgABytes + kPgBBytes) + 1024 + 256;

struct PairGemmParams {
  void* C;
  int64_t ldc;
  int64_t m_rows;    // M
  int64_t n_cols;    // output columns (SILU_MUL: the intermediate size)
  int64_t up_row0;   // SILU_MUL: row of W where the "up" half starts
  int x_tiles, y_tiles, kblocks;
  int panel;         // y tiles per L2 rasterisation panel (see gemm.cu: tile_coords)
  int out_f32;
};

__device__ __forceinline__ void pair_tile_coords(const PairGemmParams& p, int tile, int& xt, int& yt) {
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

// --- cta_group::2 forms of the tcgen05 / TMA instructions ---
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// the mbarrier operand is a shared::cluster address (the LEADER's full barrier)
__device__ __forceinline__ void tma_load_2d_2cta(void* smem_dst, const void* tmap, uint32_t bar_cluster_addr,
                                                 int32_t c_inner, int32_t c_outer, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(bar_cluster_addr), "r"(c_inner), "r"(c_outer), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void umma_ss_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in every CTA of `mask` once all MMAs issued
// so far by this thread have completed
__device__ __forceinline__ void umma_commit_2cta_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

template <typename T, int kEpi>
__global__ void __launch_bounds__(kPgThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                 const PairGemmParams p) {
  constexpr int kBNOut = (kEpi == SPD_EPI_SILU_MUL) ? kPgBN / 2 : kPgBN;
  constexpr int kTmemCols = 2 * kPgBN;  // two accumulator stages

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kPgStages * kPgABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPgStages * (kPgABytes + kPgBBytes));
  uint64_t* full = bars;                        // [kPgStages]  (leader's are the live ones)
  uint64_t* empty = bars + kPgStages;           // [kPgStages]
  uint64_t* tfull = bars + 2 * kPgStages;       // [2]
  uint64_t* tempty = bars + 2 * kPgStages + 2;  // [2]          (leader's are the live ones)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kPgStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  const int num_tiles = p.x_tiles * p.y_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kPgStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta(tmem_slot, kTmemCols);
    tmem_relinquish_2cta();
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers exist before any cross-CTA signal
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  griddep_wait();

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0, phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        int xt, yt;
        pair_tile_coords(p, tile, xt, yt);
        const int a_row = xt * (2 * kPgBM) + static_cast<int>(rank) * kPgBM;
        const int w_row = (kEpi == SPD_EPI_SILU_MUL)
                              ? (rank == 0 ? yt * kBNOut : static_cast<int>(p.up_row0) + yt * kBNOut)
                              : yt * kPgBN + static_cast<int>(rank) * (kPgBN / 2);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (rank == 0) mbar_expect_tx(&full[stage], 2 * (kPgABytes + kPgBBytes));
          const uint32_t lead_full = map_to_rank(smem_u32(&full[stage]), 0);
          tma_load_2d_2cta(sA + stage * kPgABytes, &tmA, lead_full, kb * kPgBK, a_row, kEvictNormal);
          tma_load_2d_2cta(sB + stage * kPgBBytes, &tmW, lead_full, kb * kPgBK, w_row, kEvictNormal);
          if (++stage == kPgStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = umma_idesc<T>(2 * kPgBM, kPgBN, 0, 0);
      int stage = 0, phase = 0, n = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        const int as = n & 1, aphase = (n >> 1) & 1;
        ++n;
        mbar_wait_cluster_acquire(&tempty[as], aphase ^ 1);  // both CTAs drained this accumulator
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * kPgBN;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after_sync();
          const uint32_t aa = smem_u32(sA + stage * kPgABytes);
          const uint32_t ba = smem_u32(sB + stage * kPgBBytes);
#pragma unroll
          for (int k = 0; k < kPgBK / 16; ++k)
            umma_ss_2cta(d_tmem, umma_desc_sw128(aa + k * 32, 16, 1024), umma_desc_sw128(ba + k * 32, 16, 1024),
                         idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_2cta_mc(&empty[stage], 0b11);  // both CTAs' slots are free once these MMAs are done
          if (++stage == kPgStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_2cta_mc(&tfull[as], 0b11);
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lead_tempty0 = map_to_rank(smem_u32(&tempty[0]), 0);
    int n = 0;
    for (int tile = pair; tile < num_tiles; tile += npairs) {
      int xt, yt;
      pair_tile_coords(p, tile, xt, yt);
      const int as = n & 1, aphase = (n >> 1) & 1;
      ++n;
      mbar_wait(&tfull[as], aphase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kPgBN;
      const int64_t m = static_cast<int64_t>(xt) * (2 * kPgBM) + rank * kPgBM + row;
#pragma unroll 1
      for (int c = 0; c < kBNOut; c += 32) {
        uint32_t r[32];
        float v[32];
        tmem_ld_32x32b_x32(taddr + c, r);
        if constexpr (kEpi == SPD_EPI_SILU_MUL) {
          uint32_t u[32];
          tmem_ld_32x32b_x32(taddr + kBNOut + c, u);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            // gate_up_proj output is rounded to T before SiluAndMul reads it
            const float g = to_f32<T>(from_f32<T>(__uint_as_float(r[j])));
            const float up = to_f32<T>(from_f32<T>(__uint_as_float(u[j])));
            v[j] = silu_f32(g) * up;
          }
        } else {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        }
        const int64_t n0 = static_cast<int64_t>(yt) * kBNOut + c;
        if (m < p.m_rows) {
          if (p.out_f32) {
            float* dst = reinterpret_cast<float*>(p.C) + m * p.ldc + n0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (n0 + j < p.n_cols) {
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
            for (int j = 0; j < 32; j += 8) {
              if (n0 + j < p.n_cols) {
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
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote_release(lead_tempty0 + as * 8);
    }
  }

  griddep_launch_dependents();
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // the peer may still read this CTA's shared memory / signal its barriers
  if (warp == 2) tmem_dealloc_2cta(tmem_base, kTmemCols);
}

template <typename T, int kEpi>
static int launch_pair_gemm(const CUtensorMap& ta, const CUtensorMap& tw, const PairGemmParams& p, int pairs,
                            cudaStream_t stream) {
  auto kern = gemm_pair_kernel<T, kEpi>;
  static bool configured = false;
  if (!configured) {
    SPD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPgSmemBytes));
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kPgThreads);
  cfg.dynamicSmemBytes = kPgSmemBytes;
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
  SPD_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tw, p));
  return 0;
}

// Called by spd_gemm for the prefill regime.  Returns 1 if this path does not apply (the caller
// falls through to the single-CTA kernel), 0 on launch, else an error code.
int gemm_pair_try(void* C, const void* A, const void* W, int64_t M, int64_t N, int64_t K, int64_t lda,
                  int64_t ldw, int64_t ldc, int dtype, int out_f32, int epilogue, int sms, int panel_mb,
                  cudaStream_t stream) {
  static const int enabled = getenv("SPD_GEMM_2CTA") ? atoi(getenv("SPD_GEMM_2CTA")) : 1;
  if (!enabled || M <= 256 || sms < 2) return 1;
  // a 256-row pair tile must not pad M more than the 128-row tiles would, unless M is large
  const int64_t rows_pair = (M + 255) / 256 * 256, rows_single = (M + 127) / 128 * 128;
  if (rows_pair != rows_single && M < 2048) return 1;
  PairGemmParams p{};
  p.C = C;
  p.ldc = ldc;
  p.m_rows = M;
  p.n_cols = N;
  p.up_row0 = N;
  p.out_f32 = out_f32;
  p.kblocks = static_cast<int>((K + kPgBK - 1) / kPgBK);
  p.x_tiles = static_cast<int>((M + 255) / 256);
  const int n_out = epilogue == SPD_EPI_SILU_MUL ? kPgBN / 2 : kPgBN;
  p.y_tiles = static_cast<int>((N + n_out - 1) / n_out);
  const long long tile_bytes = static_cast<long long>(kPgBN) * K * 2;
  const long long q = (static_cast<long long>(panel_mb) << 20) / tile_bytes;
  p.panel = static_cast<int>(q < 1 ? 1 : q);
  const int total = p.x_tiles * p.y_tiles;
  const int pairs = total < sms / 2 ? total : sms / 2;
  CUtensorMap ta, tw;
  int rc;
  if ((rc = get_tmap(&ta, A, M, K, lda, kPgBM))) return rc;
  if ((rc = get_tmap(&tw, W, epilogue == SPD_EPI_SILU_MUL ? 2 * N : N, K, ldw, kPgBN / 2))) return rc;
  if (epilogue == SPD_EPI_SILU_MUL) {
    SPD_DISPATCH_DTYPE(dtype, { return launch_pair_gemm<T, SPD_EPI_SILU_MUL>(ta, tw, p, pairs, stream); });
  } else {
    SPD_DISPATCH_DTYPE(dtype, { return launch_pair_gemm<T, SPD_EPI_NONE>(ta, tw, p, pairs, stream); });
  }
  return 1;
}

}  // namespace spd
