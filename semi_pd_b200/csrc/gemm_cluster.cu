// Decode-regime GEMM with split-K across a THREAD-BLOCK CLUSTER, reduced on chip through
// distributed shared memory — for the layers whose N gives fewer 128-row weight tiles than there
// are SMs (W_qkv: 48 tiles, W_o and W_down: 32 tiles on Llama-3-8B).  One SM streams ~90-100 GB/s
// however deep its TMA ring is, so those layers need 3-4 SMs per tile to approach HBM speed; a
// global-memory split-K fix-up costs ~8 us (measured, profiles/r01_gemm_decode_sweep.log), more
// than the 10-20 us the whole GEMM should take.
//
// cluster = S CTAs working on the same [128 weight rows] x [BN tokens] output tile (swap-AB as in
// gemm.cu), rank r owning k-blocks [r*kb/S, (r+1)*kb/S).  When its main loop is done
//   * a peer (r > 0) waits for the leader's "go" (the leader's TMA ring is idle), writes its fp32
//     accumulator tile straight into the LEADER's shared memory (st.shared::cluster) and arrives on
//     the leader's mbarrier with release.cluster semantics;
//   * the leader adds its own TMEM accumulator and the S-1 staged partials in rank order
//     (bit-deterministic), converts and stores C.
// No workspace, no atomics, no second kernel; the hardware co-schedules the cluster, so there is
// nothing to spin on.
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace spd {

constexpr int kCgThreads = 256;
constexpr int kCgBM = 128;
constexpr int kCgBK = 64;
constexpr int kCgSmemBudget = 220 * 1024;

struct ClusterGemmParams {
  void* C;
  int64_t ldc;
  int64_t n_rows;  // N (weight rows)
  int64_t m_rows;  // M (tokens)
  int kblocks;
  int out_f32;
};

template <int BN>
__host__ __device__ constexpr int cg_stage_bytes() { return (kCgBM + BN) * kCgBK * 2; }
template <int BN>
__host__ __device__ constexpr int cg_stages() {
  return (kCgSmemBudget / cg_stage_bytes<BN>()) < 12 ? (kCgSmemBudget / cg_stage_bytes<BN>()) : 12;
}
template <int BN>
__host__ __device__ constexpr int cg_smem_bytes() { return cg_stages<BN>() * cg_stage_bytes<BN>() + 1024 + 256; }
// staged partial: [128 rows][BN floats + 4 pad] — the pad rotates rows over the smem banks
template <int BN>
__host__ __device__ constexpr int cg_part_bytes() { return kCgBM * (BN * 4 + 16); }

template <typename T, int BN>
__global__ void __launch_bounds__(kCgThreads, 1)
gemm_cluster_splitk_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA,
                           const ClusterGemmParams p) {
  constexpr int kStages = cg_stages<BN>();
  constexpr int kXBytes = kCgBM * kCgBK * 2;
  constexpr int kYBytes = BN * kCgBK * 2;
  constexpr int kTmemCols = BN < 32 ? 32 : BN;
  constexpr int kChunk = BN >= 32 ? 32 : 16;
  constexpr int kRowBytes = BN * 4 + 16;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sX = smem;
  uint8_t* sY = smem + kStages * kXBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * (kXBytes + kYBytes));
  uint64_t* full = bars;              // [kStages]
  uint64_t* empty = bars + kStages;   // [kStages]
  uint64_t* tfull = bars + 2 * kStages;
  uint64_t* go = bars + 2 * kStages + 1;    // peer: the leader's ring is idle, staging may be written
  uint64_t* done = bars + 2 * kStages + 2;  // leader: all peers' partials have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 3);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t S = cluster_nctarank();
  const uint32_t rank = cluster_ctarank();
  const int tile = blockIdx.x / S;
  const int kb0 = static_cast<int>(static_cast<long long>(rank) * p.kblocks / S);
  const int kb1 = static_cast<int>(static_cast<long long>(rank + 1) * p.kblocks / S);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmA);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    mbar_init(go, 1);
    mbar_init(done, 128 * (S - 1));
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  cluster_sync_all();  // every CTA's barriers are initialised before any remote arrive
  const uint32_t tmem_base = *tmem_slot;

  // PDL at both ends (see gemm_pair_decode.cu): dependents are released at once, and the ring is
  // filled with weight tiles before this grid waits for its predecessor
  griddep_launch_dependents();
  if (warp != 0) griddep_wait();

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0, phase = 0;
      int ahead = 0;
      for (int kb = kb0; kb < kb1 && ahead < kStages; ++kb, ++ahead) {
        mbar_expect_tx(&full[ahead], kXBytes + kYBytes);
        tma_load_2d(sX + ahead * kXBytes, &tmW, &full[ahead], kb * kCgBK, tile * kCgBM, kEvictFirst);
      }
      griddep_wait();
      for (int kb = kb0; kb < kb1; ++kb) {
        if (kb - kb0 >= ahead) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], kXBytes + kYBytes);
          tma_load_2d(sX + stage * kXBytes, &tmW, &full[stage], kb * kCgBK, tile * kCgBM, kEvictFirst);
        }
        tma_load_2d(sY + stage * kYBytes, &tmA, &full[stage], kb * kCgBK, 0, kEvictLast);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc<T>(kCgBM, BN, 0, 0);
      int stage = 0, phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after_sync();
        const uint32_t xa = smem_u32(sX + stage * kXBytes);
        const uint32_t ya = smem_u32(sY + stage * kYBytes);
#pragma unroll
        for (int k = 0; k < kCgBK / 16; ++k)
          umma_ss(tmem_base, umma_desc_sw128(xa + k * 32, 16, 1024), umma_desc_sw128(ya + k * 32, 16, 1024),
                  idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        umma_commit(&empty[stage]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(tfull);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 128;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    mbar_wait(tfull, 0);  // all of this CTA's MMAs are complete: its TMA ring is idle too
    tc_fence_after_sync();
    if (rank != 0) {
      mbar_wait_cluster_acquire(go, 0);
      const uint32_t stage_local = smem_u32(smem) + (rank - 1) * cg_part_bytes<BN>() + row * kRowBytes;
      const uint32_t dst = map_to_rank(stage_local, 0);
#pragma unroll 1
      for (int c = 0; c < BN; c += kChunk) {
        uint32_t r[kChunk];
        if constexpr (kChunk == 32) tmem_ld_32x32b_x32(taddr + c, r);
        else tmem_ld_32x32b_x16(taddr + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < kChunk; j += 4) st_cluster_v4(dst + (c + j) * 4, r[j], r[j + 1], r[j + 2], r[j + 3]);
      }
      mbar_arrive_remote_release(map_to_rank(smem_u32(done), 0));
    } else {
      if (et < static_cast<int>(S) - 1) mbar_arrive_remote_release(map_to_rank(smem_u32(go), et + 1));
      if (S > 1) mbar_wait_cluster_acquire(done, 0);
      const int64_t nn = static_cast<int64_t>(tile) * kCgBM + row;
#pragma unroll 1
      for (int c = 0; c < BN; c += kChunk) {
        uint32_t r[kChunk];
        if constexpr (kChunk == 32) tmem_ld_32x32b_x32(taddr + c, r);
        else tmem_ld_32x32b_x16(taddr + c, r);
        tmem_ld_wait();
        float v[kChunk];
#pragma unroll
        for (int j = 0; j < kChunk; ++j) v[j] = __uint_as_float(r[j]);
        for (uint32_t s = 1; s < S; ++s) {  // fixed order: rank 0 (own), 1, 2, ...
          const uint8_t* ps = smem + (s - 1) * cg_part_bytes<BN>() + row * kRowBytes + c * 4;
#pragma unroll
          for (int j = 0; j < kChunk; j += 4) {
            const float4 f = *reinterpret_cast<const float4*>(ps + j * 4);
            v[j] += f.x; v[j + 1] += f.y; v[j + 2] += f.z; v[j + 3] += f.w;
          }
        }
        if (nn < p.n_rows) {
          const int m_left = static_cast<int>(p.m_rows) - c;  // tokens of this chunk that exist
          if (p.out_f32) {
            float* dst = reinterpret_cast<float*>(p.C) + static_cast<int64_t>(c) * p.ldc + nn;
#pragma unroll
            for (int j = 0; j < kChunk; ++j)
              if (j < m_left) dst[static_cast<int64_t>(j) * p.ldc] = to_f32<T>(from_f32<T>(v[j]));
          } else {
            T* dst = reinterpret_cast<T*>(p.C) + static_cast<int64_t>(c) * p.ldc + nn;
#pragma unroll
            for (int j = 0; j < kChunk; ++j)
              if (j < m_left) dst[static_cast<int64_t>(j) * p.ldc] = from_f32<T>(v[j]);
          }
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // nobody leaves while a peer may still touch its shared memory / barriers
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

template <typename T, int BN>
static int launch_cluster_gemm(const CUtensorMap& tw, const CUtensorMap& ta, const ClusterGemmParams& p,
                               int tiles, int S, cudaStream_t stream) {
  auto kern = gemm_cluster_splitk_kernel<T, BN>;
  static bool configured = false;
  if (!configured) {
    SPD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, cg_smem_bytes<BN>()));
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tiles * S);
  cfg.blockDim = dim3(kCgThreads);
  cfg.dynamicSmemBytes = cg_smem_bytes<BN>();
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  SPD_CUDA(cudaLaunchKernelEx(&cfg, kern, tw, ta, p));
  return 0;
}

// Called by spd_gemm for the swap-AB regime.  Returns 1 if this path does not apply (caller falls
// through to the single-CTA-per-tile kernel), 0 on launch, else an error code.
int gemm_cluster_try(void* C, const void* A, const void* W, int64_t M, int64_t N, int64_t K, int64_t lda,
                     int64_t ldw, int64_t ldc, int dtype, int out_f32, int sms, cudaStream_t stream) {
  static const int enabled = getenv("SPD_GEMM_CLUSTER") ? atoi(getenv("SPD_GEMM_CLUSTER")) : 1;
  if (!enabled || M > 128) return 1;
  const int tiles = static_cast<int>((N + kCgBM - 1) / kCgBM);
  const int kblocks = static_cast<int>((K + kCgBK - 1) / kCgBK);
  int BN = 16;
  while (BN < M) BN <<= 1;
  int S = sms / tiles;            // CTAs per tile that still fit the partition in one wave
  S = S >= 4 ? 4 : (S >= 2 ? 2 : 1);  // clusters of 3 schedule poorly on the 16-20-SM GPCs (measured)
  if (S > kblocks / 8) S = kblocks / 8;   // >= 8 k-blocks per rank
  const int stage_area = (BN == 16 ? cg_stages<16>() * cg_stage_bytes<16>()
                          : BN == 32 ? cg_stages<32>() * cg_stage_bytes<32>()
                          : BN == 64 ? cg_stages<64>() * cg_stage_bytes<64>()
                                     : cg_stages<128>() * cg_stage_bytes<128>());
  const int part = kCgBM * (BN * 4 + 16);
  while (S > 1 && (S - 1) * part > stage_area) --S;
  if (S < 2) return 1;
  ClusterGemmParams p{};
  p.C = C;
  p.ldc = ldc;
  p.n_rows = N;
  p.m_rows = M;
  p.kblocks = kblocks;
  p.out_f32 = out_f32;
  CUtensorMap tw, ta;
  int rc;
  if ((rc = get_tmap(&tw, W, N, K, ldw, kCgBM))) return rc;
  if ((rc = get_tmap(&ta, A, M, K, lda, BN))) return rc;
#define SPD_CG(BN_)                                                                              \
  case BN_:                                                                                      \
    SPD_DISPATCH_DTYPE(dtype, { return launch_cluster_gemm<T, BN_>(tw, ta, p, tiles, S, stream); }); \
    break;
  switch (BN) {
    SPD_CG(16)
    SPD_CG(32)
    SPD_CG(64)
    SPD_CG(128)
  }
#undef SPD_CG
  return 1;
}

}  // namespace spd
