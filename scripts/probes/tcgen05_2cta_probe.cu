// Probe: cost per instruction of tcgen05.mma.cta_group::2 (M = 256 over a CTA pair, K = 16) in the
// SS form (A and B in shared memory; B split between the two CTAs) and the TS form (A in TMEM),
// N = 64 / 128 / 256, against the single-CTA numbers of tcgen05_rates_probe.cu.  Question: does
// the pair form hide the fixed ~43 clk the single-CTA SS form pays per instruction?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -I semi_pd_b200/csrc -o gpurun_out/tc2_probe scripts/probes/tcgen05_2cta_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include "ptx.cuh"

using namespace spd;

__device__ __forceinline__ void p_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void p_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void p_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void p_mma_ss2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void p_mma_ts2(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void p_commit2(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// mode 0: SS, one accumulator; 1: TS, one accumulator; 2: SS alternating two accumulators and two
// A / B addresses (independent instruction streams, as the attention kernel issues them)
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1) probe2(int mode, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 4) {
    if (lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    __syncwarp();
    p_alloc2(slot, 512);
    p_relinquish2();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tm = *slot;
  if (warp == 4 && rank == 0) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc<__nv_bfloat16>(256, N, 0, 0);
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        const int alt = (mode == 2) ? (r & 1) : 0;
        const uint64_t ad = umma_desc_sw128(smem_u32(smem) + alt * 16384 + (r & 3) * 32, 16, 1024);
        const uint64_t bd = umma_desc_sw128(smem_u32(smem + 32768) + alt * 16384 + (r & 3) * 32, 16, 1024);
        if (mode == 1) p_mma_ts2(tm, tm + 256 + (r & 7) * 8, bd, idesc, 1u);
        else p_mma_ss2(tm + alt * (N < 256 ? N : 0), ad, bd, idesc, 1u);
      }
      p_commit2(bar, 0b11);
      mbar_wait(bar, 0);
      out[0] = clock64() - t0;
    }
  } else if (warp == 4 && rank == 1) {
    if (elect_one()) mbar_wait(bar, 0);  // the multicast commit also lands here
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 4) p_dealloc2(tm, 512);
}

// single-CTA reference with the same alternating pattern (mode 2 of the pair probe)
template <int N>
__global__ void __launch_bounds__(192, 1) probe1_alt(int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 4) {
    if (lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = *slot;
  if (warp == 4 && elect_one()) {
    constexpr uint32_t idesc = umma_idesc<__nv_bfloat16>(128, N, 0, 0);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const int alt = r & 1;
      const uint64_t ad = umma_desc_sw128(smem_u32(smem) + alt * 16384 + (r & 3) * 32, 16, 1024);
      const uint64_t bd = umma_desc_sw128(smem_u32(smem + 32768) + alt * 16384 + (r & 3) * 32, 16, 1024);
      umma_ss(tm + alt * (N < 256 ? N : 0), ad, bd, idesc, 1u);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    out[0] = clock64() - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tm, 512);
}

template <int N>
void run(long long* d) {
  const int reps = 4000;
  cudaFuncSetAttribute(probe2<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(probe1_alt<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const char* modes[3] = {"pair SS", "pair TS", "pair SS alternating"};
  for (int mode = 0; mode < 3; ++mode) {
    long long h = 0;
    cudaMemset(d, 0, 64);
    probe2<N><<<2, 192, 100 * 1024>>>(mode, reps, d);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("N=%-4d %-22s err=%s  %.1f clk/instr (ideal %d -> %.0f%% of 8192 FLOP/clk/SM)\n", N, modes[mode],
           cudaGetErrorString(e), (double)h / reps, N / 2, 100.0 * (N / 2) / ((double)h / reps));
  }
  long long h = 0;
  cudaMemset(d, 0, 64);
  probe1_alt<N><<<1, 192, 100 * 1024>>>(reps, d);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("N=%-4d %-22s err=%s  %.1f clk/instr (ideal %d -> %.0f%%)\n", N, "single SS alternating", cudaGetErrorString(e),
         (double)h / reps, N / 2, 100.0 * (N / 2) / ((double)h / reps));
}

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  long long* d;
  cudaMalloc(&d, 1024);
  run<64>(d);
  run<128>(d);
  run<256>(d);
  return 0;
}
