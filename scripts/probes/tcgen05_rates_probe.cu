// Probe: sustained rates of the tcgen05 building blocks the prefill attention is made of, one CTA
// on one SM (sm_100a):
//   * tcgen05.mma kind::f16, SS form, M = 128, K = 16, N = 64 / 128 / 256      (clk per instruction)
//   * the TS form (A operand from TMEM) at N = 128
//   * tcgen05.ld 32x32b.x32 from 4 warps                                         (bytes per clk)
//   * MUFU.EX2 from 4 and 8 warps                                                (clk per warp instruction)
//   * the MMA stream and the tcgen05.ld stream TOGETHER (do they share a port?)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -I semi_pd_b200/csrc -o /tmp/tc_probe scripts/probes/tcgen05_rates_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include "ptx.cuh"

using namespace spd;

__device__ __forceinline__ void umma_ts_probe(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// mode: 0 = MMA SS only, 1 = MMA TS only, 2 = LD only, 3 = MMA SS + LD, 4 = MMA TS + LD
template <int N>
__global__ void __launch_bounds__(192, 1) probe(int mode, int reps, long long* out) {
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
  const bool do_mma = mode != 2, do_ld = mode >= 2, ts = (mode == 1 || mode == 4);
  if (warp == 4 && do_mma) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc<__nv_bfloat16>(128, N, 0, 0);
      const uint64_t ad = umma_desc_sw128(smem_u32(smem), 16, 1024);
      const uint64_t bd = umma_desc_sw128(smem_u32(smem + 32768), 16, 1024);
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (ts) umma_ts_probe(tm, tm + 256 + (r & 7) * 8, bd, idesc, 1u);
        else umma_ss(tm, ad, bd, idesc, 1u);
      }
      umma_commit(bar);
      mbar_wait(bar, 0);
      out[0] = clock64() - t0;
    }
  } else if (warp < 4 && do_ld) {
    const uint32_t base = tm + (static_cast<uint32_t>(warp * 32) << 16) + 256;  // columns 256..383: not the MMA's D
    uint32_t acc = 0;
    __syncwarp();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(base + (r & 3) * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= v[i];
    }
    const long long t1 = clock64();
    if (lane == 0) out[1 + warp] = t1 - t0;
    if (acc == 0x12345678u) out[7] = acc;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tm, 512);
}

__global__ void mufu_probe(int reps, long long* out, float* sink) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = 0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) out[threadIdx.x >> 5] = t1 - t0;
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 123.f) *sink = s;
}

template <int N>
void run(const char* name, long long* d) {
  cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int reps = 4000;
  const char* modes[5] = {"MMA SS alone", "MMA TS alone", "tcgen05.ld alone", "MMA SS + tcgen05.ld", "MMA TS + tcgen05.ld"};
  for (int mode = 0; mode < 5; ++mode) {
    long long h[8] = {0};
    cudaMemset(d, 0, 64);
    probe<N><<<1, 192, 100 * 1024>>>(mode, reps, d);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("%-8s %-22s err=%s", name, modes[mode], cudaGetErrorString(e));
    if (mode != 2) printf("  MMA %.1f clk/instr (%.0f%% of the 8192 FLOP/clk peak)", (double)h[0] / reps,
                          100.0 * (2.0 * 128 * N * 16 / 8192) / ((double)h[0] / reps));
    if (mode >= 2) printf("  LD %.1f clk per x32 load per warp -> %.0f B/clk/SM", (double)h[1] / reps,
                          4.0 * 32 * 32 * 4 / ((double)h[1] / reps));
    printf("\n");
  }
}

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  long long* d; float* sink;
  cudaMalloc(&d, 1024); cudaMalloc(&sink, 4);
  run<128>("N=128", d);
  run<64>("N=64", d);
  run<256>("N=256", d);
  for (int warps : {4, 8, 16}) {
    long long h[16] = {0};
    mufu_probe<<<1, warps * 32>>>(2000, d, sink);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
    printf("MUFU.EX2 %2d warps: %.2f clk per warp instruction per warp -> %.1f ex2/clk/SM\n", warps,
           (double)h[0] / (2000 * 8), warps * 32.0 / ((double)h[0] / (2000 * 8)));
  }
  return 0;
}
