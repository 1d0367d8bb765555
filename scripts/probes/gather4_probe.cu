// Probe: semantics of cp.async.bulk.tensor.2d ... tile::gather4 on sm_100a — smem placement under
// the 128-byte swizzle, out-of-bounds rows (zero fill?) and the bytes credited to the mbarrier.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/gather4_probe scripts/probes/gather4_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tm, uint16_t* out, int oob_row, int* flag, int expect) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(base + 4096);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 1024; ++i) ((uint32_t*)base)[i] = 0xdeadbeefu;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    if (threadIdx.x == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(expect) : "memory");
    __syncwarp(0x3u);
    int r0, r1, r2, r3;
    if (threadIdx.x == 0) { r0 = 5; r1 = 63; r2 = oob_row; r3 = 2; }
    else { r0 = 7; r1 = 8; r2 = 9; r3 = 10; }
    const uint32_t dst = smem_u32(base) + threadIdx.x * 512;
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst), "l"(&tm), "r"(smem_u32(bar)), "r"(64), "r"(r0), "r"(r1),
        "r"(r2), "r"(r3)
        : "memory");
  }
  if (threadIdx.x == 0) {
    uint32_t done = 0;
    const long long t0 = clock64();
    while (!done && clock64() - t0 < 20000000ll) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(bar)), "r"(0) : "memory");
    }
    *flag = done ? 1 : 0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = ((uint16_t*)base)[i];
}

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  const int R = 64, C = 256;
  std::vector<uint16_t> h(R * C);
  for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[r * C + c] = (uint16_t)(r * 256 + c);
  uint16_t *d, *o; int* f;
  cudaMalloc(&d, h.size() * 2); cudaMalloc(&o, 1024); cudaMalloc(&f, 4);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tm;
  cuuint64_t dims[2] = {C, R};
  cuuint64_t strides[1] = {C * 2};
  cuuint32_t box[2] = {64, 1};
  cuuint32_t es[2] = {1, 1};
  CUresult rc = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)rc);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
  for (int oob : {3, 1000, -1, 0x40000000}) for (int expect : {1024, 768}) {
    if (oob == 3 && expect != 1024) continue;
    probe<<<1, 64, 8192>>>(tm, o, oob, f, expect);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<uint16_t> g(512); int flag = -1;
    cudaMemcpy(g.data(), o, 1024, cudaMemcpyDeviceToHost);
    cudaMemcpy(&flag, f, 4, cudaMemcpyDeviceToHost);
    printf("row2=%d expect_tx=%d: err=%s barrier_completed=%d\n", oob, expect, cudaGetErrorString(e), flag);
    // expected: smem row i (128 B) holds pool row rows[i], cols 64..127, 16-byte chunk c at (c ^ (i & 7))
    int rows[8] = {5, 63, oob, 2, 7, 8, 9, 10};
    int bad = 0;
    for (int i = 0; i < 8; ++i) {
      bool isoob = rows[i] < 0 || rows[i] >= R;
      for (int c = 0; c < 8; ++c) for (int e2 = 0; e2 < 8; ++e2) {
        uint16_t got = g[i * 64 + ((c ^ (i & 7)) * 8) + e2];
        uint16_t want = isoob ? 0 : (uint16_t)(rows[i] * 256 + 64 + c * 8 + e2);
        if (got != want) { if (bad < 6) printf("  mismatch smem row %d chunk %d elem %d: got %04x want %04x\n", i, c, e2, got, want); ++bad; }
      }
    }
    printf("  mismatches: %d  (row2 first words: %04x %04x %04x %04x)\n", bad, g[2 * 64], g[2 * 64 + 1], g[2 * 64 + 2], g[2 * 64 + 3]);
  }
  return 0;
}
