// Tensor-parallel all-reduce over NVLink 5 / NVSwitch peer memory (one process per GPU).
//
// Replaces the custom all-reduce the reference dispatches its row-parallel outputs to when the
// message is below 8 MiB (python/sglang/srt/distributed/device_communicators/
// custom_all_reduce.py:147-160,447-480 → vLLM `_C`; TRT-LLM variant
// sgl-kernel/csrc/allreduce/trt_reduce_internal.cu:190-300).  Same contract — SUM over the TP
// ranks of a 16-bit [T, hidden] tensor, usable inside a captured CUDA graph — but built for the
// single-process-per-GPU layout of this repo:
//
//   * every rank owns ONE cudaMalloc'ed symmetric buffer [flags | counters | slot0 | slot1 | res],
//     exported with cudaIpcGetMemHandle and mapped by all peers (spd_comm_create / _connect; the
//     64-byte handles travel through torch.distributed, the plumbing);
//   * the kernel copies its input into the slot of the call's parity, signals the peers
//     (st.release.sys into THEIR flag words), waits for their signals (ld.acquire.sys on its own)
//     and then either reduces all peers' slots directly (one-shot, latency-bound messages) or
//     reduces only its 1/W partition, publishes it and gathers the other partitions (two-shot);
//   * all state a call needs (barrier sequence number, slot parity) lives in device memory and
//     is advanced by the kernel itself, so a captured graph replays correctly any number of
//     times; no host-side sequence numbers, no per-call registration of buffers;
//   * every rank sums the W inputs in rank order 0..W-1 in fp32 and rounds once: the result is
//     bit-identical on all ranks and run to run;
//   * a spin that exceeds the time-out (a dead peer) raises a sticky error flag instead of
//     hanging the GPU (spd_comm_status).
//
// Why a constant grid: block b only ever talks to block b of the peers, and the double-buffered
// slots are safe because a rank's call n+2 cannot start before ALL its blocks passed a barrier of
// call n+1, i.e. before all peer blocks finished reading call n.  That argument needs every block
// to take part in every call, whatever the message size — so the grid never changes.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <new>

#include "common.cuh"
#include "ptx.cuh"

namespace spd {

constexpr int kArMaxRanks = 8;
constexpr int kArBlocks = 32;
constexpr int kArThreads = 512;
constexpr int64_t kArFlagBytes = 4096;      // uint32 flags[kArBlocks][kArMaxRanks], written by peers
constexpr int64_t kArCounterBytes = 4096;   // uint32 bar[kArBlocks], calls[kArBlocks], error

struct ArPtrs {
  uint8_t* base[kArMaxRanks];  // every rank's symmetric buffer as mapped in THIS process
};

struct ArParams {
  ArPtrs peers;
  int rank;
  int world;
  int64_t slot_bytes;   // size of one data slot
  const void* in;
  void* out;
  int64_t count;        // elements
  uint64_t timeout_ns;
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ uint32_t* flags_of(uint8_t* base) { return reinterpret_cast<uint32_t*>(base); }
__device__ __forceinline__ uint32_t* counters_of(uint8_t* base) {
  return reinterpret_cast<uint32_t*>(base + kArFlagBytes);
}

// Cross-rank barrier among the blocks with this blockIdx: thread t < world signals peer t and
// waits for peer t's signal.  `val` is the block's barrier sequence number (monotonic, compared
// modulo 2^32).  Returns after a __syncthreads, so every thread of the block may read peer data.
__device__ __forceinline__ void block_barrier(const ArParams& p, uint32_t val) {
  __syncthreads();  // this block's stores precede the signal (release is cumulative over bar.sync)
  if (threadIdx.x < p.world) {
    const int peer = threadIdx.x;
    uint32_t* theirs = flags_of(p.peers.base[peer]) + blockIdx.x * kArMaxRanks + p.rank;
    const uint32_t* mine = flags_of(p.peers.base[p.rank]) + blockIdx.x * kArMaxRanks + peer;
    volatile uint32_t* err = counters_of(p.peers.base[p.rank]) + 2 * kArBlocks;
    __threadfence_system();
    st_release_sys(theirs, val);
    const uint64_t t0 = globaltimer_ns();
    // a peer that never signals (crashed rank) must not hang the GPU: after the time-out the
    // error word is set and stays set — later barriers do not wait at all — and the host reads
    // it through spd_comm_status
    while (static_cast<int32_t>(ld_acquire_sys(mine) - val) < 0) {
      if (*err != 0u) break;
      if (globaltimer_ns() - t0 > p.timeout_ns) {
        *err = 1u;
        break;
      }
    }
  }
  __syncthreads();
}

template <typename T>
__device__ __forceinline__ void accum8(float (&acc)[8], const int4& v) {
  const T* h = reinterpret_cast<const T*>(&v);
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] += static_cast<float>(h[i]);
}
template <typename T>
__device__ __forceinline__ int4 pack8(const float (&acc)[8]) {
  int4 v;
  T* h = reinterpret_cast<T*>(&v);
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = static_cast<T>(acc[i]);
  return v;
}

// out[i] (and res[i] if given) = sum over ranks of slot_r[i], i in [lo, hi), rank order, fp32.
// U vectors per thread are loaded from every rank before any is consumed: with one 16-byte load
// in flight per thread a block pulls ~8 KB per NVLink round trip (~3 us) — 32 blocks then move
// < 100 GB/s, which made a 64 MiB prefill all-reduce cost as much as the layer it follows.
template <typename T, int W, int U>
__device__ __forceinline__ void reduce_span(const ArParams& p, int world, int64_t slot_off, int64_t lo, int64_t hi,
                                            int4* out, int4* res) {
  for (int64_t i0 = lo + threadIdx.x; i0 < hi; i0 += static_cast<int64_t>(kArThreads) * U) {
    int4 v[U][kArMaxRanks];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + static_cast<int64_t>(u) * kArThreads;
      if (i < hi) {
#pragma unroll
        for (int r = 0; r < kArMaxRanks; ++r)
          if (r < world) v[u][r] = __ldcv(reinterpret_cast<const int4*>(p.peers.base[r] + slot_off) + i);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + static_cast<int64_t>(u) * kArThreads;
      if (i < hi) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < kArMaxRanks; ++r)
          if (r < world) accum8<T>(acc, v[u][r]);
        const int4 o = pack8<T>(acc);
        if (res) res[i] = o;
        out[i] = o;
      }
    }
  }
}

// dst[i] = src[i], i in [lo, hi), four 16-byte loads per thread in flight (src may be peer memory)
__device__ __forceinline__ void copy_span(int4* dst, const int4* src, int64_t lo, int64_t hi, bool peer) {
  for (int64_t i0 = lo + threadIdx.x; i0 < hi; i0 += static_cast<int64_t>(kArThreads) * 4) {
    int4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + static_cast<int64_t>(u) * kArThreads;
      if (i < hi) v[u] = peer ? __ldcv(src + i) : src[i];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + static_cast<int64_t>(u) * kArThreads;
      if (i < hi) dst[i] = v[u];
    }
  }
}

// Sub-range [lo, hi) (in 8-element vectors) of block b within a span of `nvec` vectors.
__device__ __forceinline__ void block_span(int64_t nvec, int64_t& lo, int64_t& hi) {
  const int64_t per = (nvec + kArBlocks - 1) / kArBlocks;
  lo = per * blockIdx.x;
  hi = lo + per < nvec ? lo + per : nvec;
  if (lo > nvec) lo = nvec;
}

template <typename T, int W, bool kTwoShot>
__global__ void __launch_bounds__(kArThreads, 1) allreduce_kernel(ArParams p) {
#if __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.wait;" ::: "memory");  // PDL: the producer GEMM's output is visible
#endif
  const int world = W > 0 ? W : p.world;
  uint8_t* self = p.peers.base[p.rank];
  uint32_t* ctr = counters_of(self);
  __shared__ uint32_t s_bar, s_call;
  if (threadIdx.x == 0) {
    s_bar = ctr[blockIdx.x];
    s_call = ctr[kArBlocks + blockIdx.x];
    ctr[blockIdx.x] = s_bar + (kTwoShot ? 2u : 1u);
    ctr[kArBlocks + blockIdx.x] = s_call + 1u;
  }
  __syncthreads();
  const uint32_t bar0 = s_bar;
  const int parity = s_call & 1u;
  const int64_t slot_off = kArFlagBytes + kArCounterBytes + parity * p.slot_bytes;
  const int64_t res_off = kArFlagBytes + kArCounterBytes + 2 * p.slot_bytes;
  const int64_t nvec = p.count / 8;
  const int4* in = reinterpret_cast<const int4*>(p.in);
  int4* out = reinterpret_cast<int4*>(p.out);
  int4* my_slot = reinterpret_cast<int4*>(self + slot_off);

  // vectors per thread in flight: U x world 16-byte loads (64 registers)
  constexpr int U = (W == 2) ? 8 : (W == 4) ? 4 : 2;
  if (!kTwoShot) {
    int64_t lo, hi;
    block_span(nvec, lo, hi);
    copy_span(my_slot, in, lo, hi, false);
    block_barrier(p, bar0 + 1u);
    reduce_span<T, W, U>(p, world, slot_off, lo, hi, out, nullptr);
  } else {
    // partition q of the message = vectors [q*pv, min((q+1)*pv, nvec))
    const int64_t pv = (nvec + world - 1) / world;
    for (int q = 0; q < world; ++q) {
      const int64_t base = q * pv;
      const int64_t len = base >= nvec ? 0 : (nvec - base < pv ? nvec - base : pv);
      int64_t lo, hi;
      block_span(len, lo, hi);
      copy_span(my_slot, in, base + lo, base + hi, false);
    }
    block_barrier(p, bar0 + 1u);
    {
      const int64_t base = p.rank * pv;
      const int64_t len = base >= nvec ? 0 : (nvec - base < pv ? nvec - base : pv);
      int64_t lo, hi;
      block_span(len, lo, hi);
      reduce_span<T, W, U>(p, world, slot_off, base + lo, base + hi, out, reinterpret_cast<int4*>(self + res_off));
    }
    block_barrier(p, bar0 + 2u);
    for (int q = 0; q < world; ++q) {
      if (q == p.rank) continue;
      const int64_t base = q * pv;
      const int64_t len = base >= nvec ? 0 : (nvec - base < pv ? nvec - base : pv);
      int64_t lo, hi;
      block_span(len, lo, hi);
      copy_span(out, reinterpret_cast<const int4*>(p.peers.base[q] + res_off), base + lo, base + hi, true);
    }
  }
#if __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

// ------------------------------------------------------------------------------------------------
// all-reduce FUSED with its consumer: x = sum over ranks of in; residual += x; out = RMSNorm(residual)
// (tensor_model_parallel_all_reduce after o_proj / down_proj, S/layers/linear.py:1265-1266, followed
// by fused_add_rmsnorm, S/layers/layernorm.py:46-57) — one kernel instead of two and no round trip
// of the reduced tensor through memory.  One-shot form: a group of 256 threads owns one token row
// (two rows per block at a time), reduces it across the W slots, adds the residual and normalises
// it in registers.  The arithmetic is that of allreduce_kernel followed by rmsnorm_kernel<fused>
// (elementwise.cu), step by step — fp32 sum in rank order rounded to T, fp32 add of the residual,
// the same per-thread accumulation order and reduction tree for the sum of squares — so the fused
// and the unfused path are bit-identical.  Same grid, counters and slots as allreduce_kernel: the
// calls of one communicator may interleave freely.
// ------------------------------------------------------------------------------------------------
struct ArNormParams {
  void* residual;       // [rows, dim] in / out
  const void* weight;   // [dim]
  int64_t rows;
  int dim;
  float eps;
};

constexpr int kArNormMaxVec = 4;  // 16-byte vectors per thread: rows up to 256 * 4 * 8 = 8192 elements

template <typename T>
__device__ __forceinline__ void ar_unpack8(const int4& v, float (&f)[8]) {
  const T* h = reinterpret_cast<const T*>(&v);
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = static_cast<float>(h[i]);
}

template <typename T, int W>
__global__ void __launch_bounds__(kArThreads, 1) allreduce_add_rmsnorm_kernel(ArParams p, ArNormParams f) {
#if __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
  const int world = W > 0 ? W : p.world;
  uint8_t* self = p.peers.base[p.rank];
  uint32_t* ctr = counters_of(self);
  __shared__ uint32_t s_bar, s_call;
  __shared__ float red[2][33];
  if (threadIdx.x == 0) {
    s_bar = ctr[blockIdx.x];
    s_call = ctr[kArBlocks + blockIdx.x];
    ctr[blockIdx.x] = s_bar + 1u;
    ctr[kArBlocks + blockIdx.x] = s_call + 1u;
  }
  __syncthreads();
  const uint32_t bar0 = s_bar;
  const int parity = s_call & 1u;
  const int64_t slot_off = kArFlagBytes + kArCounterBytes + parity * p.slot_bytes;
  const int g = threadIdx.x >> 8;   // row group of the block
  const int t = threadIdx.x & 255;  // thread within the group == thread of rmsnorm_kernel
  const int nv = f.dim >> 3;
  const int4* in = reinterpret_cast<const int4*>(p.in);
  int4* my_slot = reinterpret_cast<int4*>(self + slot_off);
  // this block's rows go into its slot region; block b of every peer holds the same rows
  for (int64_t row = blockIdx.x * 2 + g; row < f.rows; row += 2 * kArBlocks)
    for (int v = t; v < nv; v += 256) my_slot[row * nv + v] = in[row * nv + v];
  block_barrier(p, bar0 + 1u);

  int4* out = reinterpret_cast<int4*>(p.out);
  int4* res = reinterpret_cast<int4*>(f.residual);
  const int4* wv = reinterpret_cast<const int4*>(f.weight);
  for (int64_t row = blockIdx.x * 2 + g; row < f.rows; row += 2 * kArBlocks) {
    float vals[kArNormMaxVec][8];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kArNormMaxVec; ++i) {
      const int v = t + i * 256;
      if (v < nv) {
        const int64_t idx = row * nv + v;
        int4 sv[kArMaxRanks];
#pragma unroll
        for (int r = 0; r < kArMaxRanks; ++r)
          if (r < world) sv[r] = __ldcv(reinterpret_cast<const int4*>(p.peers.base[r] + slot_off) + idx);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < kArMaxRanks; ++r)
          if (r < world) accum8<T>(acc, sv[r]);
        const int4 x16 = pack8<T>(acc);  // the all-reduce's output, rounded to T as the unfused kernel stores it
        float rr[8];
        ar_unpack8<T>(x16, vals[i]);
        ar_unpack8<T>(res[idx], rr);
#pragma unroll
        for (int j = 0; j < 8; ++j) vals[i][j] += rr[j];
        res[idx] = pack8<T>(vals[i]);  // residual_out = (x + residual).to(dtype)
#pragma unroll
        for (int j = 0; j < 8; ++j) ss += vals[i][j] * vals[i][j];
      }
    }
    // sum of squares over the row: the reduction tree of elementwise.cu's block_sum (8 warps)
    ss = warp_sum(ss);
    const int gw = t >> 5;
    if ((t & 31) == 0) red[g][gw] = ss;
    asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");
    float tt = (t < 8) ? red[g][t] : 0.f;
    if (gw == 0) {
      tt = warp_sum(tt);
      if (t == 0) red[g][32] = tt;
    }
    asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");
    const float inv = rsqrtf(red[g][32] / static_cast<float>(f.dim) + f.eps);
#pragma unroll
    for (int i = 0; i < kArNormMaxVec; ++i) {
      const int v = t + i * 256;
      if (v < nv) {
        float wf[8], o[8];
        ar_unpack8<T>(wv[v], wf);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = vals[i][j] * inv * wf[j];
        out[row * nv + v] = pack8<T>(o);
      }
    }
  }
#if __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

// out[r * nvec + i] = rank r's in[i]: the all-gather of equally sized 16-byte-vector payloads
// (the [bs, 16 B] (max, argmax) rows of the vocab-sharded lm_head; small logits blocks).
__global__ void __launch_bounds__(kArThreads, 1) allgather_kernel(ArParams p) {
#if __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
  uint8_t* self = p.peers.base[p.rank];
  uint32_t* ctr = counters_of(self);
  __shared__ uint32_t s_bar, s_call;
  if (threadIdx.x == 0) {
    s_bar = ctr[blockIdx.x];
    s_call = ctr[kArBlocks + blockIdx.x];
    ctr[blockIdx.x] = s_bar + 1u;
    ctr[kArBlocks + blockIdx.x] = s_call + 1u;
  }
  __syncthreads();
  const int parity = s_call & 1u;
  const int64_t slot_off = kArFlagBytes + kArCounterBytes + parity * p.slot_bytes;
  const int64_t nvec = p.count;  // 16-byte vectors per rank
  const int4* in = reinterpret_cast<const int4*>(p.in);
  int4* out = reinterpret_cast<int4*>(p.out);
  int4* my_slot = reinterpret_cast<int4*>(self + slot_off);
  int64_t lo, hi;
  block_span(nvec, lo, hi);
  for (int64_t i = lo + threadIdx.x; i < hi; i += kArThreads) my_slot[i] = in[i];
  block_barrier(p, s_bar + 1u);
  for (int r = 0; r < p.world; ++r) {
    const int4* src = reinterpret_cast<const int4*>(p.peers.base[r] + slot_off);
    for (int64_t i = lo + threadIdx.x; i < hi; i += kArThreads) out[r * nvec + i] = __ldcv(src + i);
  }
#if __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

}  // namespace spd

using namespace spd;

struct spd_comm {
  int rank;
  int world;
  int device;
  int64_t max_bytes;      // largest message (bytes) = size of one slot
  int64_t total_bytes;
  uint8_t* local;         // this rank's symmetric buffer
  uint8_t* mapped[kArMaxRanks];
  bool connected;
  uint64_t timeout_ns;
};

extern "C" {

int spd_comm_create(int rank, int world, int64_t max_message_bytes, spd_comm_t** out,
                    void* ipc_handle_out) {
  SPD_REQUIRE(out && ipc_handle_out, SPD_E_BADARG, "spd_comm_create: null pointer");
  SPD_REQUIRE(world >= 2 && world <= kArMaxRanks && rank >= 0 && rank < world, SPD_E_BADARG,
              "spd_comm_create: bad rank/world (%d/%d; world must be 2..%d)", rank, world, kArMaxRanks);
  SPD_REQUIRE(max_message_bytes > 0 && max_message_bytes % 16 == 0, SPD_E_BADARG,
              "spd_comm_create: max_message_bytes must be a positive multiple of 16");
  static_assert(sizeof(cudaIpcMemHandle_t) == SPD_IPC_HANDLE_BYTES, "IPC handle size");
  spd_comm* c = new (std::nothrow) spd_comm();
  SPD_REQUIRE(c, SPD_E_BADARG, "spd_comm_create: out of host memory");
  memset(c, 0, sizeof(*c));
  c->rank = rank;
  c->world = world;
  c->max_bytes = max_message_bytes;
  c->total_bytes = kArFlagBytes + kArCounterBytes + 3 * max_message_bytes;
  c->timeout_ns = 10ull * 1000ull * 1000ull * 1000ull;
  if (const char* e = getenv("SPD_AR_TIMEOUT_MS")) c->timeout_ns = strtoull(e, nullptr, 10) * 1000000ull;
  cudaError_t e = cudaGetDevice(&c->device);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&c->local), c->total_bytes);
  if (e == cudaSuccess) e = cudaMemset(c->local, 0, c->total_bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, c->local);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    if (c->local) cudaFree(c->local);
    delete c;
    return set_error(static_cast<int>(e), "spd_comm_create: %s", cudaGetErrorString(e));
  }
  memcpy(ipc_handle_out, &h, sizeof(h));
  c->mapped[rank] = c->local;
  *out = c;
  return 0;
}

int spd_comm_connect(spd_comm_t* c, const void* all_handles) {
  SPD_REQUIRE(c && all_handles, SPD_E_BADARG, "spd_comm_connect: null pointer");
  SPD_REQUIRE(!c->connected, SPD_E_BADARG, "spd_comm_connect: already connected");
  const uint8_t* hs = static_cast<const uint8_t*>(all_handles);
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, hs + r * sizeof(h), sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
      return set_error(static_cast<int>(e), "spd_comm_connect: cudaIpcOpenMemHandle(rank %d): %s", r,
                       cudaGetErrorString(e));
    c->mapped[r] = static_cast<uint8_t*>(p);
  }
  c->connected = true;
  return 0;
}

int64_t spd_comm_max_bytes(const spd_comm_t* c) { return c ? c->max_bytes : 0; }

int spd_comm_status(spd_comm_t* c, int* timed_out) {
  SPD_REQUIRE(c && timed_out, SPD_E_BADARG, "spd_comm_status: null pointer");
  uint32_t v = 0;
  SPD_CUDA(cudaMemcpy(&v, c->local + kArFlagBytes + 2 * kArBlocks * sizeof(uint32_t), sizeof(v),
                      cudaMemcpyDeviceToHost));
  *timed_out = static_cast<int>(v);
  return 0;
}

int spd_allreduce(spd_comm_t* c, void* out, const void* in, int64_t count, int dtype, int algo,
                  spd_stream_t stream) {
  SPD_REQUIRE(c && out && in, SPD_E_BADARG, "spd_allreduce: null pointer");
  SPD_REQUIRE(c->connected, SPD_E_BADARG, "spd_allreduce: communicator is not connected");
  SPD_REQUIRE(count >= 0 && count % 8 == 0, SPD_E_ALIGN, "spd_allreduce: count must be a multiple of 8");
  SPD_REQUIRE(aligned16(out) && aligned16(in), SPD_E_ALIGN, "spd_allreduce: pointers must be 16-byte aligned");
  SPD_REQUIRE(count * 2 <= c->max_bytes, SPD_E_WORKSPACE,
              "spd_allreduce: message of %lld bytes exceeds the communicator's %lld",
              (long long)(count * 2), (long long)c->max_bytes);
  SPD_REQUIRE(algo >= 0 && algo <= 2, SPD_E_BADARG, "spd_allreduce: algo must be 0 (auto), 1 or 2");
  if (count == 0) return 0;
  ArParams p{};
  for (int r = 0; r < c->world; ++r) p.peers.base[r] = c->mapped[r];
  p.rank = c->rank;
  p.world = c->world;
  p.slot_bytes = c->max_bytes;
  p.in = in;
  p.out = out;
  p.count = count;
  p.timeout_ns = c->timeout_ns;
  // dispatch as custom_all_reduce.py / the vLLM kernel it wraps: one-shot for latency-bound
  // messages, two-shot (reduce-scatter + all-gather) once the W-fold read volume dominates
  const int64_t bytes = count * 2;
  bool two_shot;
  if (algo == 1) two_shot = false;
  else if (algo == 2) two_shot = true;
  else if (c->world == 2) two_shot = false;
  else if (c->world <= 4) two_shot = bytes >= 512 * 1024;
  else two_shot = bytes >= 256 * 1024;
  const bool pdl = pdl_enabled();
  cudaStream_t s = as_stream(stream);
#define SPD_AR_LAUNCH(W)                                                                          \
  do {                                                                                            \
    if (two_shot)                                                                                 \
      SPD_CUDA(launch(allreduce_kernel<T, W, true>, dim3(kArBlocks), dim3(kArThreads), 0, s, pdl, p)); \
    else                                                                                          \
      SPD_CUDA(launch(allreduce_kernel<T, W, false>, dim3(kArBlocks), dim3(kArThreads), 0, s, pdl, p)); \
  } while (0)
  SPD_DISPATCH_DTYPE(dtype, {
    switch (c->world) {
      case 2: SPD_AR_LAUNCH(2); break;
      case 4: SPD_AR_LAUNCH(4); break;
      case 8: SPD_AR_LAUNCH(8); break;
      default: SPD_AR_LAUNCH(0); break;
    }
  });
#undef SPD_AR_LAUNCH
  return 0;
}

int spd_allreduce_add_rmsnorm(spd_comm_t* c, void* out, const void* in, void* residual, const void* weight,
                              int64_t rows, int64_t dim, float eps, int dtype, spd_stream_t stream) {
  SPD_REQUIRE(c && out && in && residual && weight, SPD_E_BADARG, "spd_allreduce_add_rmsnorm: null pointer");
  SPD_REQUIRE(c->connected, SPD_E_BADARG, "spd_allreduce_add_rmsnorm: communicator is not connected");
  SPD_REQUIRE(rows >= 0 && dim > 0 && dim % 8 == 0 && dim <= 256 * kArNormMaxVec * 8, SPD_E_SHAPE,
              "spd_allreduce_add_rmsnorm: dim must be a multiple of 8 up to %d (got %lld)",
              256 * kArNormMaxVec * 8, (long long)dim);
  SPD_REQUIRE(aligned16(out) && aligned16(in) && aligned16(residual) && aligned16(weight), SPD_E_ALIGN,
              "spd_allreduce_add_rmsnorm: pointers must be 16-byte aligned");
  SPD_REQUIRE(rows * dim * 2 <= c->max_bytes, SPD_E_WORKSPACE,
              "spd_allreduce_add_rmsnorm: message of %lld bytes exceeds the communicator's %lld",
              (long long)(rows * dim * 2), (long long)c->max_bytes);
  if (rows == 0) return 0;
  ArParams p{};
  for (int r = 0; r < c->world; ++r) p.peers.base[r] = c->mapped[r];
  p.rank = c->rank;
  p.world = c->world;
  p.slot_bytes = c->max_bytes;
  p.in = in;
  p.out = out;
  p.count = rows * dim;
  p.timeout_ns = c->timeout_ns;
  ArNormParams f{};
  f.residual = residual;
  f.weight = weight;
  f.rows = rows;
  f.dim = static_cast<int>(dim);
  f.eps = eps;
  const bool pdl = pdl_enabled();
  cudaStream_t s = as_stream(stream);
#define SPD_ARN_LAUNCH(W) \
  SPD_CUDA(launch(allreduce_add_rmsnorm_kernel<T, W>, dim3(kArBlocks), dim3(kArThreads), 0, s, pdl, p, f))
  SPD_DISPATCH_DTYPE(dtype, {
    switch (c->world) {
      case 2: SPD_ARN_LAUNCH(2); break;
      case 4: SPD_ARN_LAUNCH(4); break;
      case 8: SPD_ARN_LAUNCH(8); break;
      default: SPD_ARN_LAUNCH(0); break;
    }
  });
#undef SPD_ARN_LAUNCH
  return 0;
}

int spd_allgather(spd_comm_t* c, void* out, const void* in, int64_t bytes_per_rank, spd_stream_t stream) {
  SPD_REQUIRE(c && out && in, SPD_E_BADARG, "spd_allgather: null pointer");
  SPD_REQUIRE(c->connected, SPD_E_BADARG, "spd_allgather: communicator is not connected");
  SPD_REQUIRE(bytes_per_rank >= 0 && bytes_per_rank % 16 == 0, SPD_E_ALIGN,
              "spd_allgather: bytes_per_rank must be a multiple of 16");
  SPD_REQUIRE(aligned16(out) && aligned16(in), SPD_E_ALIGN, "spd_allgather: pointers must be 16-byte aligned");
  SPD_REQUIRE(bytes_per_rank <= c->max_bytes, SPD_E_WORKSPACE,
              "spd_allgather: %lld bytes per rank exceed the communicator's %lld",
              (long long)bytes_per_rank, (long long)c->max_bytes);
  if (bytes_per_rank == 0) return 0;
  ArParams p{};
  for (int r = 0; r < c->world; ++r) p.peers.base[r] = c->mapped[r];
  p.rank = c->rank;
  p.world = c->world;
  p.slot_bytes = c->max_bytes;
  p.in = in;
  p.out = out;
  p.count = bytes_per_rank / 16;
  p.timeout_ns = c->timeout_ns;
  SPD_CUDA(launch(allgather_kernel, dim3(kArBlocks), dim3(kArThreads), 0, as_stream(stream),
                  pdl_enabled(), p));
  return 0;
}

int spd_comm_destroy(spd_comm_t* c) {
  if (!c) return 0;
  for (int r = 0; r < c->world; ++r)
    if (r != c->rank && c->mapped[r]) cudaIpcCloseMemHandle(c->mapped[r]);
  if (c->local) cudaFree(c->local);
  delete c;
  return 0;
}

}  // extern "C"
