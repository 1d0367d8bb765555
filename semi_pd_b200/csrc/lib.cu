// Library core: error text, device queries, SM partitions (CUDA green contexts), TMA descriptor
// encoding.  Driver entry points are resolved at run time through cudaGetDriverEntryPoint so the
// shared object has no link-time dependency on libcuda (it loads, and exports its symbols, on a
// box without a GPU; every compute call then fails loudly with a CUDA error).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>

#include "common.cuh"
#include "tmap.cuh"

namespace spd {

char* last_error_buf() {
  static thread_local char buf[512] = "";
  return buf;
}

int device_sm_count_cached() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SPD_PDL");  // on by default; SPD_PDL=0 turns it off
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <typename Fn>
static Fn driver_fn(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  return reinterpret_cast<Fn>(p);
}

// ------------------------------------------------------------------------------------------------
// TMA descriptors: 2-D, 16-bit elements, 128-byte swizzle, box = {64 elements, box_rows}
// ------------------------------------------------------------------------------------------------
int make_tmap_2d_sw128(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                       uint64_t row_stride_elems, uint32_t box_rows, uint32_t box_cols) {
  using Fn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                          const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                          CUtensorMapFloatOOBfill);
  static Fn fn = driver_fn<Fn>("cuTensorMapEncodeTiled");
  if (!fn) return set_error(SPD_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  // 16-bit payload moved as raw bits (no conversion), so one encoding serves bf16 and fp16
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(SPD_E_DRIVER,
                     "cuTensorMapEncodeTiled failed (%d): base=%p rows=%llu cols=%llu stride=%llu box=%ux%u",
                     (int)r, base, (unsigned long long)rows, (unsigned long long)cols,
                     (unsigned long long)row_stride_elems, box_rows, box_cols);
  return 0;
}

}  // namespace spd

using namespace spd;

// ------------------------------------------------------------------------------------------------
// SM partitions
// ------------------------------------------------------------------------------------------------
struct spd_partition {
  int device;
  CUgreenCtx ctx[2];
  CUstream stream[2];
  int sms[2];
};

extern "C" {

int spd_version(void) { return 100; }

const char* spd_last_error_string(void) { return last_error_buf(); }

int spd_device_sm_count(int device, int* sm_count) {
  SPD_REQUIRE(sm_count, SPD_E_BADARG, "spd_device_sm_count: null pointer");
  int n = 0;
  SPD_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
  *sm_count = n;
  return 0;
}

int spd_partition_create(int device, int prefill_sms, int decode_sms, spd_partition_t** out) {
  SPD_REQUIRE(out, SPD_E_BADARG, "spd_partition_create: null pointer");
  SPD_REQUIRE(prefill_sms > 0 && decode_sms > 0, SPD_E_BADARG,
              "spd_partition_create: SM counts must be positive (%d, %d)", prefill_sms, decode_sms);
  using GetRes = CUresult (*)(CUdevice, CUdevResource*, CUdevResourceType);
  using Split = CUresult (*)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*,
                             unsigned int, unsigned int);
  using GenDesc = CUresult (*)(CUdevResourceDesc*, CUdevResource*, unsigned int);
  using Create = CUresult (*)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
  using StreamCreate = CUresult (*)(CUstream*, CUgreenCtx, unsigned int, int);
  using DevGet = CUresult (*)(CUdevice*, int);
  SPD_CUDA(cudaSetDevice(device));
  SPD_CUDA(cudaFree(0));  // make sure the primary context exists
  auto get_res = driver_fn<GetRes>("cuDeviceGetDevResource");
  auto split = driver_fn<Split>("cuDevSmResourceSplitByCount");
  auto gen = driver_fn<GenDesc>("cuDevResourceGenerateDesc");
  auto create = driver_fn<Create>("cuGreenCtxCreate");
  auto screate = driver_fn<StreamCreate>("cuGreenCtxStreamCreate");
  auto devget = driver_fn<DevGet>("cuDeviceGet");
  SPD_REQUIRE(get_res && split && gen && create && screate && devget, SPD_E_DRIVER,
              "green-context driver entry points are not available");
  CUdevice dev;
  CUresult r = devget(&dev, device);
  SPD_REQUIRE(r == CUDA_SUCCESS, SPD_E_DRIVER, "cuDeviceGet failed (%d)", (int)r);
  CUdevResource all;
  r = get_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM);
  SPD_REQUIRE(r == CUDA_SUCCESS, SPD_E_DRIVER, "cuDeviceGetDevResource failed (%d)", (int)r);
  const int total = (int)all.sm.smCount;
  SPD_REQUIRE(prefill_sms + decode_sms <= total, SPD_E_BADARG,
              "spd_partition_create: %d + %d SMs exceed the device's %d", prefill_sms, decode_sms,
              total);
  // Groups come in multiples of the device's SM granularity (8 on sm_100); the group whose
  // request is such a multiple is carved first and the other one gets exactly what remains (so a
  // 76 + 72 split of 148 SMs is expressible either way round).
  const int first = (prefill_sms % 8 != 0 && decode_sms % 8 == 0) ? 1 : 0;
  const int want[2] = {prefill_sms, decode_sms};
  CUdevResource grp[2], rem0, rem1;
  unsigned int n = 1;
  r = split(&grp[first], &n, &all, &rem0, 0, (unsigned)want[first]);
  SPD_REQUIRE(r == CUDA_SUCCESS && n == 1, SPD_E_DRIVER,
              "cuDevSmResourceSplitByCount(%d) failed (%d)", want[first], (int)r);
  n = 1;
  r = split(&grp[1 - first], &n, &rem0, &rem1, 0, (unsigned)want[1 - first]);
  if (r != CUDA_SUCCESS || n != 1) {
    // the remainder is smaller than the rounded-up request: the second group is everything left
    grp[1 - first] = rem0;
  }
  spd_partition* p = new spd_partition();
  p->device = device;
  for (int i = 0; i < 2; ++i) {
    CUdevResourceDesc desc;
    r = gen(&desc, &grp[i], 1);
    if (r != CUDA_SUCCESS) {
      delete p;
      return set_error(SPD_E_DRIVER, "cuDevResourceGenerateDesc failed (%d)", (int)r);
    }
    r = create(&p->ctx[i], desc, dev, CU_GREEN_CTX_DEFAULT_STREAM);
    if (r != CUDA_SUCCESS) {
      delete p;
      return set_error(SPD_E_DRIVER, "cuGreenCtxCreate failed (%d)", (int)r);
    }
    r = screate(&p->stream[i], p->ctx[i], CU_STREAM_NON_BLOCKING, 0);
    if (r != CUDA_SUCCESS) {
      delete p;
      return set_error(SPD_E_DRIVER, "cuGreenCtxStreamCreate failed (%d)", (int)r);
    }
    p->sms[i] = (int)grp[i].sm.smCount;
  }
  *out = p;
  return 0;
}

int spd_partition_sm_counts(const spd_partition_t* p, int* prefill_sms, int* decode_sms) {
  SPD_REQUIRE(p && prefill_sms && decode_sms, SPD_E_BADARG, "spd_partition_sm_counts: null pointer");
  *prefill_sms = p->sms[0];
  *decode_sms = p->sms[1];
  return 0;
}

int spd_partition_stream(const spd_partition_t* p, int which, spd_stream_t* stream) {
  SPD_REQUIRE(p && stream && (which == 0 || which == 1), SPD_E_BADARG,
              "spd_partition_stream: bad argument");
  *stream = reinterpret_cast<spd_stream_t>(p->stream[which]);
  return 0;
}

int spd_partition_destroy(spd_partition_t* p) {
  if (!p) return 0;
  using Destroy = CUresult (*)(CUgreenCtx);
  using SDestroy = CUresult (*)(CUstream);
  auto destroy = driver_fn<Destroy>("cuGreenCtxDestroy");
  auto sdestroy = driver_fn<SDestroy>("cuStreamDestroy");
  for (int i = 0; i < 2; ++i) {
    if (sdestroy && p->stream[i]) sdestroy(p->stream[i]);
    if (destroy && p->ctx[i]) destroy(p->ctx[i]);
  }
  delete p;
  return 0;
}

}  // extern "C"
