// placeholder until the tcgen05 prefill kernel lands (next commit)
#include "common.cuh"
using namespace spd;
extern "C" int spd_attn_prefill_ragged(void*, const void*, const void*, const void*, const int32_t*,
                                       const int64_t*, const int32_t*, const int32_t*, const int32_t*,
                                       int64_t, int64_t, int64_t, int, int, int, int64_t, int64_t,
                                       int64_t, float, int, spd_stream_t) {
  return set_error(SPD_E_SHAPE, "spd_attn_prefill_ragged: not built yet");
}
