// placeholder until the tcgen05 decode kernel lands (next commit)
#include "common.cuh"
using namespace spd;
extern "C" int spd_attn_decode_paged(void*, const void*, const void*, const void*, const int32_t*,
                                     const int64_t*, const int64_t*, int64_t, int, int, int, int64_t,
                                     int64_t, int64_t, float, int, void*, int64_t, int, spd_stream_t) {
  return set_error(SPD_E_SHAPE, "spd_attn_decode_paged: not built yet");
}
