// Host-side TMA descriptor encoding (implemented in lib.cu).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace spd {

// 2-D row-major tensor of 16-bit elements [rows, cols] with `row_stride_elems` between rows;
// box = box_rows x box_cols (box_cols * 2 bytes must be 128 for the 128-byte swizzle).
// Out-of-bounds parts of a box are zero-filled by the hardware.
int make_tmap_2d_sw128(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                       uint64_t row_stride_elems, uint32_t box_rows, uint32_t box_cols);

// Cached variant (implemented in gemm.cu): descriptors are pure functions of (pointer, shape,
// box), so steady-state launches do no driver work.  box_cols is fixed at 64 (one swizzle atom).
int get_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t stride,
             uint32_t box_rows);

}  // namespace spd
