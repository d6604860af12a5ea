// Cached creation of TMA tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point; no -lcuda).
#pragma once
#include "common.cuh"

// bf16 tensor [d2][d1][d0] (d0 contiguous), strides in ELEMENTS (ld1 between d1 rows, ld2 between d2 slabs; d2 = 1 for
// a matrix).  Box = box0 x box1 x 1 elements, 128-byte swizzle (box0 * 2 bytes must be <= 128), zero fill out of bounds.
int make_tmap_bf16(const bf16* ptr, int d0, int d1, int d2, int64_t ld1, int64_t ld2, int box0, int box1, CUtensorMap* out);
// 4-D: [d3][d2][d1][d0], box0 x 1 x box2 x 1 (attention: head dim, head, position, batch).
int make_tmap_bf16_4d(const bf16* ptr, int d0, int d1, int d2, int d3, int64_t ld1, int64_t ld2, int64_t ld3, int box0, int box2,
                      CUtensorMap* out);
