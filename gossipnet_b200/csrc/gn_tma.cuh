// Host side of tensor-map TMA: build a CUtensorMap for a 2D bf16 matrix.  The encoder is a
// driver function; it is resolved through the runtime (cudaGetDriverEntryPoint), so the
// library keeps linking against cudart only.  A tensor map is 128 bytes of plain data: the
// launchers build it on their stack per call and pass it by value (__grid_constant__), so
// there is no global state and CUDA graphs capture it with the other kernel parameters.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gn {

typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// nullptr when the driver does not export cuTensorMapEncodeTiled
tmap_encode_fn tmap_encoder();

// rows x cols bf16 matrix, row pitch in bytes (multiple of 16), box = box_rows x box_cols with
// box_cols * 2 <= 128 bytes, SWIZZLE_128B, out-of-bounds elements read as zero.
// Returns 0 on success, the CUresult otherwise (-1: no encoder).
int encode_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                        uint64_t pitch_bytes, uint32_t box_rows, uint32_t box_cols);
// the same for a float32 matrix (box_cols * 4 <= 128 bytes)
int encode_tmap_2d_f32(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                       uint64_t pitch_bytes, uint32_t box_rows, uint32_t box_cols);

}  // namespace gn
