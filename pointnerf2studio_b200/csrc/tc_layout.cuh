// Tile layout shared by the tensor-core forward (field_tc.cu) and backward (field_tc_train.cu) kernels.
// A tile = 128 rows; an operand tile is stored as 8-column "k-slabs" of 128 rows x 16 bytes (see umma.cuh), in shared memory and --
// for training -- verbatim in global memory, so a backward kernel bulk-copies a saved tile straight into an MMA operand buffer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pnerf_b200.h"

namespace pnerf {
namespace tcl {
constexpr int ROWS = 128;
constexpr int SLAB = ROWS * 16;                  // bytes of one 8-wide k-slab of a 128-row operand
// forward operands kept per tile for the backward pass: [X0 36 slabs | H1 32 | X3 36 | H3 32 | H4 32]
constexpr int SAVE_X0 = 0, SAVE_H1 = 36, SAVE_X3 = 68, SAVE_H3 = 104, SAVE_H4 = 136, SAVE_SLABS = 168;
constexpr int64_t SAVE_TILE_BYTES = (int64_t)SAVE_SLABS * SLAB;     // 344 064
constexpr int64_t DELTA_TILE_BYTES = (int64_t)32 * SLAB;            // a 128 x 256 bf16 gradient tile
// colour network operands kept per 128-sample tile: [C0 = [F_s | PE(v) | 0] 36 slabs | C1 16 | C2 16 | C3 16]
constexpr int CSAVE_C0 = 0, CSAVE_C1 = 36, CSAVE_C2 = 52, CSAVE_C3 = 68, CSAVE_SLABS = 84;
constexpr int64_t CSAVE_TILE_BYTES = (int64_t)CSAVE_SLABS * SLAB;   // 172 032
constexpr int64_t CDELTA_TILE_BYTES = (int64_t)16 * SLAB;           // a 128 x 128 bf16 gradient tile
// aggregated features F between the two forward kernels: per 128-sample tile 32 k-slabs = the first 32 slabs of the colour network's
// A operand, so the colour kernel fetches a tile with bulk copies instead of a per-thread gather
constexpr int64_t F_TILE_BYTES = (int64_t)32 * SLAB;                // 65 536
}  // namespace tcl

int field_tc_launch(const pnerf_points* pts, const pnerf_camera* cam, const pnerf_mlp* mlp, const void* wpack, const pnerf_mode* mode,
                    const float* dirs, const float* sample_loc, const int* sample_pidx, const int* sample_ids, int S, const int* S_dev,
                    int SR, int K, float* sigma, float* rgb, void* F, uint8_t* save, float* save_w, float* save_raw, bool color,
                    uint8_t* csave, cudaStream_t st, int kp_override = 0, int si0 = 0, bool field = true);

}  // namespace pnerf
