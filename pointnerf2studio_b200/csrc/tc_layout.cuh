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
}  // namespace tcl

int field_tc_launch(const pnerf_points* pts, const pnerf_camera* cam, const pnerf_mlp* mlp, const void* wpack, const pnerf_mode* mode,
                    const float* dirs, const float* sample_loc, const int* sample_pidx, const int* sample_ids, int S, int SR, int K,
                    float* sigma, float* rgb, void* F, uint8_t* save, float* save_w, float* save_raw, bool color, cudaStream_t st);

// fp32 colour network (mlp_color + rgb head) on an (S,256) bf16 feature matrix: used by the tensor-core training path, implemented
// with the SIMT kernels of field_f32.cu.  `ws` holds color_f32_ws_floats(S) floats and carries the activations to the backward.
int64_t color_f32_ws_floats(int64_t S);
int color_forward_f32(const pnerf_points* pts, const pnerf_camera* cam, const pnerf_mlp* mlp, const pnerf_mode* mode, const float* dirs,
                      const int* sample_ids, int S, int SR, const void* F_bf16, float* ws, float* rgb, cudaStream_t st);
// -> *dF = (S, ldF) fp32 gradient of the aggregated features (inside ws)
int color_backward_f32(const pnerf_mlp* mlp, const pnerf_mlp_grad* gm, const pnerf_mode* mode, const int* sample_ids, int S,
                       const float* d_rgb, float* ws, const float** dF, int* ldF, cudaStream_t st);
}  // namespace pnerf
