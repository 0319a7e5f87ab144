// Shared device/host helpers for libpnerf_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pnerf_b200.h"

namespace pnerf {

extern thread_local char g_last_error[256];
int set_cuda_error(cudaError_t e, const char* where);

#define PNERF_CUDA(call)                                                      \
    do {                                                                      \
        cudaError_t _e = (call);                                              \
        if (_e != cudaSuccess) return ::pnerf::set_cuda_error(_e, #call);     \
    } while (0)

#define PNERF_LAUNCH_CHECK()                                                  \
    do {                                                                      \
        cudaError_t _e = cudaGetLastError();                                  \
        if (_e != cudaSuccess) return ::pnerf::set_cuda_error(_e, __func__);  \
    } while (0)

constexpr int kSMs = 148;

struct Frame {
    float lo[3];
    float sv[3];
    int dim[3];
};

static inline Frame frame_of(const pnerf_grid_view* g) {
    Frame f;
    for (int a = 0; a < 3; a++) { f.lo[a] = g->lo[a]; f.sv[a] = g->sv[a]; f.dim[a] = g->dim[a]; }
    return f;
}

// (int) floor((p - lo) / sv), IEEE fp32 sub + div exactly like query_worldcoords.cu:38-40; returns
// false when outside the grid (CU:44).  NaN compares false -> outside.
__device__ __forceinline__ bool voxel_of(const Frame& f, float x, float y, float z, int& vx, int& vy, int& vz) {
    float fx = floorf(__fdiv_rn(__fsub_rn(x, f.lo[0]), f.sv[0]));
    float fy = floorf(__fdiv_rn(__fsub_rn(y, f.lo[1]), f.sv[1]));
    float fz = floorf(__fdiv_rn(__fsub_rn(z, f.lo[2]), f.sv[2]));
    bool in = (fx >= 0.f) && (fx < (float)f.dim[0]) && (fy >= 0.f) && (fy < (float)f.dim[1]) && (fz >= 0.f) &&
              (fz < (float)f.dim[2]);
    vx = in ? (int)fx : 0;
    vy = in ? (int)fy : 0;
    vz = in ? (int)fz : 0;
    return in;
}

__device__ __forceinline__ int cell_lin(const Frame& f, int x, int y, int z) {
    return (x * f.dim[1] + y) * f.dim[2] + z;
}

// exclusive scan of int32 (n up to 2^31-1); out may alias in; out has n+1 entries when `with_total`.
int64_t scan_workspace_bytes(int64_t n);
int exclusive_scan_i32(const int* in, int* out, int64_t n, bool with_total, void* ws, int64_t ws_bytes,
                       cudaStream_t st);

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

}  // namespace pnerf
