// Grid frame reduction (row H) and voxel-bucket build (row G1) -- replaces claim_occ / map_coor2occ /
// fill_occ2pnts of the reference (query_worldcoords.cu:18-162).
//
// B200-first layout: instead of three racy hash-like tables (coor_2_occ, occ_2_coor, occ_2_pnts) that
// the reference re-allocates and re-fills on every forward, the cloud is bucket-sorted ONCE per
// point-cloud version into a CSR over the dense cell grid:
//   cell_start[G+1]  int32 offsets          (4 B per cell, L2 resident: 22 MB for a 1.4-unit object)
//   recs[n] float4   (x, y, z, index|vz<<28) sorted by (cell, index)
// Cell ids run z-fastest, so the three z-neighbours of a voxel row are ONE contiguous run of 16-byte
// records: a 3x3x3 neighbourhood is 9 coalesced runs instead of 27 x (table + count + P indices + P
// scattered xyz loads).  Occupancy is a G-bit mask (3.5 MB for the full +-1.2 box, L1/L2 resident).
// HBM-bound integer work: grids are sized in multiples of the SM count with grid-stride loops.
#include "pnerf_common.cuh"

namespace pnerf {
namespace {

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// float atomic min/max through the ordered-int trick
__device__ __forceinline__ void atomic_minf(float* a, float v) {
    if (v >= 0.f) atomicMin((int*)a, __float_as_int(v)); else atomicMax((unsigned*)a, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_maxf(float* a, float v) {
    if (v >= 0.f) atomicMax((int*)a, __float_as_int(v)); else atomicMin((unsigned*)a, __float_as_uint(v));
}

__global__ void bbox_init_kernel(float* out) {
    if (threadIdx.x < 3) out[threadIdx.x] = INFINITY;
    else if (threadIdx.x < 6) out[threadIdx.x] = -INFINITY;
}

__global__ void __launch_bounds__(256) bbox_kernel(const float* __restrict__ xyz, int64_t n, float* __restrict__ out) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            float v = xyz[3 * i + a];
            mn[a] = fminf(mn[a], v);
            mx[a] = fmaxf(mx[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
        float lo = warp_min(mn[a]), hi = warp_max(mx[a]);
        if ((threadIdx.x & 31) == 0) {
            if (lo != INFINITY) atomic_minf(out + a, lo);
            if (hi != -INFINITY) atomic_maxf(out + 3 + a, hi);
        }
    }
}

// pass 1: cell id of each point (-1 outside the clipped grid, CU:44) + per-cell counts
__global__ void __launch_bounds__(256) count_kernel(const float* __restrict__ xyz, int64_t n, Frame f,
                                                     int* __restrict__ pt_cell, int* __restrict__ cnt) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int vx, vy, vz;
        bool in = voxel_of(f, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], vx, vy, vz);
        int c = in ? cell_lin(f, vx, vy, vz) : -1;
        pt_cell[i] = c;
        if (in) atomicAdd(cnt + c, 1);
    }
}

// pass 2: drop every point into its cell's segment of `tmp` (slot order is race order; fixed by pass 3)
__global__ void __launch_bounds__(256) fill_kernel(const int* __restrict__ pt_cell, int64_t n,
                                                    const int* __restrict__ start_full, int* __restrict__ cursor,
                                                    int* __restrict__ tmp) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int c = pt_cell[i];
        if (c >= 0) tmp[start_full[c] + atomicAdd(cursor + c, 1)] = (int)i;
    }
}

// pass 3a: capped counts min(n_c, P) (the reference keeps at most P per voxel, CU:149-151)
__global__ void __launch_bounds__(256) cap_kernel(const int* __restrict__ start_full, int64_t G, int P,
                                                   int* __restrict__ capped) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < G; c += (int64_t)gridDim.x * blockDim.x)
        capped[c] = min(start_full[c + 1] - start_full[c], P);
}

// pass 3b: per occupied cell, order its segment by point index, keep the first P, emit the records,
// and set the dilated occupancy bits (CU:105-112).
__global__ void __launch_bounds__(256) finalize_kernel(const float* __restrict__ xyz, const int* __restrict__ start_full,
                                                        const int* __restrict__ cell_start, int* __restrict__ tmp,
                                                        int64_t G, int P, Frame f, int q0, int q1, int q2,
                                                        float4* __restrict__ recs, uint32_t* __restrict__ occ_bits) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < G; c += (int64_t)gridDim.x * blockDim.x) {
        const int a = start_full[c], n = start_full[c + 1] - a;
        if (n == 0) continue;
        int* seg = tmp + a;
        const int keep = min(n, P);
        if (n <= 64) {  // insertion sort
            for (int i = 1; i < n; i++) {
                int v = seg[i], j = i - 1;
                while (j >= 0 && seg[j] > v) { seg[j + 1] = seg[j]; j--; }
                seg[j + 1] = v;
            }
        } else {        // selection of the `keep` smallest
            for (int i = 0; i < keep; i++) {
                int best = i;
                for (int j = i + 1; j < n; j++) if (seg[j] < seg[best]) best = j;
                int t = seg[i]; seg[i] = seg[best]; seg[best] = t;
            }
        }
        const int vz = (int)(c % f.dim[2]);
        const int vy = (int)((c / f.dim[2]) % f.dim[1]);
        const int vx = (int)(c / ((int64_t)f.dim[2] * f.dim[1]));
        const int o = cell_start[c];
        for (int i = 0; i < keep; i++) {
            int p = seg[i];
            recs[o + i] = make_float4(xyz[3 * (int64_t)p], xyz[3 * (int64_t)p + 1], xyz[3 * (int64_t)p + 2],
                                      __int_as_float(p | ((vz & 7) << 28)));
        }
        for (int x = max(0, vx - q0 / 2); x < min(f.dim[0], vx + (q0 + 1) / 2); x++)
            for (int y = max(0, vy - q1 / 2); y < min(f.dim[1], vy + (q1 + 1) / 2); y++)
                for (int z = max(0, vz - q2 / 2); z < min(f.dim[2], vz + (q2 + 1) / 2); z++) {
                    int id = cell_lin(f, x, y, z);
                    atomicOr(occ_bits + (id >> 5), 1u << (id & 31));
                }
    }
}

int grid_for(int64_t n, int threads) {
    int64_t b = (n + threads - 1) / threads;
    int64_t cap = (int64_t)kSMs * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
}  // namespace
}  // namespace pnerf

using namespace pnerf;

extern "C" int pnerf_bbox(const float* xyz, int64_t n, float* out_minmax, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!out_minmax || n < 0 || (n > 0 && !xyz)) return PNERF_ERR_ARG;
    bbox_init_kernel<<<1, 32, 0, st>>>(out_minmax);
    PNERF_LAUNCH_CHECK();
    if (n > 0) {
        bbox_kernel<<<grid_for(n, 256), 256, 0, st>>>(xyz, n, out_minmax);
        PNERF_LAUNCH_CHECK();
    }
    return PNERF_OK;
}

extern "C" int64_t pnerf_grid_workspace_bytes(int64_t n, int64_t cells) {
    // pt_cell[n] + tmp[n] + cnt[G] + start_full[G+1] + scan workspace
    return align_up(n * 4, 256) * 2 + align_up(cells * 4, 256) + align_up((cells + 1) * 4, 256) +
           scan_workspace_bytes(cells + 1) + 1024;
}

extern "C" int pnerf_grid_build(const float* xyz, int64_t n, const float* lo_h, const float* sv_h, const int* dim_h,
                                int P, const int* query_size_h, int* cell_start, float* recs, uint32_t* occ_bits,
                                void* workspace, int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!lo_h || !sv_h || !dim_h || !query_size_h || !cell_start || !occ_bits || !workspace || n < 0 || P <= 0)
        return PNERF_ERR_ARG;
    if (n > 0 && (!xyz || !recs)) return PNERF_ERR_ARG;
    if (n >= (1 << 28)) return PNERF_ERR_ARG;  // index shares a word with 3 bits of vz
    Frame f;
    for (int a = 0; a < 3; a++) {
        f.lo[a] = lo_h[a]; f.sv[a] = sv_h[a]; f.dim[a] = dim_h[a];
        if (dim_h[a] <= 0 || !(sv_h[a] > 0.f)) return PNERF_ERR_ARG;
    }
    const int64_t G = (int64_t)f.dim[0] * f.dim[1] * f.dim[2];
    if (G >= (int64_t)1 << 31) return PNERF_ERR_ARG;
    if (workspace_bytes < pnerf_grid_workspace_bytes(n, G)) return PNERF_ERR_WORKSPACE;
    char* w = (char*)workspace;
    int* pt_cell = (int*)w; w += align_up(n * 4, 256);
    int* tmp = (int*)w; w += align_up(n * 4, 256);
    int* cnt = (int*)w; w += align_up(G * 4, 256);
    int* start_full = (int*)w; w += align_up((G + 1) * 4, 256);
    void* scan_ws = w;
    const int64_t scan_bytes = workspace_bytes - (w - (char*)workspace);

    PNERF_CUDA(cudaMemsetAsync(cnt, 0, G * 4, st));
    PNERF_CUDA(cudaMemsetAsync(occ_bits, 0, ((G + 31) / 32) * 4, st));
    if (n > 0) {
        count_kernel<<<grid_for(n, 256), 256, 0, st>>>(xyz, n, f, pt_cell, cnt);
        PNERF_LAUNCH_CHECK();
    }
    int rc = exclusive_scan_i32(cnt, start_full, G, true, scan_ws, scan_bytes, st);
    if (rc) return rc;
    PNERF_CUDA(cudaMemsetAsync(cnt, 0, G * 4, st));
    if (n > 0) {
        fill_kernel<<<grid_for(n, 256), 256, 0, st>>>(pt_cell, n, start_full, cnt, tmp);
        PNERF_LAUNCH_CHECK();
    }
    cap_kernel<<<grid_for(G, 256), 256, 0, st>>>(start_full, G, P, cnt);
    PNERF_LAUNCH_CHECK();
    rc = exclusive_scan_i32(cnt, cell_start, G, true, scan_ws, scan_bytes, st);
    if (rc) return rc;
    finalize_kernel<<<grid_for(G, 256), 256, 0, st>>>(xyz, start_full, cell_start, tmp, G, P, f, query_size_h[0],
                                                     query_size_h[1], query_size_h[2], (float4*)recs, occ_bits);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}
