// Point-cloud maintenance either side of the per-ray path:
//   probe_filter  -- which probed rays become new neural points: the image-space tensor code of probe_hole
//                    (run/train_studio.py:414-423: miss-ray mask, 3x3 bloat_inds dilation :447-455, far-distance rays, opacity threshold)
//                    as one kernel over the H x W probe maps;
//   vox_closest   -- construct_vox_points_closest (models/mvs/mvs_utils.py:537-561): one point per occupied voxel, the one closest to
//                    the voxel's centroid (torch.unique + torch_scatter scatter_mean / scatter_min in the reference), on a dense
//                    counting grid like grid.cu's: accumulate -> arg-min -> scan -> emit, voxels in (x, y, z) lexicographic order.
#include "pnerf_common.cuh"

namespace pnerf {
namespace {

__device__ __forceinline__ float norm3(float a, float b, float c) { return sqrtf(a * a + b * b + c * c); }

__global__ void __launch_bounds__(256) probe_filter_kernel(const int8_t* __restrict__ ray_mask, const float* __restrict__ gt,
                                                           const float* __restrict__ color, const float* __restrict__ far_dist,
                                                           const float* __restrict__ opacity, const uint8_t* __restrict__ edge, float bg0,
                                                           float bg1, float bg2, int H, int W, float far_thresh, float opacity_thresh,
                                                           uint8_t* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= H * W) return;
    const int y = p / W, x = p % W;
    bool keep = false;
    if (ray_mask[p] > 0 && opacity[p] > opacity_thresh) {
        // neighbouring_miss_mask: some pixel q of the 3x3 window (clamped at the border like bloat_inds) is a miss ray --
        // inside the edge mask, no neural point along it, ground truth not background (TS:414-419)
        for (int dy = -1; dy <= 1 && !keep; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                const int qy = y + dy, qx = x + dx;
                if (qy < 0 || qy >= H || qx < 0 || qx >= W) continue;
                const int q = qy * W + qx;
                if ((edge == nullptr || edge[q]) && ray_mask[q] < 1 &&
                    norm3(gt[3 * q] - bg0, gt[3 * q + 1] - bg1, gt[3 * q + 2] - bg2) > 0.002f) { keep = true; break; }
            }
        // far rays: a neural point was found but the densest sample is far from its nearest neighbour and the colour is right (TS:420-422)
        if (!keep && far_thresh > 0.f && far_dist[p] > far_thresh &&
            norm3(gt[3 * p] - color[3 * p], gt[3 * p + 1] - color[3 * p + 1], gt[3 * p + 2] - color[3 * p + 2]) < 0.1f)
            keep = true;
    }
    out[p] = keep ? 1 : 0;
}

struct VoxFrame { float mn[3]; float sz[3]; int dim[3]; };

__device__ __forceinline__ int64_t vox_cell(const VoxFrame& f, const float* __restrict__ p, bool& inside) {
    // floor((xyz - space_min) / construct_vox_sz) in fp32, exactly as MU:552-553 evaluates it
    const float fx = floorf(__fdiv_rn(__fsub_rn(p[0], f.mn[0]), f.sz[0]));
    const float fy = floorf(__fdiv_rn(__fsub_rn(p[1], f.mn[1]), f.sz[1]));
    const float fz = floorf(__fdiv_rn(__fsub_rn(p[2], f.mn[2]), f.sz[2]));
    inside = fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)f.dim[0] && fy < (float)f.dim[1] && fz < (float)f.dim[2];
    return inside ? ((int64_t)fx * f.dim[1] + (int64_t)fy) * f.dim[2] + (int64_t)fz : -1;
}

__global__ void __launch_bounds__(256) vox_accumulate_kernel(VoxFrame f, const float* __restrict__ xyz, int64_t n, int* __restrict__ cnt,
                                                             double* __restrict__ sum, int* __restrict__ n_outside) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        bool in;
        const int64_t c = vox_cell(f, xyz + 3 * i, in);
        if (!in) { atomicAdd(n_outside, 1); continue; }
        atomicAdd(cnt + c, 1);
        atomicAdd(sum + 3 * c, (double)xyz[3 * i]);
        atomicAdd(sum + 3 * c + 1, (double)xyz[3 * i + 1]);
        atomicAdd(sum + 3 * c + 2, (double)xyz[3 * i + 2]);
    }
}

__global__ void __launch_bounds__(256) vox_argmin_kernel(VoxFrame f, const float* __restrict__ xyz, int64_t n, const int* __restrict__ cnt,
                                                         const double* __restrict__ sum, unsigned long long* __restrict__ best) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        bool in;
        const int64_t c = vox_cell(f, xyz + 3 * i, in);
        if (!in) continue;
        const double inv = 1.0 / (double)cnt[c];
        const float cx = (float)(sum[3 * c] * inv), cy = (float)(sum[3 * c + 1] * inv), cz = (float)(sum[3 * c + 2] * inv);
        const float r = norm3(xyz[3 * i] - cx, xyz[3 * i + 1] - cy, xyz[3 * i + 2] - cz);       // MU:555-556
        // smallest residual, lowest point index on ties (scatter_min's first-minimum rule): one 64-bit atomicMin
        atomicMin(best + c, ((unsigned long long)__float_as_uint(r) << 32) | (unsigned long long)(uint32_t)i);
    }
}

__global__ void __launch_bounds__(256) vox_flags_kernel(const int* __restrict__ cnt, int64_t cells, int* __restrict__ flags) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < cells) flags[c] = cnt[c] > 0 ? 1 : 0;
}

__global__ void __launch_bounds__(256) vox_emit_kernel(VoxFrame f, const int* __restrict__ cnt, const double* __restrict__ sum,
                                                       const unsigned long long* __restrict__ best, const int* __restrict__ pos, int64_t cells,
                                                       int max_out, float* __restrict__ centroid, int* __restrict__ grid_idx,
                                                       int* __restrict__ min_idx, int* __restrict__ n_out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0) *n_out = pos[cells];
    if (c >= cells || cnt[c] == 0) return;
    const int o = pos[c];
    if (o >= max_out) return;
    const double inv = 1.0 / (double)cnt[c];
    centroid[3 * o] = (float)(sum[3 * c] * inv); centroid[3 * o + 1] = (float)(sum[3 * c + 1] * inv); centroid[3 * o + 2] = (float)(sum[3 * c + 2] * inv);
    const int z = (int)(c % f.dim[2]), y = (int)((c / f.dim[2]) % f.dim[1]), x = (int)(c / ((int64_t)f.dim[1] * f.dim[2]));
    grid_idx[3 * o] = x; grid_idx[3 * o + 1] = y; grid_idx[3 * o + 2] = z;
    min_idx[o] = (int)(best[c] & 0xffffffffull);
}

int64_t vox_cells(const int* dim) { return (int64_t)dim[0] * dim[1] * dim[2]; }

}  // namespace
}  // namespace pnerf

using namespace pnerf;

extern "C" int pnerf_probe_filter(const int8_t* ray_mask, const float* gt, const float* color, const float* far_dist, const float* opacity,
                                  const uint8_t* edge_mask, const float* bg_h, int H, int W, float far_thresh, float opacity_thresh,
                                  uint8_t* keep, void* stream) {
    if (H < 0 || W < 0 || !bg_h) return PNERF_ERR_ARG;
    if ((int64_t)H * W == 0) return PNERF_OK;
    if ((int64_t)H * W > 0x7fffffff || !ray_mask || !gt || !opacity || !keep || (far_thresh > 0.f && (!far_dist || !color))) return PNERF_ERR_ARG;
    probe_filter_kernel<<<(unsigned)(((int64_t)H * W + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ray_mask, gt, color, far_dist, opacity, edge_mask,
                                                                                                   bg_h[0], bg_h[1], bg_h[2], H, W, far_thresh,
                                                                                                   opacity_thresh, keep);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}

extern "C" int64_t pnerf_vox_closest_workspace_bytes(const int* dim_h) {
    if (!dim_h || dim_h[0] <= 0 || dim_h[1] <= 0 || dim_h[2] <= 0) return 0;
    const int64_t g = vox_cells(dim_h);
    return align_up(g * 4, 256) + align_up(g * 24, 256) + align_up(g * 8, 256) + align_up((g + 1) * 4, 256) + scan_workspace_bytes(g + 1) + 512;
}

extern "C" int pnerf_vox_closest(const float* xyz, int64_t n, const float* space_min_h, const float* vox_size_h, const int* dim_h, int max_out,
                                 float* centroid, int* grid_idx, int* min_idx, int* n_out, int* n_outside, void* workspace,
                                 int64_t workspace_bytes, void* stream) {
    if (n < 0 || !space_min_h || !vox_size_h || !dim_h || max_out < 0 || !n_out || !n_outside) return PNERF_ERR_ARG;
    if (dim_h[0] <= 0 || dim_h[1] <= 0 || dim_h[2] <= 0 || vox_cells(dim_h) >= 0x7fffffff) return PNERF_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    PNERF_CUDA(cudaMemsetAsync(n_out, 0, 4, st));
    PNERF_CUDA(cudaMemsetAsync(n_outside, 0, 4, st));
    if (n == 0) return PNERF_OK;
    if (!xyz || !centroid || !grid_idx || !min_idx || !workspace) return PNERF_ERR_ARG;
    if (workspace_bytes < pnerf_vox_closest_workspace_bytes(dim_h)) return PNERF_ERR_WORKSPACE;
    const int64_t g = vox_cells(dim_h);
    VoxFrame f;
    for (int a = 0; a < 3; a++) { f.mn[a] = space_min_h[a]; f.sz[a] = vox_size_h[a]; f.dim[a] = dim_h[a]; }
    uint8_t* w = (uint8_t*)workspace;
    int* cnt = (int*)w; w += align_up(g * 4, 256);
    double* sum = (double*)w; w += align_up(g * 24, 256);
    unsigned long long* best = (unsigned long long*)w; w += align_up(g * 8, 256);
    int* pos = (int*)w; w += align_up((g + 1) * 4, 256);
    const int64_t scan_bytes = workspace_bytes - (w - (uint8_t*)workspace);
    PNERF_CUDA(cudaMemsetAsync(cnt, 0, (size_t)g * 4, st));
    PNERF_CUDA(cudaMemsetAsync(sum, 0, (size_t)g * 24, st));
    PNERF_CUDA(cudaMemsetAsync(best, 0xff, (size_t)g * 8, st));
    const int64_t pb = (n + 255) / 256;
    const unsigned pblocks = (unsigned)(pb > (int64_t)kSMs * 32 ? (int64_t)kSMs * 32 : pb);
    vox_accumulate_kernel<<<pblocks, 256, 0, st>>>(f, xyz, n, cnt, sum, n_outside);
    PNERF_LAUNCH_CHECK();
    vox_argmin_kernel<<<pblocks, 256, 0, st>>>(f, xyz, n, cnt, sum, best);
    PNERF_LAUNCH_CHECK();
    const unsigned cblocks = (unsigned)((g + 255) / 256);
    vox_flags_kernel<<<cblocks, 256, 0, st>>>(cnt, g, pos);
    PNERF_LAUNCH_CHECK();
    int rc = exclusive_scan_i32(pos, pos, g, true, w, scan_bytes, st);
    if (rc) return rc;
    vox_emit_kernel<<<cblocks, 256, 0, st>>>(f, cnt, sum, best, pos, g, max_out, centroid, grid_idx, min_idx, n_out);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}
