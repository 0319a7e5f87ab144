// Exclusive int32 scan (multi-level reduce / scan / add) and the stream compactions built on it.
// Used by the grid build (cell offsets) and by the ray / sample compaction of the per-call path.
#include "pnerf_common.cuh"

namespace pnerf {

thread_local char g_last_error[256] = {0};

int set_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s", where, cudaGetErrorString(e));
    cudaGetLastError();
    return PNERF_ERR_CUDA;
}

namespace {
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
    __shared__ int warp_sums[kScanThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < kScanThreads / 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        if (lane < kScanThreads / 32) warp_sums[lane] = w;  // inclusive
    }
    __syncthreads();
    int base = warp > 0 ? warp_sums[warp - 1] : 0;
    *total = warp_sums[kScanThreads / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

__global__ void __launch_bounds__(kScanThreads) tile_sums_kernel(const int* __restrict__ in, int64_t n,
                                                                  int* __restrict__ sums) {
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        int64_t j = base + i;
        if (j < n) s += in[j];
    }
    int total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// out[i] = offsets[tile] + exclusive prefix inside the tile.  `offsets` may be null (single tile).
__global__ void __launch_bounds__(kScanThreads) tile_scan_kernel(const int* in, int* out, int64_t n,
                                                                  const int* __restrict__ offsets, int write_total) {
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int v[kScanItems];
    int s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        int64_t j = base + i;
        v[i] = j < n ? in[j] : 0;
        s += v[i];
    }
    int total;
    int pre = block_exclusive_scan(s, &total) + (offsets ? offsets[blockIdx.x] : 0);
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        int64_t j = base + i;
        if (j < n) out[j] = pre;
        pre += v[i];
    }
    if (write_total && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0)
        out[n] = (offsets ? offsets[blockIdx.x] : 0) + total;
}
}  // namespace

int64_t scan_workspace_bytes(int64_t n) {
    int64_t bytes = 0;
    while (n > kScanTile) {
        n = (n + kScanTile - 1) / kScanTile;
        bytes += align_up(n * 4, 256);
    }
    return bytes + 256;
}

int exclusive_scan_i32(const int* in, int* out, int64_t n, bool with_total, void* ws, int64_t ws_bytes,
                       cudaStream_t st) {
    if (n < 0) return PNERF_ERR_ARG;
    if (n == 0) {
        if (with_total) PNERF_CUDA(cudaMemsetAsync(out, 0, 4, st));
        return PNERF_OK;
    }
    if (ws_bytes < scan_workspace_bytes(n)) return PNERF_ERR_WORKSPACE;
    const int64_t tiles = (n + kScanTile - 1) / kScanTile;
    if (tiles == 1) {
        tile_scan_kernel<<<1, kScanThreads, 0, st>>>(in, out, n, nullptr, with_total ? 1 : 0);
        PNERF_LAUNCH_CHECK();
        return PNERF_OK;
    }
    int* sums = (int*)ws;
    tile_sums_kernel<<<(unsigned)tiles, kScanThreads, 0, st>>>(in, n, sums);
    PNERF_LAUNCH_CHECK();
    char* next_ws = (char*)ws + align_up(tiles * 4, 256);
    int rc = exclusive_scan_i32(sums, sums, tiles, false, next_ws, ws_bytes - align_up(tiles * 4, 256), st);
    if (rc) return rc;
    tile_scan_kernel<<<(unsigned)tiles, kScanThreads, 0, st>>>(in, out, n, sums, with_total ? 1 : 0);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}

// ------------------------------------------------------------------------------------------------
namespace {
__global__ void flags_to_i32_kernel(const uint8_t* __restrict__ flags, int64_t n, int* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = flags[i] ? 1 : 0;
}
__global__ void scatter_ids_kernel(const uint8_t* __restrict__ flags, const int* __restrict__ pos, int64_t n,
                                   int* __restrict__ ids, int* __restrict__ count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flags[i]) ids[pos[i]] = (int)i;
    if (i == 0) *count = pos[n];
}
// Samples bucketed by neighbour count: class c holds the samples whose count n satisfies lo_c < n <= kp_c (kp = 2, 4, 8, ...): the
// tensor-core field kernel then runs class c with kp_c rows per sample instead of K rows for everybody (16 % of the rows of the
// render bench are padding at 8 rows per sample: 19 % of its samples have <= 4 neighbours).
constexpr int kMaxClasses = 8;
struct ClassBounds { int lo[kMaxClasses], hi[kMaxClasses]; int n; };
__device__ __forceinline__ int class_of(const ClassBounds& b, int c) {
    int cls = -1;
#pragma unroll
    for (int k = 0; k < kMaxClasses; k++)
        if (k < b.n && c > b.lo[k] && c <= b.hi[k]) cls = k;
    return cls;
}
// Three phases over tiles of 256 slots: per-tile class counts -> exclusive scan of each class's tile counts (n_slots / 256 entries,
// not n_slots) -> scatter with the tile bases.  Deterministic, ascending inside every class (neighbouring samples of a ray, which
// share neural points, stay neighbours: an arrival-ordered variant with atomics cost the field kernels 2 % of their time).
__global__ void __launch_bounds__(256) class_tile_count_kernel(const uint8_t* __restrict__ cnt, int64_t n, ClassBounds b, int64_t n_tiles,
                                                               int* __restrict__ tile_counts /* [class][n_tiles + 1] */) {
    __shared__ int s_tot[kMaxClasses];
    const int lane = threadIdx.x & 31;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        if (threadIdx.x < kMaxClasses) s_tot[threadIdx.x] = 0;
        __syncthreads();
        const int64_t i = t * 256 + threadIdx.x;
        const int cls = i < n ? class_of(b, cnt[i]) : -1;
        for (int k = 0; k < b.n; k++) {
            const unsigned m = __ballot_sync(0xffffffffu, cls == k);
            if (lane == 0 && m) atomicAdd(&s_tot[k], __popc(m));
        }
        __syncthreads();
        if (threadIdx.x < b.n) tile_counts[(int64_t)threadIdx.x * (n_tiles + 1) + t] = s_tot[threadIdx.x];
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) class_tile_scatter_kernel(const uint8_t* __restrict__ cnt, int64_t n, ClassBounds b, int64_t n_tiles,
                                                                 const int* __restrict__ tile_pos /* scanned, [class][n_tiles + 1] */,
                                                                 int* __restrict__ ids, int* __restrict__ counts) {
    __shared__ int s_w[8][kMaxClasses];      // per warp and class: count, then exclusive offset inside the tile
    __shared__ int s_base[kMaxClasses];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (blockIdx.x == 0 && threadIdx.x < b.n) counts[threadIdx.x] = tile_pos[(int64_t)threadIdx.x * (n_tiles + 1) + n_tiles];
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t i = t * 256 + threadIdx.x;
        const int cls = i < n ? class_of(b, cnt[i]) : -1;
        unsigned mine = 0;
        for (int k = 0; k < b.n; k++) {
            const unsigned m = __ballot_sync(0xffffffffu, cls == k);
            if (lane == 0) s_w[warp][k] = __popc(m);
            if (cls == k) mine = m;
        }
        __syncthreads();
        if (threadIdx.x < b.n) {
            const int k = threadIdx.x;
            int tot = 0, base = 0;
            for (int w = 0; w < 8; w++) { const int c = s_w[w][k]; s_w[w][k] = tot; tot += c; }
            for (int c = 0; c < k; c++) base += tile_pos[(int64_t)c * (n_tiles + 1) + n_tiles];     // sizes of the classes before this one
            s_base[k] = base + tile_pos[(int64_t)k * (n_tiles + 1) + t];
        }
        __syncthreads();
        if (cls >= 0) ids[s_base[cls] + s_w[warp][cls] + __popc(mine & ((1u << lane) - 1u))] = (int)i;
        __syncthreads();
    }
}
__global__ void ray_flags_kernel(const uint8_t* __restrict__ sample_valid, int R, int SR, int8_t* __restrict__ ray_mask,
                                 int* __restrict__ flag_i32) {
    // one warp per ray: ray survives iff any of its SR slots has a neighbour (CU:425-427)
    int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= R) return;
    int any = 0;
    for (int s = lane; s < SR; s += 32) any |= sample_valid[(int64_t)r * SR + s];
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) { ray_mask[r] = any ? 1 : 0; flag_i32[r] = any ? 1 : 0; }
}
__global__ void scatter_rays_kernel(const int* __restrict__ flag_pos, const int8_t* __restrict__ ray_mask, int R,
                                    int* __restrict__ ray_index, int* __restrict__ n_rays) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R && ray_mask[r]) ray_index[flag_pos[r]] = r;
    if (r == 0) *n_rays = flag_pos[R];
}
__global__ void gather_rays_kernel(const int* __restrict__ ray_index, int n_rays, int SR, int K,
                                   const int* __restrict__ pidx, const float* __restrict__ loc,
                                   int* __restrict__ out_pidx, float* __restrict__ out_loc) {
    const int per_ray = SR * (K + 3);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_rays * per_ray) return;
    int rr = (int)(i / per_ray), e = (int)(i % per_ray);
    int r = ray_index[rr];
    if (e < SR * K) out_pidx[(int64_t)rr * SR * K + e] = pidx[(int64_t)r * SR * K + e];
    else { e -= SR * K; out_loc[(int64_t)rr * SR * 3 + e] = loc[(int64_t)r * SR * 3 + e]; }
}
// ---- hit-ray compaction: rays whose sample selection found at least one occupied position (R' of SURVEY.md section 8)
__global__ void hit_flags_kernel(const int* __restrict__ cnt, int R, int* __restrict__ flag) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R) flag[r] = cnt[r] > 0 ? 1 : 0;
}
__global__ void scatter_hits_kernel(const int* __restrict__ pos, const int* __restrict__ cnt, int R, int* __restrict__ ray_index,
                                    int* __restrict__ n_rays) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R && cnt[r] > 0) ray_index[pos[r]] = r;
    if (r == 0) *n_rays = pos[R];
}
// one warp per compact ray: its SR sample positions, its count and its direction
__global__ void __launch_bounds__(256) gather_hits_kernel(const int* __restrict__ ray_index, int n, int SR, const float* __restrict__ loc,
                                                           const int* __restrict__ cnt, const float* __restrict__ dirs,
                                                           float* __restrict__ loc_out, int* __restrict__ cnt_out,
                                                           float* __restrict__ dirs_out) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        const int r = ray_index[i];
        const float* src = loc + (int64_t)r * SR * 3;
        float* dst = loc_out + (int64_t)i * SR * 3;
        for (int e = lane; e < 3 * SR; e += 32) dst[e] = src[e];
        if (lane == 0) cnt_out[i] = cnt[r];
        if (lane < 3 && dirs) dirs_out[3 * (int64_t)i + lane] = dirs[3 * (int64_t)r + lane];
    }
}
}  // namespace
}  // namespace pnerf

using namespace pnerf;

extern "C" int pnerf_hit_rays(const int* sample_cnt, int R, int* ray_index, int* n_rays, void* workspace, int64_t workspace_bytes,
                              void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (R < 0 || !n_rays) return PNERF_ERR_ARG;
    if (R == 0) { PNERF_CUDA(cudaMemsetAsync(n_rays, 0, 4, st)); return PNERF_OK; }
    if (!sample_cnt || !ray_index || !workspace) return PNERF_ERR_ARG;
    int64_t pos_bytes = align_up(((int64_t)R + 1) * 4, 256);
    if (workspace_bytes < pos_bytes + scan_workspace_bytes(R)) return PNERF_ERR_WORKSPACE;
    int* pos = (int*)workspace;
    hit_flags_kernel<<<(R + 255) / 256, 256, 0, st>>>(sample_cnt, R, pos);
    PNERF_LAUNCH_CHECK();
    int rc = exclusive_scan_i32(pos, pos, R, true, (char*)workspace + pos_bytes, workspace_bytes - pos_bytes, st);
    if (rc) return rc;
    scatter_hits_kernel<<<(R + 255) / 256, 256, 0, st>>>(pos, sample_cnt, R, ray_index, n_rays);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}

extern "C" int pnerf_gather_hit_rays(const int* ray_index, int n_rays, int SR, const float* sample_loc, const int* sample_cnt,
                                     const float* dirs, float* loc_out, int* cnt_out, float* dirs_out, void* stream) {
    if (n_rays < 0 || SR <= 0) return PNERF_ERR_ARG;
    if (n_rays == 0) return PNERF_OK;
    if (!ray_index || !sample_loc || !sample_cnt || !loc_out || !cnt_out || (dirs && !dirs_out)) return PNERF_ERR_ARG;
    const int64_t b = ((int64_t)n_rays * 32 + 255) / 256;
    const int blocks = (int)(b > (int64_t)kSMs * 16 ? (int64_t)kSMs * 16 : b);
    gather_hits_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ray_index, n_rays, SR, sample_loc, sample_cnt, dirs, loc_out, cnt_out, dirs_out);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}

extern "C" int64_t pnerf_scan_workspace_bytes(int64_t n) { return scan_workspace_bytes(n + 1) + align_up((n + 1) * 4, 256); }

extern "C" int pnerf_sample_compact(const uint8_t* sample_valid, int64_t n_slots, int* sample_ids, int* n_samples,
                                    void* workspace, int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_slots < 0 || !sample_ids || !n_samples) return PNERF_ERR_ARG;
    if (n_slots == 0) { PNERF_CUDA(cudaMemsetAsync(n_samples, 0, 4, st)); return PNERF_OK; }
    if (!sample_valid || !workspace) return PNERF_ERR_ARG;
    int64_t pos_bytes = align_up((n_slots + 1) * 4, 256);
    if (workspace_bytes < pos_bytes + scan_workspace_bytes(n_slots)) return PNERF_ERR_WORKSPACE;
    int* pos = (int*)workspace;
    unsigned blocks = (unsigned)((n_slots + 255) / 256);
    flags_to_i32_kernel<<<blocks, 256, 0, st>>>(sample_valid, n_slots, pos);
    PNERF_LAUNCH_CHECK();
    int rc = exclusive_scan_i32(pos, pos, n_slots, true, (char*)workspace + pos_bytes, workspace_bytes - pos_bytes, st);
    if (rc) return rc;
    scatter_ids_kernel<<<blocks, 256, 0, st>>>(sample_valid, pos, n_slots, sample_ids, n_samples);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}

extern "C" int pnerf_sample_compact_classes(const uint8_t* sample_count, int64_t n_slots, int K, int n_classes, const int* class_kp_h,
                                            int* sample_ids, int* n_per_class, void* workspace, int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_slots < 0 || K <= 0 || n_classes < 1 || n_classes > kMaxClasses || !class_kp_h || !sample_ids || !n_per_class) return PNERF_ERR_ARG;
    for (int c = 0; c < n_classes; c++)
        if (class_kp_h[c] < 1 || (c > 0 && class_kp_h[c] >= class_kp_h[c - 1])) return PNERF_ERR_ARG;      // strictly descending
    if (class_kp_h[0] < K) return PNERF_ERR_ARG;                                                            // the first class takes the fullest samples
    PNERF_CUDA(cudaMemsetAsync(n_per_class, 0, 4 * (size_t)n_classes, st));
    if (n_slots == 0) return PNERF_OK;
    if (!sample_count || !workspace) return PNERF_ERR_ARG;
    ClassBounds b;
    b.n = n_classes;
    for (int c = 0; c < n_classes; c++) { b.hi[c] = c == 0 ? K : class_kp_h[c]; b.lo[c] = c + 1 < n_classes ? class_kp_h[c + 1] : 0; }
    const int64_t n_tiles = (n_slots + 255) / 256;
    const int64_t tc_bytes = align_up((int64_t)n_classes * (n_tiles + 1) * 4, 256);
    if (workspace_bytes < tc_bytes + scan_workspace_bytes(n_tiles + 1)) return PNERF_ERR_WORKSPACE;
    int* tile_counts = (int*)workspace;
    const unsigned blocks = (unsigned)(n_tiles > (int64_t)kSMs * 32 ? (int64_t)kSMs * 32 : n_tiles);
    class_tile_count_kernel<<<blocks, 256, 0, st>>>(sample_count, n_slots, b, n_tiles, tile_counts);
    PNERF_LAUNCH_CHECK();
    for (int c = 0; c < n_classes; c++) {
        int* tc = tile_counts + (int64_t)c * (n_tiles + 1);
        int rc = exclusive_scan_i32(tc, tc, n_tiles, true, (char*)workspace + tc_bytes, workspace_bytes - tc_bytes, st);
        if (rc) return rc;
    }
    class_tile_scatter_kernel<<<blocks, 256, 0, st>>>(sample_count, n_slots, b, n_tiles, tile_counts, sample_ids, n_per_class);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}

extern "C" int pnerf_ray_compact(const uint8_t* sample_valid, int R, int SR, int8_t* ray_mask, int* ray_index,
                                 int* n_rays, void* workspace, int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (R < 0 || SR <= 0 || !n_rays) return PNERF_ERR_ARG;
    if (R == 0) { PNERF_CUDA(cudaMemsetAsync(n_rays, 0, 4, st)); return PNERF_OK; }
    if (!sample_valid || !ray_mask || !ray_index || !workspace) return PNERF_ERR_ARG;
    int64_t pos_bytes = align_up(((int64_t)R + 1) * 4, 256);
    if (workspace_bytes < pos_bytes + scan_workspace_bytes(R)) return PNERF_ERR_WORKSPACE;
    int* pos = (int*)workspace;
    ray_flags_kernel<<<(unsigned)(((int64_t)R * 32 + 255) / 256), 256, 0, st>>>(sample_valid, R, SR, ray_mask, pos);
    PNERF_LAUNCH_CHECK();
    int rc = exclusive_scan_i32(pos, pos, R, true, (char*)workspace + pos_bytes, workspace_bytes - pos_bytes, st);
    if (rc) return rc;
    scatter_rays_kernel<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(pos, ray_mask, R, ray_index, n_rays);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}

extern "C" int pnerf_gather_rays(const int* ray_index, int n_rays, int SR, int K, const int* sample_pidx,
                                 const float* sample_loc, int* out_pidx, float* out_loc, void* stream) {
    if (n_rays < 0) return PNERF_ERR_ARG;
    if (n_rays == 0) return PNERF_OK;
    int64_t total = (int64_t)n_rays * SR * (K + 3);
    gather_rays_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ray_index, n_rays, SR, K, sample_pidx,
                                                                                         sample_loc, out_pidx, out_loc);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}

extern "C" int pnerf_version(void) { return 200; }
extern "C" int pnerf_host_register(void* host_ptr, int64_t bytes) {
    if (!host_ptr || bytes <= 0) return PNERF_ERR_ARG;
    PNERF_CUDA(cudaHostRegister(host_ptr, (size_t)bytes, cudaHostRegisterPortable));
    return PNERF_OK;
}
extern "C" int pnerf_host_unregister(void* host_ptr) {
    if (!host_ptr) return PNERF_ERR_ARG;
    PNERF_CUDA(cudaHostUnregister(host_ptr));
    return PNERF_OK;
}
extern "C" int pnerf_copy_rows_to_host(void* dst_h, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t row_bytes,
                                       int64_t n_rows, void* stream) {
    if (!dst_h || !src || row_bytes <= 0 || n_rows < 0 || dst_pitch < row_bytes || src_pitch < row_bytes) return PNERF_ERR_ARG;
    if (n_rows == 0) return PNERF_OK;
    PNERF_CUDA(cudaMemcpy2DAsync(dst_h, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)row_bytes, (size_t)n_rows,
                                 cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return PNERF_OK;
}
extern "C" const char* pnerf_last_cuda_error(void) { return g_last_error; }
extern "C" int pnerf_device_check(void) {
    int dev = 0;
    cudaDeviceProp p;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) return PNERF_ERR_CUDA;
    return p.major == 10 ? PNERF_OK : PNERF_ERR_ARCH;
}
