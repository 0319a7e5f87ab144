// Sample selection (rows G0/G2) and layered K-nearest query (row Q) -- replaces mask_raypos,
// get_shadingloc, query_neigh_along_ray_layered and the torch glue between them
// (query_worldcoords.cu:165-302, 368-422).
//
// sample_select: one warp per ray walks the D coarse positions 32 at a time, probes the occupancy
//   bitmask (L1/L2 resident), ballots the hits and compacts the first SR of them with popc prefixes.
//   No (R,D) mask tensor, no cumsum, no masked_select, no host sync.
// query: a group of Kp lanes (Kp = 8, 16 or 32 >= K) owns one sample, so a warp serves 32/Kp
//   consecutive slots of one ray (neighbouring samples share cells -> L1 hits).  Candidates come as
//   coalesced runs of 16-byte records (see grid.cu); the group keeps its K best sorted across lanes and
//   inserts candidates with one ballot + one shuffle-up.  Shell by shell, stop when >= K in-radius
//   candidates have been seen (CU:300).
#include "pnerf_common.cuh"

namespace pnerf {
namespace {

__device__ __forceinline__ bool occ_probe(const Frame& f, const uint32_t* __restrict__ occ, float x, float y, float z) {
    int vx, vy, vz;
    if (!voxel_of(f, x, y, z, vx, vy, vz)) return false;
    int id = cell_lin(f, vx, vy, vz);
    return (__ldg(occ + (id >> 5)) >> (id & 31)) & 1u;
}

__global__ void __launch_bounds__(256) sample_select_kernel(Frame f, const uint32_t* __restrict__ occ,
                                                             const float* __restrict__ raypos, float ox, float oy, float oz,
                                                             const float* __restrict__ dirs, const float* __restrict__ t_vals,
                                                             int t_stride, int R, int D, int SR,
                                                             float* __restrict__ sample_loc, int* __restrict__ sample_cnt) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < R; r += warps) {
        float dx = 0.f, dy = 0.f, dz = 0.f;
        if (!raypos) { dx = dirs[3 * (int64_t)r]; dy = dirs[3 * (int64_t)r + 1]; dz = dirs[3 * (int64_t)r + 2]; }
        float* loc = sample_loc + (int64_t)r * SR * 3;
        int n = 0;
        for (int j0 = 0; j0 < D && n < SR; j0 += 32) {
            const int j = j0 + lane;
            float x = 0.f, y = 0.f, z = 0.f;
            bool hit = false;
            if (j < D) {
                if (raypos) {
                    const float* p = raypos + ((int64_t)r * D + j) * 3;
                    x = p[0]; y = p[1]; z = p[2];
                } else {
                    const float t = t_vals[(int64_t)r * t_stride + j];
                    x = __fadd_rn(ox, __fmul_rn(dx, t));   // campos + raydir * t, two roundings (RM:330)
                    y = __fadd_rn(oy, __fmul_rn(dy, t));
                    z = __fadd_rn(oz, __fmul_rn(dz, t));
                }
                hit = occ_probe(f, occ, x, y, z);
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            const int slot = n + __popc(m & ((1u << lane) - 1u));
            if (hit && slot < SR) { loc[3 * slot] = x; loc[3 * slot + 1] = y; loc[3 * slot + 2] = z; }
            n += __popc(m);
        }
        n = min(n, SR);
        for (int e = 3 * n + lane; e < 3 * SR; e += 32) loc[e] = 0.f;   // unfilled slots stay (0,0,0) (CU:383)
        if (lane == 0) sample_cnt[r] = n;
    }
}

// ------------------------------------------------------------------------------------------------ query
struct Best {          // one entry of the group's sorted K-best list, distributed one per lane
    uint32_t d2;       // float bits of d2 (>= 0, so unsigned order == float order); 0xffffffff = empty
    uint32_t ord;      // visit order inside the sample: (shell << 24) | (run << 12) | position
    int idx;
};

template <int KP>
__device__ __forceinline__ void group_insert(Best& b, uint32_t cd2, uint32_t cord, int cidx, unsigned gmask, int glane,
                                             int gshift) {
    const bool before = (b.d2 < cd2) || (b.d2 == cd2 && b.ord < cord);   // my entry stays ahead of the candidate
    const int pos = __popc((__ballot_sync(gmask, before) >> gshift) & (KP == 32 ? 0xffffffffu : ((1u << KP) - 1u)));
    const uint32_t ud2 = __shfl_up_sync(gmask, b.d2, 1, KP);
    const uint32_t uord = __shfl_up_sync(gmask, b.ord, 1, KP);
    const int uidx = __shfl_up_sync(gmask, b.idx, 1, KP);
    if (glane > pos) { b.d2 = ud2; b.ord = uord; b.idx = uidx; }
    else if (glane == pos) { b.d2 = cd2; b.ord = cord; b.idx = cidx; }
}

// Scan one run of records [a, b) for the group's sample.
template <int KP>
__device__ __forceinline__ void scan_run(const float4* __restrict__ recs, int a, int b, float qx, float qy, float qz,
                                         float r2, uint32_t ord_base, int K, Best& best, int& seen, unsigned gmask,
                                         int glane, int gshift, unsigned long long& n_cand) {
    for (int base = a; base < b; base += KP) {
        const int i = base + glane;
        uint32_t cd2 = 0xffffffffu;
        int cidx = -1;
        if (i < b) {
            const float4 rec = __ldg(recs + i);
            const float dx = __fsub_rn(rec.x, qx), dy = __fsub_rn(rec.y, qy), dz = __fsub_rn(rec.z, qz);
            const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));   // CU:271 as nvcc contracts it
            if (r2 == 0.f || d2 <= r2) { cd2 = __float_as_uint(d2); cidx = __float_as_int(rec.w) & 0x0fffffff; }
        }
        unsigned m = (__ballot_sync(gmask, cidx >= 0) >> gshift) & (KP == 32 ? 0xffffffffu : ((1u << KP) - 1u));
        seen += __popc(m);
        n_cand += (glane == 0) ? (unsigned long long)(min(b, base + KP) - base) : 0ull;
        while (m) {   // group-uniform loop: insert the in-radius candidates in visit order
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t sd2 = __shfl_sync(gmask, cd2, src, KP);
            const int sidx = __shfl_sync(gmask, cidx, src, KP);
            const uint32_t sord = ord_base + (uint32_t)(base - a + src);
            // cheap reject: not better than the current K-th (lane K-1 holds it)
            const uint32_t wd2 = __shfl_sync(gmask, best.d2, K - 1, KP);
            const uint32_t word = __shfl_sync(gmask, best.ord, K - 1, KP);
            if (sd2 < wd2 || (sd2 == wd2 && sord < word)) group_insert<KP>(best, sd2, sord, sidx, gmask, glane, gshift);
        }
    }
}

template <int KP>
__global__ void __launch_bounds__(256) query_kernel(Frame f, const int* __restrict__ cell_start,
                                                     const float4* __restrict__ recs, const float* __restrict__ sample_loc,
                                                     const int* __restrict__ sample_cnt, int R, int SR, int K, int layers,
                                                     float r2, int* __restrict__ sample_pidx, uint8_t* __restrict__ sample_valid,
                                                     unsigned long long* __restrict__ stats) {
    constexpr int GPW = 32 / KP;                       // groups (samples) per warp
    const int lane = threadIdx.x & 31;
    const int glane = lane % KP, gid = lane / KP, gshift = gid * KP;
    const unsigned gmask = (KP == 32) ? 0xffffffffu : (((1u << KP) - 1u) << gshift);
    const int groups_per_ray = (SR + GPW - 1) / GPW;   // warp w serves slots [w%gpr * GPW, +GPW) of ray w/gpr
    const int64_t n_warps_work = (int64_t)R * groups_per_ray;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long n_vis = 0, n_cand = 0;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_warps_work; w += warps) {
        const int r = (int)(w / groups_per_ray);
        const int slot = (int)(w % groups_per_ray) * GPW + gid;
        if (slot >= SR) continue;                      // group-uniform
        const int64_t sid = (int64_t)r * SR + slot;
        Best best = {0xffffffffu, 0xffffffffu, -1};
        int seen = 0;
        if (slot < sample_cnt[r]) {
            const float qx = sample_loc[3 * sid], qy = sample_loc[3 * sid + 1], qz = sample_loc[3 * sid + 2];
            int vx, vy, vz;
            if (voxel_of(f, qx, qy, qz, vx, vy, vz)) {
                for (int shell = 0; shell < layers; shell++) {
                    uint32_t run = 0;
                    for (int dx = -shell; dx <= shell; dx++) {
                        const int x = vx + dx;
                        for (int dy = -shell; dy <= shell; dy++) {
                            const int y = vy + dy;
                            const bool in_xy = (x >= 0) && (x < f.dim[0]) && (y >= 0) && (y < f.dim[1]);
                            const bool rim = max(abs(dx), abs(dy)) == shell;   // whole z range belongs to this shell
                            // rim rows: one run z in [vz-shell, vz+shell]; inner rows: two single cells z = vz -+ shell
                            const int parts = rim ? 1 : 2;
                            for (int part = 0; part < parts; part++, run++) {
                                if (!in_xy) continue;
                                int z0, z1;
                                if (rim) { z0 = vz - shell; z1 = vz + shell; }
                                else { z0 = z1 = part == 0 ? vz - shell : vz + shell; }
                                z0 = max(z0, 0); z1 = min(z1, f.dim[2] - 1);
                                if (z0 > z1) continue;
                                const int c0 = cell_lin(f, x, y, z0);
                                const int a = __ldg(cell_start + c0), b = __ldg(cell_start + c0 + (z1 - z0) + 1);
                                if (glane == 0) n_vis += (unsigned long long)(z1 - z0 + 1);
                                scan_run<KP>(recs, a, b, qx, qy, qz, r2, ((uint32_t)shell << 24) | (run << 12), K, best, seen,
                                             gmask, glane, gshift, n_cand);
                            }
                        }
                    }
                    if (seen >= K) break;   // CU:300
                }
            }
        }
        if (glane < K) sample_pidx[sid * K + glane] = best.idx;
        if (glane == 0) sample_valid[sid] = best.idx >= 0 ? 1 : 0;
    }
    if (stats) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            n_vis += __shfl_xor_sync(0xffffffffu, n_vis, o);
            n_cand += __shfl_xor_sync(0xffffffffu, n_cand, o);
        }
        if (lane == 0 && (n_vis | n_cand)) { atomicAdd(stats, n_vis); atomicAdd(stats + 1, n_cand); }
    }
}

}  // namespace
}  // namespace pnerf

using namespace pnerf;

extern "C" int pnerf_sample_select(const pnerf_grid_view* g, const float* raypos, const float* origin_h, const float* dirs,
                                   const float* t_vals, int t_stride, int R, int D, int SR, float* sample_loc,
                                   int* sample_cnt, void* stream) {
    if (!g || R < 0 || D <= 0 || SR <= 0) return PNERF_ERR_ARG;
    if (R == 0) return PNERF_OK;
    if (!sample_loc || !sample_cnt || !g->occ_bits) return PNERF_ERR_ARG;
    if (!raypos && (!origin_h || !dirs || !t_vals || (t_stride != 0 && t_stride != D))) return PNERF_ERR_ARG;
    const Frame f = frame_of(g);
    const int blocks = (int)min((int64_t)kSMs * 8, ((int64_t)R * 32 + 255) / 256);
    sample_select_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(f, g->occ_bits, raypos, raypos ? 0.f : origin_h[0],
                                                                  raypos ? 0.f : origin_h[1], raypos ? 0.f : origin_h[2], dirs,
                                                                  t_vals, t_stride, R, D, SR, sample_loc, sample_cnt);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}

extern "C" int pnerf_query(const pnerf_grid_view* g, const float* sample_loc, const int* sample_cnt, int R, int SR, int K,
                           int kernel_size0, float radius, int* sample_pidx, uint8_t* sample_valid,
                           unsigned long long* stats, void* stream) {
    if (!g || R < 0 || SR <= 0 || K <= 0 || K > 32) return PNERF_ERR_ARG;
    const int layers = (kernel_size0 + 1) / 2;   // CU:256 reads kernel_size[0] only
    if (layers < 1 || layers > 3) return PNERF_ERR_ARG;
    if (R == 0) return PNERF_OK;
    if (!sample_loc || !sample_cnt || !sample_pidx || !sample_valid || !g->cell_start || !g->recs) return PNERF_ERR_ARG;
    const Frame f = frame_of(g);
    const float r2 = radius * radius;            // CU:410, fp32 on the host
    const int KP = K <= 8 ? 8 : (K <= 16 ? 16 : 32);
    const int64_t warps = (int64_t)R * ((SR + 32 / KP - 1) / (32 / KP));
    const int blocks = (int)min((int64_t)kSMs * 16, (warps * 32 + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    const float4* recs = (const float4*)g->recs;
    if (KP == 8)
        query_kernel<8><<<blocks, 256, 0, st>>>(f, g->cell_start, recs, sample_loc, sample_cnt, R, SR, K, layers, r2, sample_pidx, sample_valid, stats);
    else if (KP == 16)
        query_kernel<16><<<blocks, 256, 0, st>>>(f, g->cell_start, recs, sample_loc, sample_cnt, R, SR, K, layers, r2, sample_pidx, sample_valid, stats);
    else
        query_kernel<32><<<blocks, 256, 0, st>>>(f, g->cell_start, recs, sample_loc, sample_cnt, R, SR, K, layers, r2, sample_pidx, sample_valid, stats);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}
