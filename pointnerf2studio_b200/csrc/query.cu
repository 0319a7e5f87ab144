// Sample selection (rows G0/G2) and layered K-nearest query (row Q) -- replaces near_far_linear_ray_generation,
// mask_raypos, get_shadingloc, query_neigh_along_ray_layered and the torch glue between them
// (diff_ray_marching.py:292-336, query_worldcoords.cu:165-302, 368-422).
//
// sample_select: one warp per ray.  A lane owns 4 consecutive coarse positions of each 128-position chunk; the
//   jittered t mid-points are generated in registers (Philox4x32-10 keyed by (seed, ray, j/4), warp scan of the
//   segment lengths), the ray is clipped against the grid box first so only positions inside it are probed
//   against the occupancy bitmask (L1/L2 resident), and the first SR hits are compacted with a warp prefix sum.
//   No (R,D,3) position tensor, no (R,D) mask, no cumsum / masked_select, no host sync.
// query: one THREAD per sample slot, a warp = 32 consecutive slots of one ray (neighbouring samples share
//   cells -> L1 hits).  Candidates come as runs of 16-byte records sorted by (cell, index) (see grid.cu); the
//   thread keeps its K best in registers as a sorted list with a branch-free unrolled insertion.  Shell by shell,
//   stop when >= K in-radius candidates have been seen (CU:300).
#include "pnerf_common.cuh"

namespace pnerf {
namespace {

__device__ __forceinline__ bool occ_probe(const Frame& f, const uint32_t* __restrict__ occ, float x, float y, float z) {
    int vx, vy, vz;
    if (!voxel_of(f, x, y, z, vx, vy, vz)) return false;
    int id = cell_lin(f, vx, vy, vz);
    return (__ldg(occ + (id >> 5)) >> (id & 31)) & 1u;
}

// ------------------------------------------------------------------------------------------------ coarse t
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }   // [0,1), 24 bits

struct TGen { float near, far, jitter; uint32_t seed_lo, seed_hi; };

// t edge j of RM:313-315: near * (1 - tau_j) + far * tau_j, tau = linspace(0, 1, D + 1) (torch: start + i*step for the
// first half, end - (steps-1-i)*step for the second)
__device__ __forceinline__ float t_edge(const TGen& g, int j, int D) {
    const float step = 1.f / (float)D;
    const float tau = (j < (D + 1) / 2) ? __fmul_rn(step, (float)j) : __fsub_rn(1.f, __fmul_rn(step, (float)(D - j)));
    return __fadd_rn(__fmul_rn(g.near, __fsub_rn(1.f, tau)), __fmul_rn(g.far, tau));
}

// The four t mid-points [jb, jb+4) of ray r (jb = j0 + 4*lane) of RM:312-329 with jitter; `carry` = sum of the segment
// lengths before j0 (warp-uniform), updated to include this chunk.  Positions with j >= D get t = +inf.
__device__ __forceinline__ void chunk_t(const TGen& g, int r, int jb, int D, int lane, float& carry, float (&t)[4]) {
    float seg[4];
    const uint4 rnd = philox4x32_10(make_uint4((uint32_t)(jb >> 2), (uint32_t)r, 0u, 0u), make_uint2(g.seed_lo, g.seed_hi));
    const uint32_t rv[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
    float e0 = jb < D ? t_edge(g, jb, D) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int j = jb + i;
        float s = 0.f;
        if (j < D) {
            const float e1 = t_edge(g, j + 1, D);
            s = __fmul_rn(__fsub_rn(e1, e0), __fadd_rn(1.f, __fmul_rn(g.jitter, __fsub_rn(u01(rv[i]), 0.5f))));   // RM:318-322
            e0 = e1;
        }
        seg[i] = s;
    }
    const float tot = (seg[0] + seg[1]) + (seg[2] + seg[3]);
    float inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    float run = carry + (inc - tot);                 // cumsum before this lane's first segment
    carry += __shfl_sync(0xffffffffu, inc, 31);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float end0 = g.near + run;             // RM:324-326
        run += seg[i];
        const float end1 = g.near + run;
        t[i] = (jb + i < D) ? 0.5f * (end0 + end1) : INFINITY;   // RM:327
    }
}

// Ray / grid-box slab test with half a voxel of margin: positions with t outside [ta, tb] are outside the grid and can
// never hit (CU:181-185), so they are not probed.
__device__ __forceinline__ void clip_ray(const Frame& f, const float o[3], const float d[3], float& ta, float& tb) {
    ta = -INFINITY; tb = INFINITY;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float lo = f.lo[a] - 0.5f * f.sv[a], hi = f.lo[a] + ((float)f.dim[a] + 0.5f) * f.sv[a];
        if (fabsf(d[a]) > 1e-12f) {
            const float t1 = (lo - o[a]) / d[a], t2 = (hi - o[a]) / d[a];
            ta = fmaxf(ta, fminf(t1, t2));
            tb = fminf(tb, fmaxf(t1, t2));
        } else if (o[a] < lo || o[a] > hi) {
            ta = INFINITY; tb = -INFINITY;
        }
    }
    const float pad = 1e-4f * fmaxf(fabsf(ta), fabsf(tb));
    if (ta <= tb) { ta -= pad; tb += pad; }
}

// MODE 0: explicit positions raypos (R,D,3); 1: t table (t_stride 0 or D); 2: jittered t generated in registers
#ifndef PNERF_SEL_MINB
#define PNERF_SEL_MINB 5      // in-kernel jitter: 61 -> 51 registers, 5 blocks per SM: 0.526 -> 0.474 ms on one box
#endif
template <int MODE>
__global__ void __launch_bounds__(256, MODE == 2 ? PNERF_SEL_MINB : 1) sample_select_kernel(Frame f, const uint32_t* __restrict__ occ,
                                                             const float* __restrict__ raypos, float ox, float oy, float oz,
                                                             const float* __restrict__ dirs, const float* __restrict__ t_vals,
                                                             int t_stride, TGen gen, int R, int D, int SR, int fill_missed,
                                                             float* __restrict__ sample_loc, int* __restrict__ sample_cnt,
                                                             const float* __restrict__ step_dev) {
    const int lane = threadIdx.x & 31;
    if (step_dev) {          // per-step constants from device memory (CUDA-graph replay): pnerf_camera.dev layout
        ox = __ldg(step_dev); oy = __ldg(step_dev + 1); oz = __ldg(step_dev + 2);
        gen.near = __ldg(step_dev + 12); gen.far = __ldg(step_dev + 13);
        gen.seed_lo = __float_as_uint(__ldg(step_dev + 14)); gen.seed_hi = __float_as_uint(__ldg(step_dev + 15));
    }
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < R; r += warps) {
        float d[3] = {0.f, 0.f, 0.f};
        float ta = -INFINITY, tb = INFINITY;
        if (MODE != 0) {
            d[0] = __ldg(dirs + 3 * (int64_t)r); d[1] = __ldg(dirs + 3 * (int64_t)r + 1); d[2] = __ldg(dirs + 3 * (int64_t)r + 2);
            const float o[3] = {ox, oy, oz};
            clip_ray(f, o, d, ta, tb);
        }
        float* loc = sample_loc + (int64_t)r * SR * 3;
        int n = 0;
        float carry = 0.f;
        if (ta <= tb) {
            for (int j0 = 0; j0 < D && n < SR; j0 += 128) {
                const int jb = j0 + 4 * lane;
                float t[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
                if (MODE == 2) {
                    if (gen.near + carry > tb) break;                     // every later position is behind the box (warp-uniform)
                    chunk_t(gen, r, jb, D, lane, carry, t);
                } else if (MODE == 1) {
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        if (jb + i < D) t[i] = __ldg(t_vals + (int64_t)r * t_stride + jb + i);
                }
                float px[4], py[4], pz[4];
                unsigned hits = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    bool probe = jb + i < D;
                    if (MODE == 0) {
                        if (probe) {
                            const float* p = raypos + ((int64_t)r * D + jb + i) * 3;
                            px[i] = __ldg(p); py[i] = __ldg(p + 1); pz[i] = __ldg(p + 2);
                        }
                    } else {
                        probe = probe && t[i] >= ta && t[i] <= tb;
                        px[i] = __fadd_rn(ox, __fmul_rn(d[0], t[i]));     // campos + raydir * t, two roundings (RM:330)
                        py[i] = __fadd_rn(oy, __fmul_rn(d[1], t[i]));
                        pz[i] = __fadd_rn(oz, __fmul_rn(d[2], t[i]));
                    }
                    if (probe && occ_probe(f, occ, px[i], py[i], pz[i])) hits |= 1u << i;
                }
                if (__ballot_sync(0xffffffffu, hits != 0) == 0) continue;
                const int c = __popc(hits);
                int inc = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                int slot = n + inc - c;
                n += __shfl_sync(0xffffffffu, inc, 31);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if ((hits >> i) & 1u) {
                        if (slot < SR) { loc[3 * slot] = px[i]; loc[3 * slot + 1] = py[i]; loc[3 * slot + 2] = pz[i]; }
                        slot++;
                    }
                }
            }
        }
        n = min(n, SR);
        if (n > 0 || fill_missed)
            for (int e = 3 * n + lane; e < 3 * SR; e += 32) loc[e] = 0.f;   // unfilled slots stay (0,0,0) (CU:383)
        if (lane == 0) sample_cnt[r] = n;
    }
}

// Test / oracle hook: the t table (R,D) and the uniforms (R,D) that MODE 2 uses, computed by the same device code.
__global__ void __launch_bounds__(256) coarse_t_kernel(TGen gen, int R, int D, float* __restrict__ t_out, float* __restrict__ u_out) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < R; r += warps) {
        float carry = 0.f;
        for (int j0 = 0; j0 < D; j0 += 128) {
            const int jb = j0 + 4 * lane;
            float t[4];
            chunk_t(gen, r, jb, D, lane, carry, t);
            const uint4 rnd = philox4x32_10(make_uint4((uint32_t)(jb >> 2), (uint32_t)r, 0u, 0u), make_uint2(gen.seed_lo, gen.seed_hi));
            const uint32_t rv[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (jb + i < D) {
                    t_out[(int64_t)r * D + jb + i] = t[i];
                    if (u_out) u_out[(int64_t)r * D + jb + i] = u01(rv[i]);
                }
        }
    }
}

// ------------------------------------------------------------------------------------------------ query
// Sorted K-best list in registers; strict '<' keeps the earlier-visited candidate ahead on equal d2, so the list is
// ordered by (d2, visit order) -- the tie-break rule of include/pnerf_b200.h.
template <int KMAX>
__device__ __forceinline__ void list_insert(float (&bd2)[KMAX], int (&bidx)[KMAX], float d2, int idx) {
#pragma unroll
    for (int i = KMAX - 1; i > 0; --i) {
        const bool shift = d2 < bd2[i - 1];
        const bool here = !shift && d2 < bd2[i];
        bidx[i] = shift ? bidx[i - 1] : (here ? idx : bidx[i]);
        bd2[i] = shift ? bd2[i - 1] : (here ? d2 : bd2[i]);
    }
    if (d2 < bd2[0]) { bd2[0] = d2; bidx[0] = idx; }
}

// Per shell the thread first lists its non-empty record runs (start, length) in shared memory, then the warp walks all
// candidate streams in lock step: the trip count is the warp's longest stream and the sorted insertion runs branch-free
// for every lane (a rejected or missing candidate carries d2 = +inf and changes nothing).  The nested per-run loops this
// replaces executed sum_runs max_lane(len) iterations with a divergent 48-instruction insertion: 979 M warp instructions
// for the render bench.
// Tried in round 2 and dropped: skipping runs whose distance lower bound (gaps to the sample's voxel faces) is not below the K-th best
// distance -- exact, and it removes ~20 % of the record loads, but the bound bookkeeping (a third shared-memory word per run, 20 B of
// spills under the 56-register cap, a data-dependent loop instead of the precomputed trip count) cost more than the loads it saved:
// 1.51 -> 1.72 ms on the render bench.
// K <= 8: 62 registers leave 8 blocks (32 warps) per SM; capping at 56 (9 blocks) costs no spills and hides more of the record
// loads' latency: -11 % on the render bench (tools/sweep_query.sh on one box: 1.65 / 1.46 / 1.53 ms at 8 / 9 / 10 blocks)
#ifndef PNERF_Q_MINB
#define PNERF_Q_MINB 9
#endif
#ifndef PNERF_Q_MINB16
#define PNERF_Q_MINB16 6      // K = 16 (stress config, 10 M points): 93 -> 85 registers, 15.0 -> 13.6 ms per 4.9 M samples
#endif
template <int KMAX>
__global__ void __launch_bounds__(128, KMAX <= 8 ? PNERF_Q_MINB : (KMAX == 16 ? PNERF_Q_MINB16 : 1)) query_kernel(Frame f, const int* __restrict__ cell_start,
                                                     const float4* __restrict__ recs, const float* __restrict__ sample_loc,
                                                     const int* __restrict__ sample_cnt, int R, int SR, int K, int layers,
                                                     float r2, int max_runs, int* __restrict__ sample_pidx,
                                                     uint8_t* __restrict__ sample_valid, unsigned long long* __restrict__ stats,
                                                     int across_rays) {
    extern __shared__ int s_runs[];                    // [max_runs][2][128]: start / length of run j of thread t at (j*2 + {0,1})*128 + t
    const int lane = threadIdx.x & 31, t = threadIdx.x;
    // Which 32 samples share a warp.  along a ray (across_rays = 0): 32 consecutive slots of one ray -- right for a batch of unrelated
    // rays (training: random pixels).  across rays (1): the SAME slot of 32 consecutive rays -- right for the hit-ray list of an image,
    // whose neighbours are neighbouring pixels: their s-th samples sit at nearly the same depth, so the 32 candidate streams are of
    // similar length (the lock-step trip count is the warp's longest stream: 58 against a mean of 34 along a ray, where a warp spans
    // the whole passage through the surface shell) and lie closer together (32 pixels ~ 14 voxels against 40 along the ray).
    const int cpr = (SR + 31) >> 5;                    // 32-slot chunks per ray
    const int64_t n_tasks = across_rays ? (int64_t)((R + 31) >> 5) * SR : (int64_t)R * cpr;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long n_vis = 0, n_cand = 0;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_tasks; w += warps) {
        const int r = across_rays ? (int)(w / SR) * 32 + lane : (int)(w / cpr);
        const int slot = across_rays ? (int)(w % SR) : (int)(w % cpr) * 32 + lane;
        const bool in_range = r < R && slot < SR;
        const int64_t sid = (int64_t)r * SR + slot;
        float bd2[KMAX];
        int bidx[KMAX];
#pragma unroll
        for (int i = 0; i < KMAX; i++) { bd2[i] = INFINITY; bidx[i] = -1; }
        const int cnt = r < R ? __ldg(sample_cnt + r) : 0;
        if (__any_sync(0xffffffffu, in_range && slot < cnt)) {      // some sample of this warp is filled
            float qx = 0.f, qy = 0.f, qz = 0.f;
            int vx = 0, vy = 0, vz = 0;
            bool searching = false;
            if (in_range && slot < cnt) {
                qx = __ldg(sample_loc + 3 * sid); qy = __ldg(sample_loc + 3 * sid + 1); qz = __ldg(sample_loc + 3 * sid + 2);
                searching = voxel_of(f, qx, qy, qz, vx, vy, vz);
            }
            int seen = 0;
            for (int shell = 0; shell < layers; shell++) {
                // ---- phase 1: this thread's runs of the shell, in visit order (ux, uy, uz)
                int nr = 0, total = 0;
                if (searching) {
                    for (int dx = -shell; dx <= shell; dx++) {
                        const int x = vx + dx;
                        if (x < 0 || x >= f.dim[0]) continue;
                        for (int dy = -shell; dy <= shell; dy++) {
                            const int y = vy + dy;
                            if (y < 0 || y >= f.dim[1]) continue;
                            const bool rim = max(abs(dx), abs(dy)) == shell;   // whole z range belongs to this shell
                            // rim rows: one run z in [vz-shell, vz+shell]; inner rows: two single cells z = vz -+ shell
                            const int parts = rim ? 1 : 2;
                            for (int part = 0; part < parts; part++) {
                                int z0, z1;
                                if (rim) { z0 = vz - shell; z1 = vz + shell; }
                                else { z0 = z1 = part == 0 ? vz - shell : vz + shell; }
                                z0 = max(z0, 0); z1 = min(z1, f.dim[2] - 1);
                                if (z0 > z1) continue;
                                const int c0 = cell_lin(f, x, y, z0);
                                const int a = __ldg(cell_start + c0), b = __ldg(cell_start + c0 + (z1 - z0) + 1);
                                n_vis += (unsigned long long)(z1 - z0 + 1);
                                n_cand += (unsigned long long)(b - a);
                                if (b > a) {
                                    s_runs[(nr * 2) * 128 + t] = a;
                                    s_runs[(nr * 2 + 1) * 128 + t] = b - a;
                                    nr++; total += b - a;
                                }
                            }
                        }
                    }
                }
                // ---- phase 2: all candidate streams of the warp in lock step
                const int trips = __reduce_max_sync(0xffffffffu, total);
                int run = 0, i = 0, rem = 0;
                if (nr > 0) { i = s_runs[t]; rem = s_runs[128 + t]; }
                for (int it = 0; it < trips; it++) {
                    float d2 = INFINITY;
                    int idx = -1;
                    if (it < total) {
                        const float4 rec = __ldg(recs + i);
                        const float ex = __fsub_rn(rec.x, qx), ey = __fsub_rn(rec.y, qy), ez = __fsub_rn(rec.z, qz);
                        const float dd = __fmaf_rn(ez, ez, __fmaf_rn(ey, ey, __fmul_rn(ex, ex)));   // CU:271 as nvcc contracts it
                        if (r2 == 0.f || dd <= r2) { seen++; d2 = dd; idx = __float_as_int(rec.w) & 0x0fffffff; }
                        i++; rem--;
                        if (rem == 0) {                          // next run of this thread's list
                            run++;
                            if (run < nr) {
                                const int* e = s_runs + (run * 2) * 128 + t;
                                i = e[0]; rem = e[128];
                            }
                        }
                    }
                    if (__any_sync(0xffffffffu, d2 < bd2[KMAX - 1])) list_insert<KMAX>(bd2, bidx, d2, idx);
                }
                if (seen >= K) searching = false;   // CU:300: this sample stops after the layer; the warp goes on for the others
                if (!__any_sync(0xffffffffu, searching)) break;
            }
        }
        // The insertion keeps exactly equal distances in visit order, which is what decides WHO is among the K nearest (CU:274-293:
        // a later candidate replaces the farthest only if it is strictly closer).  The emitted order is the canonical (d2, point index)
        // of the oracle: adjacent entries of the first K with equal d2 are put in ascending index -- one bubble pass, repeated only
        // while some lane of the warp still swapped (ties are rare: coincident points).
        for (;;) {
            bool swapped = false;
#pragma unroll
            for (int i = 0; i + 1 < KMAX; i++) {
                const bool sw = i + 1 < K && bd2[i] == bd2[i + 1] && bidx[i] > bidx[i + 1] && bidx[i + 1] >= 0;
                const int a = bidx[i], b = bidx[i + 1];
                bidx[i] = sw ? b : a; bidx[i + 1] = sw ? a : b;
                swapped |= sw;
            }
            if (!__any_sync(0xffffffffu, swapped)) break;
        }
        if (in_range) {
            int* out = sample_pidx + sid * K;
            if (K == KMAX && (KMAX % 4) == 0) {
#pragma unroll
                for (int i = 0; i < KMAX; i += 4)
                    *reinterpret_cast<int4*>(out + i) = make_int4(bidx[i], bidx[i + 1], bidx[i + 2], bidx[i + 3]);
            } else {
#pragma unroll
                for (int i = 0; i < KMAX; i++)
                    if (i < K) out[i] = bidx[i];
            }
            int nv = 0;                      // number of neighbours found: the valid entries are the first ones of the list
#pragma unroll
            for (int i = 0; i < KMAX; i++) nv += (i < K && bidx[i] >= 0) ? 1 : 0;
            sample_valid[sid] = (uint8_t)nv;
        }
    }
    if (stats) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            n_vis += __shfl_xor_sync(0xffffffffu, n_vis, o);
            n_cand += __shfl_xor_sync(0xffffffffu, n_cand, o);
        }
        if (lane == 0 && (n_vis | n_cand)) { atomicAdd(stats, n_vis); atomicAdd(stats + 1, n_cand); }
    }
    (void)max_runs;
}

int ray_warps_grid(int R) {
    const int64_t b = ((int64_t)R * 32 + 255) / 256;
    const int64_t cap = (int64_t)kSMs * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace
}  // namespace pnerf

using namespace pnerf;

extern "C" int pnerf_sample_select(const pnerf_grid_view* g, const float* raypos, const float* origin_h, const float* dirs,
                                   const float* t_vals, int t_stride, int R, int D, int SR, int fill_missed, float* sample_loc,
                                   int* sample_cnt, void* stream) {
    if (!g || R < 0 || D <= 0 || SR <= 0) return PNERF_ERR_ARG;
    if (R == 0) return PNERF_OK;
    if (!sample_loc || !sample_cnt || !g->occ_bits) return PNERF_ERR_ARG;
    if (!raypos && (!origin_h || !dirs || !t_vals || (t_stride != 0 && t_stride != D))) return PNERF_ERR_ARG;
    const Frame f = frame_of(g);
    const TGen gen = {0.f, 0.f, 0.f, 0u, 0u};
    cudaStream_t st = (cudaStream_t)stream;
    if (raypos)
        sample_select_kernel<0><<<ray_warps_grid(R), 256, 0, st>>>(f, g->occ_bits, raypos, 0.f, 0.f, 0.f, nullptr, nullptr, 0, gen, R, D, SR,
                                                                  fill_missed, sample_loc, sample_cnt, nullptr);
    else
        sample_select_kernel<1><<<ray_warps_grid(R), 256, 0, st>>>(f, g->occ_bits, nullptr, origin_h[0], origin_h[1], origin_h[2], dirs,
                                                                  t_vals, t_stride, gen, R, D, SR, fill_missed, sample_loc, sample_cnt, nullptr);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}

extern "C" int pnerf_sample_select_jitter(const pnerf_grid_view* g, const float* origin_h, const float* dirs, float near, float far,
                                          float jitter, uint64_t seed, int R, int D, int SR, int fill_missed, float* sample_loc,
                                          int* sample_cnt, void* stream) {
    if (!g || R < 0 || D <= 0 || SR <= 0 || !(far > near)) return PNERF_ERR_ARG;
    if (R == 0) return PNERF_OK;
    if (!sample_loc || !sample_cnt || !g->occ_bits || !origin_h || !dirs) return PNERF_ERR_ARG;
    const Frame f = frame_of(g);
    const TGen gen = {near, far, jitter, (uint32_t)seed, (uint32_t)(seed >> 32)};
    sample_select_kernel<2><<<ray_warps_grid(R), 256, 0, (cudaStream_t)stream>>>(f, g->occ_bits, nullptr, origin_h[0], origin_h[1],
                                                                                origin_h[2], dirs, nullptr, 0, gen, R, D, SR,
                                                                                fill_missed, sample_loc, sample_cnt, nullptr);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}

extern "C" int pnerf_sample_select_jitter_dev(const pnerf_grid_view* g, const float* step_dev, const float* dirs, float jitter, int R, int D,
                                              int SR, int fill_missed, float* sample_loc, int* sample_cnt, void* stream) {
    if (!g || R < 0 || D <= 0 || SR <= 0 || !step_dev) return PNERF_ERR_ARG;
    if (R == 0) return PNERF_OK;
    if (!sample_loc || !sample_cnt || !g->occ_bits || !dirs) return PNERF_ERR_ARG;
    const Frame f = frame_of(g);
    const TGen gen = {0.f, 1.f, jitter, 0u, 0u};
    sample_select_kernel<2><<<ray_warps_grid(R), 256, 0, (cudaStream_t)stream>>>(f, g->occ_bits, nullptr, 0.f, 0.f, 0.f, dirs, nullptr, 0, gen, R, D,
                                                                                SR, fill_missed, sample_loc, sample_cnt, step_dev);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}

extern "C" int pnerf_coarse_t(float near, float far, float jitter, uint64_t seed, int R, int D, float* t_out, float* u_out,
                              void* stream) {
    if (R < 0 || D <= 0 || !(far > near)) return PNERF_ERR_ARG;
    if (R == 0) return PNERF_OK;
    if (!t_out) return PNERF_ERR_ARG;
    const TGen gen = {near, far, jitter, (uint32_t)seed, (uint32_t)(seed >> 32)};
    coarse_t_kernel<<<ray_warps_grid(R), 256, 0, (cudaStream_t)stream>>>(gen, R, D, t_out, u_out);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}

extern "C" int pnerf_query(const pnerf_grid_view* g, const float* sample_loc, const int* sample_cnt, int R, int SR, int K,
                           int kernel_size0, float radius, int* sample_pidx, uint8_t* sample_valid,
                           unsigned long long* stats, int rays_are_neighbours, void* stream) {
    if (!g || R < 0 || SR <= 0 || K <= 0 || K > 32) return PNERF_ERR_ARG;
    const int layers = (kernel_size0 + 1) / 2;   // CU:256 reads kernel_size[0] only
    if (layers < 1 || layers > 3) return PNERF_ERR_ARG;
    if (R == 0) return PNERF_OK;
    if (!sample_loc || !sample_cnt || !sample_pidx || !sample_valid || !g->cell_start || !g->recs) return PNERF_ERR_ARG;
    const Frame f = frame_of(g);
    const float r2 = radius * radius;            // CU:410, fp32 on the host
    const int64_t tasks = rays_are_neighbours ? (int64_t)((R + 31) / 32) * SR : (int64_t)R * ((SR + 31) / 32);
    const int blocks = (int)min((int64_t)kSMs * 16, (tasks * 32 + 127) / 128);
    cudaStream_t st = (cudaStream_t)stream;
    const float4* recs = (const float4*)g->recs;
    const int max_runs = layers == 1 ? 1 : (layers == 2 ? 10 : 34);     // runs of the largest shell: 1, 8 + 2, 16 + 18
    const size_t smem = (size_t)max_runs * 2 * 128 * sizeof(int);
#define PNERF_LAUNCH_Q(KM) \
    query_kernel<KM><<<blocks, 128, smem, st>>>(f, g->cell_start, recs, sample_loc, sample_cnt, R, SR, K, layers, r2, max_runs, sample_pidx, sample_valid, stats, rays_are_neighbours ? 1 : 0)
    if (K <= 4) PNERF_LAUNCH_Q(4);
    else if (K <= 8) PNERF_LAUNCH_Q(8);
    else if (K <= 16) PNERF_LAUNCH_Q(16);
    else PNERF_LAUNCH_Q(32);
#undef PNERF_LAUNCH_Q
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}
