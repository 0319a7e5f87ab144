// Step length + alpha compositing, forward and backward (rows D, C, F, L of SURVEY.md section 8a).
// Replaces SM:368-390 (+ nerfstudio RGBRenderer) and fill_invalid SM:491-504; original-flow twin
// NPV:271-279 + ray_march RM:495-541.  One warp per ray: lanes own slots s = lane + 32 c, the
// cummax / exclusive-transmittance cumprod / suffix sums are warp scans with a carry between chunks,
// loads and stores are coalesced along the slot axis.  HBM-bound: 20 B per sample in, 12 B per ray out.
#include "pnerf_common.cuh"

namespace pnerf {
namespace {

constexpr int MAXC = 4;   // SR <= 128

struct CompCam { float o[3]; float Rz[3]; const float* dev; };   // only the camera z axis matters for the depth; dev: pnerf_camera.dev
__device__ __forceinline__ void load_cam(CompCam& c) {
    if (c.dev) {
#pragma unroll
        for (int i = 0; i < 3; i++) { c.o[i] = __ldg(c.dev + i); c.Rz[i] = __ldg(c.dev + 3 + 3 * i + 2); }
    }
}

// z of R_c2w^T (p - o), mul-then-add like SU:140-141
__device__ __forceinline__ float depth_of(const CompCam& c, const float* __restrict__ p) {
    const float sx = __fsub_rn(p[0], c.o[0]), sy = __fsub_rn(p[1], c.o[1]), sz = __fsub_rn(p[2], c.o[2]);
    return __fadd_rn(__fadd_rn(__fmul_rn(sx, c.Rz[0]), __fmul_rn(sy, c.Rz[1])), __fmul_rn(sz, c.Rz[2]));
}

// Per-lane quantities of one ray after the forward recurrences.
struct RayState {
    float alpha[MAXC], T[MAXC], delta[MAXC];   // opacity, exclusive transmittance, step length (already * valid)
    float T_end;
};

__device__ __forceinline__ void ray_forward(const CompCam& cam, float vsize_z, const float* __restrict__ loc,
                                            const uint8_t* __restrict__ valid, const float* __restrict__ sigma, int SR,
                                            int lane, RayState& st) {
    float run_max = -INFINITY, run_T = 1.f;
    const float thr = 2.f * vsize_z;
#pragma unroll
    for (int c = 0; c < MAXC; c++) {
        const int s = c * 32 + lane;
        st.alpha[c] = 0.f; st.T[c] = 1.f; st.delta[c] = 0.f;
        if (c * 32 >= SR) continue;                 // warp-uniform
        const bool in = s < SR;
        // cummax of the depth over slots, including the zero-filled empty slots (SM:368)
        float z = in ? depth_of(cam, loc + 3 * s) : -INFINITY;
        float m = z;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float t = __shfl_up_sync(0xffffffffu, m, o);
            if (lane >= o) m = fmaxf(m, t);
        }
        m = fmaxf(m, run_max);
        run_max = __shfl_sync(0xffffffffu, m, 31);
        float d = 0.f;
        if (in) {
            if (s == SR - 1) d = vsize_z;                                           // SM:369
            else d = __fsub_rn(fmaxf(m, depth_of(cam, loc + 3 * (s + 1))), m);      // m_{s+1} - m_s
            const float msk = (d < 1e-8f || d > thr) ? 1.f : 0.f;                    // SM:371-373
            d = d * (1.f - msk) + msk * vsize_z;                                    // SM:374
            const float v = valid[s] ? 1.f : 0.f;
            d *= v;                                                                  // SM:375
            const float sg = sigma[s] * v;                                           // SM:379
            st.alpha[c] = 1.f - expf(-sg * d);                                       // SM:380
        }
        st.delta[c] = d;
        // exclusive cumprod of (1 - alpha + 1e-10) (SM:382-385)
        float f = in ? (1.f - st.alpha[c] + 1e-10f) : 1.f;
        float inc = f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc *= t;
        }
        float exc = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) exc = 1.f;
        st.T[c] = run_T * exc;
        run_T = run_T * __shfl_sync(0xffffffffu, inc, 31);
    }
    st.T_end = run_T;
}

__global__ void __launch_bounds__(256) composite_fwd_kernel(CompCam cam, pnerf_mode mode, const float* __restrict__ sample_loc,
                                                             const uint8_t* __restrict__ sample_valid,
                                                             const float* __restrict__ sigma, const float* __restrict__ rgb,
                                                             int R, int SR, float* __restrict__ out_rgb,
                                                             float* __restrict__ out_w, float* __restrict__ out_T) {
    load_cam(cam);
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < R; r += warps) {
        const int64_t base = (int64_t)r * SR;
        RayState st;
        ray_forward(cam, mode.vsize_z, sample_loc + 3 * base, sample_valid + base, sigma + base, SR, lane, st);
        float acc[3] = {0.f, 0.f, 0.f}, wsum = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; c++) {
            const int s = c * 32 + lane;
            if (s >= SR) continue;
            const float w = st.alpha[c] * st.T[c];                                   // SM:386
            if (out_w) out_w[base + s] = w;
            wsum += w;
#pragma unroll
            for (int j = 0; j < 3; j++) {
                float v = rgb[3 * (base + s) + j];
                if (mode.eval_clamp && !(v == v)) v = 0.f;                           // nan_to_num (eval)
                acc[j] = fmaf(w, v, acc[j]);
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
#pragma unroll
            for (int j = 0; j < 3; j++) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
        }
        if (lane < 3) {
            const float a = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : acc[2]);
            float v = a + mode.bg[lane] * (mode.bg_mode ? st.T_end : (1.f - wsum));   // RM:529-532 / RGBRenderer
            if (mode.eval_clamp) v = fminf(fmaxf(v, 0.f), 1.f);
            out_rgb[3 * (int64_t)r + lane] = v;
        }
        if (out_T && lane == 0) out_T[r] = st.T_end;
    }
}

// dL/d sigma_s = delta_s (1 - alpha_s) [ T_s u_s - (sum_{j>s} w_j u_j + B) / (1 - alpha_s + 1e-10) ],
//   u_s = <dC, c_s - bg> and B = 0 (plugin)  |  u_s = <dC, c_s> and B = <dC, bg> T_end (original);  dL/d c_s = w_s dC.
__global__ void __launch_bounds__(256) composite_bwd_kernel(CompCam cam, pnerf_mode mode, const float* __restrict__ sample_loc,
                                                             const uint8_t* __restrict__ sample_valid,
                                                             const float* __restrict__ sigma, const float* __restrict__ rgb,
                                                             const float* __restrict__ d_out, int R, int SR,
                                                             float* __restrict__ d_sigma, float* __restrict__ d_rgb) {
    load_cam(cam);
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < R; r += warps) {
        const int64_t base = (int64_t)r * SR;
        RayState st;
        ray_forward(cam, mode.vsize_z, sample_loc + 3 * base, sample_valid + base, sigma + base, SR, lane, st);
        const float dC[3] = {d_out[3 * (int64_t)r], d_out[3 * (int64_t)r + 1], d_out[3 * (int64_t)r + 2]};
        const float dbg = dC[0] * mode.bg[0] + dC[1] * mode.bg[1] + dC[2] * mode.bg[2];
        float carry = mode.bg_mode ? dbg * st.T_end : 0.f;   // everything "behind" the current chunk
#pragma unroll
        for (int c = MAXC - 1; c >= 0; c--) {
            const int s = c * 32 + lane;
            if (c * 32 >= SR) continue;
            const bool in = s < SR;
            float u = 0.f, w = 0.f;
            if (in) {
                w = st.alpha[c] * st.T[c];
                const float cr = rgb[3 * (base + s)], cg = rgb[3 * (base + s) + 1], cb = rgb[3 * (base + s) + 2];
                u = dC[0] * cr + dC[1] * cg + dC[2] * cb - (mode.bg_mode ? 0.f : dbg);
                d_rgb[3 * (base + s)] = w * dC[0];
                d_rgb[3 * (base + s) + 1] = w * dC[1];
                d_rgb[3 * (base + s) + 2] = w * dC[2];
            }
            // inclusive suffix sum of w u over lanes >= lane
            float suf = w * u;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float t = __shfl_down_sync(0xffffffffu, suf, o);
                if (lane + o < 32) suf += t;
            }
            const float behind = suf - w * u + carry;          // strictly after s
            carry += __shfl_sync(0xffffffffu, suf, 0);
            if (in) {
                const float one_m = 1.f - st.alpha[c];
                const float dalpha = st.T[c] * u - behind / (one_m + 1e-10f);
                d_sigma[base + s] = st.delta[c] * one_m * dalpha;   // delta already carries `valid`
            }
        }
    }
}

__global__ void __launch_bounds__(256) conf_loss_kernel(const float* __restrict__ conf, const int* __restrict__ pidx,
                                                         const int8_t* __restrict__ ray_mask, int64_t total, int per_ray,
                                                         float eps, float weight, const int* __restrict__ n_rays,
                                                         float* __restrict__ loss_out, float* __restrict__ g_conf, float gscale) {
    const float inv_n = 1.f / fmaxf((float)((int64_t)(*n_rays) * per_ray), 1.f);
    float part = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        if (!ray_mask[i / per_ray]) continue;
        const int p = max(pidx[i], 0);                                   // invalid slots read point 0 (SU:194)
        const float cs = fminf(fmaxf(conf[p], 1e-4f), 1.f);              // straight-through clamp value (SM:289-292)
        const float val = fminf(fmaxf(cs, eps), 1.f - eps);              // SM:427
        part += logf(val) + logf(1.f - val);
        // gradient: most slots are empty and all of those read point 0 (SU:194) -- combine equal addresses inside the warp first
        // (one atomic per distinct point instead of thousands of serialised atomics on g_conf[0])
        const bool live = g_conf && cs >= eps && cs <= 1.f - eps;
        const float g = live ? gscale * weight * inv_n * (1.f / val - 1.f / (1.f - val)) : 0.f;
        const unsigned act = __activemask();
        const unsigned same = __match_any_sync(act, live ? p : -1);
        float sum = 0.f;
        for (unsigned m = same; m; m &= m - 1) sum += __shfl_sync(same, g, __ffs(m) - 1);
        if (live && (int)(threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(g_conf + p, sum);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0 && part != 0.f && loss_out) atomicAdd(loss_out, part * weight * inv_n);
}

// Masked-ray MSE of get_loss_dict (SM:415-426: MSELoss over the rays with ray_mask > 0, + 1e-6) and its gradient.  The torch
// expression is seven elementwise / reduction launches forward and as many backward; a training step is launch-bound.
// acc[0] = sum m (p - g)^2, acc[1] = sum m, acc[2] = block ticket (all zero on entry); the last block writes the loss.
__global__ void __launch_bounds__(256) masked_mse_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ image,
                                                              const int8_t* __restrict__ ray_mask, int R, float* __restrict__ acc,
                                                              float* __restrict__ loss_out) {
    float se = 0.f, cnt = 0.f;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x) {
        if (ray_mask[r] <= 0) continue;
        const float d0 = pred[3 * r] - image[3 * r], d1 = pred[3 * r + 1] - image[3 * r + 1], d2 = pred[3 * r + 2] - image[3 * r + 2];
        se += d0 * d0 + d1 * d1 + d2 * d2;
        cnt += 1.f;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { se += __shfl_xor_sync(0xffffffffu, se, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
    if ((threadIdx.x & 31) == 0 && cnt > 0.f) { atomicAdd(acc, se); atomicAdd(acc + 1, cnt); }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(acc + 2), 1u);
        if (ticket == gridDim.x - 1) {
            __threadfence();
            const float a = *(volatile float*)acc, n = *(volatile float*)(acc + 1);
            loss_out[0] = a / (3.f * n) + 1e-6f;        // no masked ray: 0 / 0 = NaN, as MSELoss over an empty selection
        }
    }
}
__global__ void __launch_bounds__(256) masked_mse_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ image,
                                                              const int8_t* __restrict__ ray_mask, int R, const float* __restrict__ acc,
                                                              const float* __restrict__ d_loss, float* __restrict__ g_pred) {
    const float k = 2.f * d_loss[0] / (3.f * acc[1]);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * R; i += gridDim.x * blockDim.x)
        g_pred[i] = ray_mask[i / 3] > 0 ? k * (pred[i] - image[i]) : 0.f;
}

// Hole probing (original flow, neural_points_volumetric_model.py:331-362): per ray the sample of largest opacity, its position, the
// distance to its nearest gathered neighbour and the (weight * confidence)-averaged attributes of its K neighbours -- the candidates
// run/train_studio.py:335-444 turns into new neural points.  One warp per ray; lanes k < K own the neighbours of the chosen sample.
struct ProbeOut { float *opacity, *loc, *far_dist, *color, *dir, *conf, *embed; };
__global__ void __launch_bounds__(256) probe_kernel(CompCam cam, pnerf_mode mode, const float* __restrict__ sample_loc,
                                                     const uint8_t* __restrict__ sample_valid, const float* __restrict__ sigma,
                                                     const int* __restrict__ sample_pidx, const float* __restrict__ xyz,
                                                     const float* __restrict__ embed, const float* __restrict__ color,
                                                     const float* __restrict__ dir, const float* __restrict__ conf, int R, int SR, int K,
                                                     ProbeOut o) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < R; r += warps) {
        const int64_t base = (int64_t)r * SR;
        RayState st;
        ray_forward(cam, mode.vsize_z, sample_loc + 3 * base, sample_valid + base, sigma + base, SR, lane, st);
        // torch.max over the SR slots: largest opacity, first index on ties
        float best = -1.f; int bi = 0;
#pragma unroll
        for (int c = 0; c < MAXC; c++) {
            const int s = c * 32 + lane;
            if (s < SR && st.alpha[c] > best) { best = st.alpha[c]; bi = s; }
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        const float* q = sample_loc + 3 * (base + bi);
        const float qx = q[0], qy = q[1], qz = q[2];
        int p = -1; float dist = INFINITY, wraw = 0.f, cc = 1.f;
        if (lane < K) {
            p = sample_pidx[(base + bi) * K + lane];
            const int idx = max(p, 0);                                  // invalid slots gather point 0 (SU:194)
            const float dx = xyz[3 * (int64_t)idx] - qx, dy = xyz[3 * (int64_t)idx + 1] - qy, dz = xyz[3 * (int64_t)idx + 2] - qz;
            dist = sqrtf(dx * dx + dy * dy + dz * dz);
            wraw = p >= 0 ? 1.f / fmaxf(dist, 1e-6f) : 0.f;
            cc = fminf(fmaxf(conf[idx], 1e-4f), 1.f);
        }
        float wsum = wraw, dmin = dist;
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            wsum += __shfl_xor_sync(0xffffffffu, wsum, off);
            dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, off));
        }
        float w = wraw / fmaxf(wsum, 1e-8f);
        if (mode.weight_conf) w *= cc;                                   // the aggregator's returned weight (PA:826)
        const float sw = w * cc;                                         // NPV:340: weight * conf_coefficient
        // averages: lane j < 32 owns embedding dim j; lanes 0..2 colour, 3..5 dir, 6 conf (second pass)
        float e_acc = 0.f, a_acc = 0.f;
        for (int k = 0; k < K; k++) {
            const float swk = __shfl_sync(0xffffffffu, sw, k);
            const int idx = max(__shfl_sync(0xffffffffu, p, k), 0);
            e_acc = fmaf(swk, embed[(int64_t)idx * 32 + lane], e_acc);
            if (lane < 3) a_acc = fmaf(swk, color[3 * (int64_t)idx + lane], a_acc);
            else if (lane < 6) a_acc = fmaf(swk, dir[3 * (int64_t)idx + lane - 3], a_acc);
            else if (lane == 6) a_acc = fmaf(swk, conf[idx], a_acc);                       // the raw gathered confidence (NPV:350,355)
        }
        o.embed[(int64_t)r * 32 + lane] = e_acc;
        if (lane < 3) { o.color[3 * (int64_t)r + lane] = a_acc; o.loc[3 * (int64_t)r + lane] = lane == 0 ? qx : (lane == 1 ? qy : qz); }
        else if (lane < 6) o.dir[3 * (int64_t)r + lane - 3] = a_acc;
        else if (lane == 6) o.conf[r] = a_acc;
        else if (lane == 7) { o.opacity[r] = best; o.far_dist[r] = dmin; }
    }
}

CompCam make_ccam(const pnerf_camera* c) {
    CompCam k;
    for (int i = 0; i < 3; i++) { k.o[i] = c->origin[i]; k.Rz[i] = c->R_c2w[3 * i + 2]; }
    k.dev = c->dev;
    return k;
}
int ray_blocks(int R) {
    int64_t b = ((int64_t)R * 32 + 255) / 256;
    int64_t cap = (int64_t)kSMs * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
}  // namespace
}  // namespace pnerf

using namespace pnerf;

extern "C" int pnerf_composite_forward(const pnerf_camera* cam, const pnerf_mode* mode, const float* sample_loc,
                                       const uint8_t* sample_valid, const float* sigma, const float* rgb, int R, int SR,
                                       float* out_rgb, float* out_weights, float* out_T_end, void* stream) {
    if (!cam || !mode || R < 0 || SR <= 0 || SR > 32 * MAXC) return PNERF_ERR_ARG;
    if (R == 0) return PNERF_OK;
    if (!sample_loc || !sample_valid || !sigma || !rgb || !out_rgb) return PNERF_ERR_ARG;
    composite_fwd_kernel<<<ray_blocks(R), 256, 0, (cudaStream_t)stream>>>(make_ccam(cam), *mode, sample_loc, sample_valid, sigma,
                                                                         rgb, R, SR, out_rgb, out_weights, out_T_end);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}

extern "C" int pnerf_composite_backward(const pnerf_camera* cam, const pnerf_mode* mode, const float* sample_loc,
                                        const uint8_t* sample_valid, const float* sigma, const float* rgb, const float* d_out,
                                        int R, int SR, float* d_sigma, float* d_rgb, void* stream) {
    if (!cam || !mode || R < 0 || SR <= 0 || SR > 32 * MAXC) return PNERF_ERR_ARG;
    if (R == 0) return PNERF_OK;
    if (!sample_loc || !sample_valid || !sigma || !rgb || !d_out || !d_sigma || !d_rgb) return PNERF_ERR_ARG;
    composite_bwd_kernel<<<ray_blocks(R), 256, 0, (cudaStream_t)stream>>>(make_ccam(cam), *mode, sample_loc, sample_valid, sigma,
                                                                         rgb, d_out, R, SR, d_sigma, d_rgb);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}

extern "C" int pnerf_probe(const pnerf_points* pts, const pnerf_camera* cam, const pnerf_mode* mode, const float* sample_loc,
                           const uint8_t* sample_valid, const float* sigma, const int* sample_pidx, int R, int SR, int K,
                           float* max_opacity, float* max_loc, float* far_dist, float* avg_color, float* avg_dir, float* avg_conf,
                           float* avg_embed, void* stream) {
    if (!pts || !cam || !mode || R < 0 || SR <= 0 || SR > 32 * MAXC || K <= 0 || K > 32) return PNERF_ERR_ARG;
    if (R == 0) return PNERF_OK;
    if (!sample_loc || !sample_valid || !sigma || !sample_pidx || !max_opacity || !max_loc || !far_dist || !avg_color || !avg_dir ||
        !avg_conf || !avg_embed)
        return PNERF_ERR_ARG;
    ProbeOut o = {max_opacity, max_loc, far_dist, avg_color, avg_dir, avg_conf, avg_embed};
    probe_kernel<<<ray_blocks(R), 256, 0, (cudaStream_t)stream>>>(make_ccam(cam), *mode, sample_loc, sample_valid, sigma, sample_pidx,
                                                                 pts->xyz, pts->embed, pts->color, pts->dir, pts->conf, R, SR, K, o);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}

extern "C" int pnerf_masked_mse_forward(const float* pred, const float* image, const int8_t* ray_mask, int R, float* acc, float* loss_out,
                                       void* stream) {
    if (R <= 0 || !pred || !image || !ray_mask || !acc || !loss_out) return PNERF_ERR_ARG;
    const int blocks = min(kSMs * 2, (R + 255) / 256);
    masked_mse_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pred, image, ray_mask, R, acc, loss_out);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}

extern "C" int pnerf_masked_mse_backward(const float* pred, const float* image, const int8_t* ray_mask, int R, const float* acc,
                                        const float* d_loss, float* g_pred, void* stream) {
    if (R <= 0 || !pred || !image || !ray_mask || !acc || !d_loss || !g_pred) return PNERF_ERR_ARG;
    const int blocks = min(kSMs * 4, (3 * R + 255) / 256);
    masked_mse_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pred, image, ray_mask, R, acc, d_loss, g_pred);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}

extern "C" int pnerf_conf_loss(const float* conf, const int* sample_pidx, const int8_t* ray_mask, int R, int SR, int K,
                               float eps, float weight, const int* n_rays, float* loss_out, float* g_conf, float grad_scale,
                               void* stream) {
    if (R < 0 || SR <= 0 || K <= 0) return PNERF_ERR_ARG;
    if (R == 0) return PNERF_OK;
    if (!conf || !sample_pidx || !ray_mask || !n_rays) return PNERF_ERR_ARG;
    const int64_t total = (int64_t)R * SR * K;
    const int blocks = (int)min((int64_t)kSMs * 8, (total + 255) / 256);
    conf_loss_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(conf, sample_pidx, ray_mask, total, SR * K, eps, weight, n_rays,
                                                              loss_out, g_conf, grad_scale);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}
