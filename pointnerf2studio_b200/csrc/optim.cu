// Fused multi-tensor Adam for the two parameter groups of the plugin ("fields" 5e-4, "neural_points" 2e-3, studio_config.py:33-48;
// torch.optim.Adam semantics: betas, eps, bias correction, no weight decay / amsgrad, as nerfstudio's AdamOptimizerConfig builds it).
// After the fused field path the dense Adam over the neural-point tensors (N x 39 floats + two moments: 624 B of traffic per point and
// step) is the largest HBM consumer of a training step; torch runs it as ~10 elementwise launches per tensor.
// Here: ONE launch for all tensors of a group, one read of (p, g, m, v) and one write of (p, m, v) per element, float4 wide.
#include "pnerf_common.cuh"

namespace pnerf {
namespace {
struct Seg { float* p; const float* g; float* m; float* v; int64_t n; float lr_c, inv_bc2_sqrt; };   // lr_c = lr / bias_correction1
struct Segs { Seg s[PNERF_ADAM_MAX_SEGS]; };

// omb1 / omb2 = 1 - beta evaluated in double on the host, as Python does for torch (1.f - 0.999f is off by 5e-5 relative)
__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float lr_c, float omb1, float b2, float omb2, float eps,
                                      float inv_bc2_sqrt) {
    m = fmaf(omb1, g - m, m);                           // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(omb2 * g, g, v * b2);                      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    const float denom = sqrtf(v) * inv_bc2_sqrt + eps;  // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    p = p - lr_c * (m / denom);                         // param.addcdiv_(exp_avg, denom, value=-lr / bias_correction1)
}

__global__ void __launch_bounds__(256) adam_kernel(Segs segs, float omb1, float b2, float omb2, float eps, float gscale) {
    const Seg sg = segs.s[blockIdx.y];
    const float lr_c = sg.lr_c, inv_bc2_sqrt = sg.inv_bc2_sqrt;
    const int64_t n4 = sg.n >> 2;
    const bool vec = ((((uintptr_t)sg.p | (uintptr_t)sg.g | (uintptr_t)sg.m | (uintptr_t)sg.v) & 15) == 0);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        float4* P = (float4*)sg.p; const float4* G = (const float4*)sg.g; float4* M = (float4*)sg.m; float4* V = (float4*)sg.v;
        for (int64_t i = t0; i < n4; i += stride) {
            float4 p = P[i], g = G[i], m = M[i], v = V[i];
            g.x *= gscale; g.y *= gscale; g.z *= gscale; g.w *= gscale;
            adam1(p.x, g.x, m.x, v.x, lr_c, omb1, b2, omb2, eps, inv_bc2_sqrt);
            adam1(p.y, g.y, m.y, v.y, lr_c, omb1, b2, omb2, eps, inv_bc2_sqrt);
            adam1(p.z, g.z, m.z, v.z, lr_c, omb1, b2, omb2, eps, inv_bc2_sqrt);
            adam1(p.w, g.w, m.w, v.w, lr_c, omb1, b2, omb2, eps, inv_bc2_sqrt);
            P[i] = p; M[i] = m; V[i] = v;
        }
    }
    for (int64_t i = (vec ? n4 * 4 : 0) + t0; i < sg.n; i += stride) {
        float p = sg.p[i], m = sg.m[i], v = sg.v[i];
        adam1(p, sg.g[i] * gscale, m, v, lr_c, omb1, b2, omb2, eps, inv_bc2_sqrt);
        sg.p[i] = p; sg.m[i] = m; sg.v[i] = v;
    }
}

// ---------------------------------------------------------------------------------------------- data-parallel step
// Gradient reduction + Adam + parameter broadcast of a ray-sharded training step as ONE kernel over peer-mapped memory (NVLink /
// NVSwitch), instead of NCCL all-reduce followed by a replicated dense Adam (what DDP + torch.optim.Adam do for the reference,
// studio_pipeline.py:48-53): all parameters and gradients of a rank live in two flat buffers with the same layout on every rank;
// rank r owns the slice [lo, hi): it reads that slice of EVERY rank's gradient buffer (local + world-1 peer loads), averages,
// applies Adam with ITS slice of the moments (optimiser state sharded: 1/world of the Adam traffic per GPU) and stores the new
// parameters into EVERY rank's parameter buffer.  Per GPU and step the links carry (world-1)/world of the flat size in each
// direction, half of what a ring / tree all-reduce moves, and the dense Adam pass over all N points disappears from 7 of 8 GPUs.
// The caller brackets the launch with two cross-rank barriers (gradients complete / parameters delivered).
struct DpArgs {
    float* p[PNERF_DP_MAX_RANKS];
    const float* g[PNERF_DP_MAX_RANKS];
    float *m, *v;
    int64_t lo, hi, boundary;
    float lr_c[2], inv_bc2_sqrt;
    const float* hyper_dev;          // {lr_c[0], lr_c[1], inv_bc2_sqrt} in device memory (graph replay), or NULL
    const float* mc_g; float* mc_p;  // NVLS multicast addresses of the gradient / parameter buffers (all ranks at once), or NULL
    int world, me;
};

// NVSwitch multicast ("multimem") forms: ONE load returns the sum over every rank's copy, reduced inside the switch, and ONE store
// lands in every rank's copy -- the links of a GPU carry 2 x 1/world of the flat size instead of 2 x (world-1)/world.
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* addr) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(addr) : "memory");
    return r;
}
__device__ __forceinline__ void multimem_st(float* addr, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int W>
__global__ void __launch_bounds__(512) dp_adam_kernel(const DpArgs a, float omb1, float b2, float omb2, float eps, float gscale) {
    const int64_t n4 = (a.hi - a.lo) >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const float lrc0 = a.hyper_dev ? __ldg(a.hyper_dev) : a.lr_c[0], lrc1 = a.hyper_dev ? __ldg(a.hyper_dev + 1) : a.lr_c[1];
    const float ibc2 = a.hyper_dev ? __ldg(a.hyper_dev + 2) : a.inv_bc2_sqrt;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int64_t e = a.lo + 4 * i;
        float4 g;
        if (W > 1 && a.mc_g) {
            g = multimem_ld_reduce_add(a.mc_g + e);
        } else {
            float4 gs[W];
#pragma unroll
            for (int w = 0; w < W; w++) gs[w] = __ldcs(reinterpret_cast<const float4*>(a.g[w] + e));   // all loads in flight before the first add
            g = gs[0];
#pragma unroll
            for (int w = 1; w < W; w++) { g.x += gs[w].x; g.y += gs[w].y; g.z += gs[w].z; g.w += gs[w].w; }
        }
        g.x *= gscale; g.y *= gscale; g.z *= gscale; g.w *= gscale;
        float4 p = *reinterpret_cast<const float4*>(a.p[a.me] + e);
        float4 m = reinterpret_cast<float4*>(a.m)[i], v = reinterpret_cast<float4*>(a.v)[i];
        const float l0 = e + 0 < a.boundary ? lrc0 : lrc1, l1 = e + 1 < a.boundary ? lrc0 : lrc1;
        const float l2 = e + 2 < a.boundary ? lrc0 : lrc1, l3 = e + 3 < a.boundary ? lrc0 : lrc1;
        adam1(p.x, g.x, m.x, v.x, l0, omb1, b2, omb2, eps, ibc2);
        adam1(p.y, g.y, m.y, v.y, l1, omb1, b2, omb2, eps, ibc2);
        adam1(p.z, g.z, m.z, v.z, l2, omb1, b2, omb2, eps, ibc2);
        adam1(p.w, g.w, m.w, v.w, l3, omb1, b2, omb2, eps, ibc2);
        reinterpret_cast<float4*>(a.m)[i] = m;
        reinterpret_cast<float4*>(a.v)[i] = v;
        if (W > 1 && a.mc_p) {
            multimem_st(a.mc_p + e, p);
        } else {
#pragma unroll
            for (int w = 0; w < W; w++) *reinterpret_cast<float4*>(a.p[w] + e) = p;
        }
    }
}
}  // namespace
}  // namespace pnerf

using namespace pnerf;

extern "C" int pnerf_dp_adam_step(const pnerf_dp_adam* h, float beta1, float beta2, float eps, float grad_scale, void* stream) {
    if (!h || h->world < 1 || h->world > PNERF_DP_MAX_RANKS || h->rank < 0 || h->rank >= h->world || h->step < 1) return PNERF_ERR_ARG;
    if (h->lo < 0 || h->hi < h->lo || ((h->lo | h->hi) & 3) || !h->m || !h->v) return PNERF_ERR_ARG;
    if (h->hi == h->lo) return PNERF_OK;
    DpArgs a;
    // slot 0 = this rank (local loads first), then the peers starting with the next rank: the ranks' peer traffic is spread over all links
    for (int i = 0; i < h->world; i++) {
        const int w = (h->rank + i) % h->world;
        if (!h->p[w] || !h->g[w] || (((uintptr_t)h->p[w] | (uintptr_t)h->g[w]) & 15)) return PNERF_ERR_ARG;
        a.p[i] = h->p[w]; a.g[i] = h->g[w];
    }
    a.me = 0;
    a.hyper_dev = h->hyper_dev;
    a.mc_g = h->mc_g; a.mc_p = h->mc_p;
    if ((a.mc_g == nullptr) != (a.mc_p == nullptr) || (((uintptr_t)a.mc_g | (uintptr_t)a.mc_p) & 15)) return PNERF_ERR_ARG;
    a.m = h->m; a.v = h->v; a.lo = h->lo; a.hi = h->hi; a.boundary = h->boundary; a.world = h->world;
    const double bc1 = 1.0 - pow((double)beta1, (double)h->step), bc2 = 1.0 - pow((double)beta2, (double)h->step);
    a.lr_c[0] = (float)(h->lr[0] / bc1); a.lr_c[1] = (float)(h->lr[1] / bc1); a.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
    const int64_t n4 = (h->hi - h->lo) >> 2;
    int64_t blocks = (n4 + 511) / 512;
    blocks = blocks < 1 ? 1 : (blocks > (int64_t)kSMs * 4 ? (int64_t)kSMs * 4 : blocks);
    const float omb1 = (float)(1.0 - (double)beta1), omb2 = (float)(1.0 - (double)beta2);
    cudaStream_t st = (cudaStream_t)stream;
#define PNERF_DP(W) case W: dp_adam_kernel<W><<<(unsigned)blocks, 512, 0, st>>>(a, omb1, beta2, omb2, eps, grad_scale); break
    switch (h->world) { PNERF_DP(1); PNERF_DP(2); PNERF_DP(3); PNERF_DP(4); PNERF_DP(5); PNERF_DP(6); PNERF_DP(7); PNERF_DP(8); default: return PNERF_ERR_ARG; }
#undef PNERF_DP
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}

extern "C" int pnerf_adam_step(const pnerf_adam_seg* segs_h, int n_segs, float beta1, float beta2, float eps, float grad_scale,
                               void* stream) {
    if (!segs_h || n_segs < 0 || n_segs > PNERF_ADAM_MAX_SEGS) return PNERF_ERR_ARG;
    if (n_segs == 0) return PNERF_OK;
    Segs segs;
    int64_t n_max = 0;
    for (int i = 0; i < n_segs; i++) {
        if (!segs_h[i].p || !segs_h[i].g || !segs_h[i].m || !segs_h[i].v || segs_h[i].n < 0 || segs_h[i].step < 1) return PNERF_ERR_ARG;
        const double bc1 = 1.0 - pow((double)beta1, (double)segs_h[i].step), bc2 = 1.0 - pow((double)beta2, (double)segs_h[i].step);
        segs.s[i] = {segs_h[i].p, segs_h[i].g, segs_h[i].m, segs_h[i].v, segs_h[i].n, (float)(segs_h[i].lr / bc1), (float)(1.0 / sqrt(bc2))};
        n_max = segs_h[i].n > n_max ? segs_h[i].n : n_max;
    }
    int64_t blocks = (n_max / 4 + 255) / 256;
    blocks = blocks < 1 ? 1 : (blocks > (int64_t)kSMs * 16 ? (int64_t)kSMs * 16 : blocks);
    adam_kernel<<<dim3((unsigned)blocks, (unsigned)n_segs), 256, 0, (cudaStream_t)stream>>>(segs, (float)(1.0 - (double)beta1), beta2,
                                                                                          (float)(1.0 - (double)beta2), eps, grad_scale);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}
