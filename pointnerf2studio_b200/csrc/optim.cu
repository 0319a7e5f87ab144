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
}  // namespace
}  // namespace pnerf

using namespace pnerf;

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
