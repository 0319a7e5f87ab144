// Micro-benchmarks of the three resources the fused field kernel leans on, run on all SMs at once (one CTA per SM):
//   which = 0 : tcgen05.mma issue rate, operands resident in shared memory in the K-slab layout (M=128, N=param, K=16)
//   which = 1 : L2 -> shared bulk-copy stream through a ring (chunk bytes = param, 8 stages), source = a 557 KB buffer
//   which = 2 : tcgen05.ld read rate, param warps (4 or 8) each reading its 32 lanes x 256 columns repeatedly
// out[blockIdx.x] = cycles for `iters` operations.  Used by tools/tc_microbench.py; numbers are quoted in DESIGN.md.
#include "pnerf_common.cuh"
#include "umma.cuh"

namespace pnerf {
namespace {
using namespace umma;

__global__ void __launch_bounds__(256, 1) microbench_kernel(int which, int iters, int param, const uint8_t* __restrict__ src,
                                                             unsigned long long* __restrict__ out, float* __restrict__ sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[8], empty[8], done;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&tmem_base, 512);
    if (tid == 0) {
        for (int i = 0; i < 8; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(&done, 1);
        fence_barrier_init();
    }
    for (int i = tid; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    unsigned long long t0 = 0, t1 = 0;
    if (which == 0) {
        if (tid == 0) {
            const int N = param;
            const uint32_t idesc = make_idesc_bf16(128, N);
            const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 16384);
            t0 = clock64();
            for (int i = 0; i < iters; i++) {
                const int ks = i & 3;
                const uint64_t ad = make_smem_desc(a_base + ks * 2 * 2048, 2048, 128);
                const uint64_t bd = make_smem_desc(b_base + ks * 2 * (N * 16), N * 16, 128);
                mma_bf16(tmem + ((i >> 2) & 1) * 256, ad, bd, idesc, 1u);
            }
            mma_commit(&done);
            mbar_wait(&done, 0);
            t1 = clock64();
            out[blockIdx.x] = t1 - t0;
        }
    } else if (which == 5) {
        // MMA stream with a tcgen05.commit to a (never waited) mbarrier after every `param` MMAs
        if (tid == 0) {
            const uint32_t idesc = make_idesc_bf16(128, 256);
            const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 16384);
            t0 = clock64();
            for (int i = 0; i < iters; i++) {
                const int ks = i & 3;
                const uint64_t ad = make_smem_desc(a_base + ks * 2 * 2048, 2048, 128);
                const uint64_t bd = make_smem_desc(b_base + ks * 2 * (256 * 16), 256 * 16, 128);
                mma_bf16(tmem, ad, bd, idesc, 1u);
                if ((i + 1) % param == 0) mma_commit(&full[(i / param) & 7]);
            }
            mma_commit(&done);
            mbar_wait(&done, 0);
            t1 = clock64();
            out[blockIdx.x] = t1 - t0;
        }
    } else if (which == 3 || which == 4) {
        // MMA stream (thread 0 of warp 0) while `param` other warps hammer shared memory with conflict-free 16-byte stores
        // (which = 3) or read the other accumulator with tcgen05.ld (which = 4)
        __shared__ volatile int stop;
        if (tid == 0) stop = 0;
        __syncthreads();
        if (tid == 0) {
            const uint32_t idesc = make_idesc_bf16(128, 256);
            const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 16384);
            t0 = clock64();
            for (int i = 0; i < iters; i++) {
                const int ks = i & 3;
                const uint64_t ad = make_smem_desc(a_base + ks * 2 * 2048, 2048, 128);
                const uint64_t bd = make_smem_desc(b_base + ks * 2 * (256 * 16), 256 * 16, 128);
                mma_bf16(tmem, ad, bd, idesc, 1u);
            }
            mma_commit(&done);
            mbar_wait(&done, 0);
            t1 = clock64();
            out[blockIdx.x] = t1 - t0;
            stop = 1;
        } else if (warp >= 1 && warp <= param) {
            unsigned long long n = 0;
            float acc = 0.f;
            uint4* dst = reinterpret_cast<uint4*>(smem + 65536 + (warp - 1) * 8192) + lane;
            const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256u;
            while (!stop) {
                if (which == 3) {
#pragma unroll
                    for (int j = 0; j < 16; j++) dst[j * 32] = make_uint4(n, j, 0, 0);
                } else {
                    float v[32];
                    tmem_ld32(ta + (uint32_t)((n & 7) * 32), v);
                    tmem_ld_wait();
                    acc += v[3];
                }
                n++;
            }
            if (lane == 0) out[148 + blockIdx.x * 8 + warp] = n;
            if (acc == 12345.f) sink[0] = acc;
        }
    } else if (which == 1) {
        const uint32_t bytes = (uint32_t)param & 0xfffff;
        const int lanes = max(param >> 20, 1);                 // producer lanes: lane l issues chunks i = l (mod lanes)
        const int n_src = 557056 / (int)bytes;
        if (warp == 0 && lane < lanes) {
            t0 = clock64();
            for (int i = lane; i < iters; i += lanes) {
                const uint32_t st = i & 7, ph = (i >> 3) & 1;
                mbar_wait(&empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&full[st], bytes);
                bulk_g2s(smem + 65536 + st * 16384, src + (size_t)(i % n_src) * bytes, bytes, &full[st]);
            }
        } else if (warp == 1 && lane == 0) {
            uint32_t st = 0, ph = 0;
            t0 = clock64();
            for (int i = 0; i < iters; i++) {
                mbar_wait(&full[st], ph);
                mbar_arrive(&empty[st]);
                if (++st == 8) { st = 0; ph ^= 1; }
            }
            t1 = clock64();
            out[blockIdx.x] = t1 - t0;
        }
    } else {
        if (warp < param) {
            float acc = 0.f;
            const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
            __syncwarp();
            t0 = clock64();
            for (int i = 0; i < iters; i++) {
#pragma unroll 1
                for (int c0 = 0; c0 < 256; c0 += 32) {
                    float v[32];
                    tmem_ld32(ta + c0, v);
                    tmem_ld_wait();
                    acc += v[0] + v[31];
                }
            }
            t1 = clock64();
            if (lane == 0 && warp == 0) out[blockIdx.x] = t1 - t0;
            if (acc == 12345.f) sink[0] = acc;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}
}  // namespace
}  // namespace pnerf

using namespace pnerf;

extern "C" int pnerf_tc_microbench(int which, int iters, int param, const void* src, unsigned long long* out, float* sink, void* stream) {
    if (which < 0 || which > 5 || iters <= 0 || !out) return PNERF_ERR_ARG;
    if (which == 1 && (!src || (param & 0xfffff) <= 0 || (param & 0xfffff) > 16384 || (param % 16))) return PNERF_ERR_ARG;
    if (which == 0 && (param < 16 || param > 256 || (param % 16))) return PNERF_ERR_ARG;
    if (which == 2 && (param < 1 || param > 8)) return PNERF_ERR_ARG;
    if ((which == 3 || which == 4) && (param < 0 || param > 7)) return PNERF_ERR_ARG;
    if (which == 5 && param < 1) return PNERF_ERR_ARG;
    const size_t smem = 65536 + 8 * 16384;
    PNERF_CUDA(cudaFuncSetAttribute(microbench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    microbench_kernel<<<kSMs, 256, smem, (cudaStream_t)stream>>>(which, iters, param, (const uint8_t*)src, out, sink);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}
