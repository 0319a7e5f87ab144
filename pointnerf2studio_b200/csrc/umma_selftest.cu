// Self-test of the tcgen05 plumbing in umma.cuh: D[128xN] = A[128xK] * W[NxK]^T with bf16 operands, fp32
// accumulation in TMEM.  A is staged by the threads (generic proxy + fence.proxy.async), W arrives as one
// bulk async copy of a pre-packed K-slab buffer: the same two mechanisms the fused field kernel uses.
#include "pnerf_common.cuh"
#include "umma.cuh"

namespace pnerf {
namespace {
using namespace umma;

__global__ void __launch_bounds__(128) umma_selftest_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ Wp,
                                                             float* __restrict__ D, int N, int K, uint32_t tmem_cols) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                               // 128 x K
    uint8_t* sW = smem + 128 * K * 2;                 // N x K
    __shared__ uint64_t bar_w, bar_mma;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base, tmem_cols);
    if (tid == 0) { mbar_init(&bar_w, 1); mbar_init(&bar_mma, 1); fence_barrier_init(); }
    for (int j = 0; j < K / 8; j++)                   // row `tid`, slab j: 8 bf16 = 16 B
        *reinterpret_cast<uint4*>(sA + slab_off(128, tid, j * 8)) = *reinterpret_cast<const uint4*>(A + (size_t)tid * K + j * 8);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tacc = tmem_base;
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar_w, (uint32_t)(N * K * 2));
        bulk_g2s(sW, Wp, (uint32_t)(N * K * 2), &bar_w);
        mbar_wait(&bar_w, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(128, N);
        for (int ks = 0; ks < K / 16; ks++) {
            const uint64_t ad = make_smem_desc(smem_u32(sA) + ks * 2 * (128 * 16), 128 * 16, 128);
            const uint64_t bd = make_smem_desc(smem_u32(sW) + ks * 2 * (N * 16), N * 16, 128);
            mma_bf16(tacc, ad, bd, idesc, ks > 0);
        }
        mma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_ld32(tacc + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j++) D[(size_t)tid * N + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tacc, tmem_cols);
}
}  // namespace
}  // namespace pnerf

using namespace pnerf;

// A: bf16 [128 x K] row major; Wp: bf16 K-slab packed [K/8][N][8]; D: fp32 [128 x N].  N in {32..256} mult of 16, K mult of 16.
extern "C" int pnerf_umma_selftest(const void* A, const void* Wp, float* D, int N, int K, void* stream) {
    if (!A || !Wp || !D || N < 16 || N > 256 || (N % 16) || K < 16 || (K % 16)) return PNERF_ERR_ARG;
    const size_t smem = (size_t)(128 + N) * K * 2;
    if (smem > 226 * 1024) return PNERF_ERR_ARG;
    uint32_t cols = 32;
    while ((int)cols < N) cols <<= 1;
    PNERF_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)Wp, D, N, K, cols);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}
