// Tensor-core training path of the per-neighbour networks (row BWD of SURVEY.md 8a): what torch autograd does for the
// reference through SU:190-209 and SM:270-353 -- dgrad and wgrad of mlp_base / mlp_head, the density head, the K-aggregation
// and the scatter-add of the row gradients into the neural-point tensors -- as bf16 tcgen05 GEMMs over the operand tiles the
// forward kernel kept (field_tc.cu, SAVE mode), fp32 accumulation in TMEM.
//
//   agg_bwd_kernel   (SIMT)  delta4 = (w dF_s + w d sigma a' wa) * lrelu'(h4)  per row; d wa, d ba; d w (confidence, original flow)
//   tile_gemm_kernel (TC)    delta_{L-1} = lrelu'(h_{L-1}) * (delta_L W_L)   per 128-row tile: the layer's transposed weights stay
//                            resident in shared memory (128 KB), a tile arrives as ONE 64 KB bulk copy, two TMEM accumulators
//                            so a tile's epilogue overlaps the next tile's copy + MMAs
//   wgrad_tc_kernel  (TC)    dW_L = delta_L^T X_{L-1}: both operands are the saved tiles read as MN-major operands (the k-slab
//                            layout IS the canonical no-swizzle MN-major layout with LBO and SBO swapped), K = the 128 rows of a
//                            tile, accumulated over a CTA's tiles in TMEM (128 x 288 fp32), fp32 red.add into dW at the end
//                            the bias gradients d b_L = delta_L^T 1 ride in the same GEMM (a constant-1 input column)
//   scatter_kernel   (SIMT)  d embed (raw + through the positional encoding), d color, d dir, d conf -> atomics by point id
// The colour network (mlp_color + rgb head) takes the same route on 128-sample tiles kept by color_tc_kernel<SAVE>: color_head_bwd_kernel
// (SIMT) -> tile_gemm_kernel x3 (K = 128) -> the same wgrad launch (job table).
#include "pnerf_common.cuh"
#include "tc_layout.cuh"
#include "umma.cuh"

namespace pnerf {
namespace {
using namespace umma;
using namespace tcl;

constexpr int HID = 256;
constexpr int NX0 = 224;                       // columns of the layer-1 input that carry a gradient: feat 32 + PE(feat) 192
constexpr int HC = 128;
constexpr int64_t WB4 = 0, WB3 = 131072, WB2 = 262144, WB1 = 393216, WBC3 = WB1 + 32 * NX0 * 16, WBC2 = WBC3 + 16 * HC * 16,
                  WBC1 = WBC2 + 16 * HC * 16, WBWD_BYTES = WBC1 + 16 * HID * 16;

struct Cam { float o[3]; float Rc[9]; float Rw[9]; };
Cam make_cam(const pnerf_points* pts, const pnerf_camera* cam) {
    Cam c;
    for (int i = 0; i < 3; i++) c.o[i] = cam->origin[i];
    for (int i = 0; i < 9; i++) { c.Rc[i] = cam->R_c2w[i]; c.Rw[i] = pts->Rw2c[i]; }
    return c;
}
__device__ __forceinline__ void rot_w2c(const Cam& c, const float* u, float* out) {
#pragma unroll
    for (int j = 0; j < 3; j++) out[j] = u[0] * c.Rw[3 * j] + u[1] * c.Rw[3 * j + 1] + u[2] * c.Rw[3 * j + 2];
}
__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; i++) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ uint4 pack8(const float* v) {
    uint4 r;
    r.x = pack_bf16(v[0], v[1]); r.y = pack_bf16(v[2], v[3]); r.z = pack_bf16(v[4], v[5]); r.w = pack_bf16(v[6], v[7]);
    return r;
}

// Device-side sample count: the host launches with a CAPACITY (grid, workspace carve) and never reads S back; every kernel clamps its
// tile loop to the tiles that hold valid samples (spt samples per tile: 128 / KP for the per-neighbour tiles, 128 for the colour tiles).
__device__ __forceinline__ int dyn_count(const int* n_dev, int cap) { return n_dev ? min(__ldg(n_dev), cap) : cap; }
__device__ __forceinline__ int dyn_tiles(const int* n_dev, int cap_tiles, int spt) {
    return n_dev ? min((__ldg(n_dev) + spt - 1) / spt, cap_tiles) : cap_tiles;
}

// ---------------------------------------------------------------------------------------------- transposed weight pack
// Bt[k/8][n][8] with Bt(n, k) = W[k][n]: k = output feature of the forward layer (256), n = input feature (first `n_cols`).
struct PackT { const float* w; int in_dim, n_cols; int64_t off; int out_dim; };
struct PackTs { PackT j[7]; };
__global__ void __launch_bounds__(256) pack_bwd_kernel(PackTs jobs, uint8_t* __restrict__ dst) {
    const PackT jb = jobs.j[blockIdx.y];
    const int total = jb.out_dim * jb.n_cols;
    __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst + jb.off);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int n = i % jb.n_cols, k = i / jb.n_cols;
        d[((int64_t)(k >> 3) * jb.n_cols + n) * 8 + (k & 7)] = __float2bfloat16(jb.w[(int64_t)k * jb.in_dim + n]);
    }
}

// ---------------------------------------------------------------------------------------------- aggregation backward
struct AggBwd {
    const uint8_t* save; const float *save_w, *save_raw;
    const int* sample_ids; const float* d_sigma; const float* dF; int ldF;
    const float* wa;
    int S, KP, n_tiles, softplus; float slope; const int* S_dev;
    uint8_t* d4; float *dwa, *dba, *dw_rows;
};
__global__ void __launch_bounds__(128) agg_bwd_kernel(const AggBwd p) {
    __shared__ float s_wa[HID];
    __shared__ float s_draw[ROWS];
    const int row = threadIdx.x;
    s_wa[row] = p.wa[row]; s_wa[row + 128] = p.wa[row + 128];
    float acc_wa[2] = {0.f, 0.f}, acc_ba = 0.f;
    const int spt = ROWS / p.KP;
    const int S = dyn_count(p.S_dev, p.S), n_tiles = dyn_tiles(p.S_dev, p.n_tiles, spt);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();
        const int64_t grow = (int64_t)tile * ROWS + row;
        const int si = tile * spt + row / p.KP;
        const float w = p.save_w[grow], raw = p.save_raw[grow];
        const bool live = si < S;
        const float ds = live ? p.d_sigma[p.sample_ids[si]] : 0.f;
        const float a = p.softplus ? softplus_f(raw - 1.f) : fmaxf(raw, 0.f);
        const float dact = p.softplus ? sigmoid_f(raw - 1.f) : (raw > 0.f ? 1.f : 0.f);
        const float draw = ds * w * dact;
        const float4* dFrow = reinterpret_cast<const float4*>(p.dF + (int64_t)(live ? si : 0) * p.ldF);
        const uint4* h4 = reinterpret_cast<const uint4*>(p.save + (int64_t)tile * SAVE_TILE_BYTES + (int64_t)SAVE_H4 * SLAB + row * 16);
        uint4* out = reinterpret_cast<uint4*>(p.d4 + (int64_t)tile * DELTA_TILE_BYTES + row * 16);
        float dwk = 0.f;
#pragma unroll 4
        for (int j = 0; j < HID / 8; j++) {
            float h[8], g[8];
            unpack8(h4[j * (SLAB / 16)], h);
            float4 d0 = make_float4(0.f, 0.f, 0.f, 0.f), d1 = d0;
            if (live) { d0 = __ldg(dFrow + 2 * j); d1 = __ldg(dFrow + 2 * j + 1); }
            const float df[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
            for (int i = 0; i < 8; i++) {
                dwk = fmaf(df[i], h[i], dwk);
                g[i] = fmaf(w, df[i], draw * s_wa[8 * j + i]) * (h[i] > 0.f ? 1.f : p.slope);
            }
            out[j * (SLAB / 16)] = pack8(g);
        }
        p.dw_rows[grow] = dwk + ds * a;
        s_draw[row] = draw;
        __syncthreads();
        // d wa[c] += sum_rows draw_r h4[r][c]: thread t owns columns 2t, 2t+1
        const __nv_bfloat16* ht = reinterpret_cast<const __nv_bfloat16*>(p.save + (int64_t)tile * SAVE_TILE_BYTES + (int64_t)SAVE_H4 * SLAB);
        const int c = 2 * row;
        const __nv_bfloat16* col = ht + (c >> 3) * (SLAB / 2) + (c & 7);
        float s0 = 0.f, s1 = 0.f, sb = 0.f;
        for (int r = 0; r < ROWS; r++) {
            const float d = s_draw[r];
            const __nv_bfloat162 hv = *reinterpret_cast<const __nv_bfloat162*>(col + r * 8);
            s0 = fmaf(d, __low2float(hv), s0);
            s1 = fmaf(d, __high2float(hv), s1);
            sb += d;
        }
        acc_wa[0] += s0; acc_wa[1] += s1; acc_ba += sb;
    }
    if (p.dwa) {
        atomicAdd(p.dwa + 2 * row, acc_wa[0]); atomicAdd(p.dwa + 2 * row + 1, acc_wa[1]);
        if (row == 0) atomicAdd(p.dba, acc_ba);
    }
}

// ---------------------------------------------------------------------------------------------- dgrad tile GEMM
struct GemmP {
    const uint8_t* in; int64_t in_stride;            // 128 x K bf16 operand tiles (K-major k-slabs), K = 8 * ks (256 or 128)
    const uint8_t* mask; int64_t mask_stride;        // tiles whose sign gives lrelu' (same layout), or NULL
    uint8_t* out_bf; int64_t out_stride;             // bf16 tile output, or
    float* out_f32; int ld_f32;                      // fp32 row-major output (row = tile * 128 + r)
    const uint8_t* w; int N; int ks; int n_tiles; float slope;
    const int* S_dev; int spt;                       // device-side sample count and samples per tile (n_tiles is the capacity)
};
struct GemmSmem {
    uint8_t W[32 * HID * 16];
    uint8_t A[DELTA_TILE_BYTES];
    uint64_t w_bar, a_full, a_free, acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

// 12 warps: warp 0 producer + MMA issuer, warps 4-7 / 8-11 the epilogue of the lower / upper column half (a warp reaches the TMEM
// lanes of its quarter warp % 4 only, so two warps share a quarter and split the columns: the epilogue -- 32 mask loads, 256
// multiplies, 32 stores per row -- was the longest stage of a tile with one warp per quarter).
__global__ void __launch_bounds__(384, 1) tile_gemm_kernel(const GemmP p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    GemmSmem& sm = *reinterpret_cast<GemmSmem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = dyn_tiles(p.S_dev, p.n_tiles, p.spt);
    const int n_my = n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (tid == 0) {
        mbar_init(&sm.w_bar, 1); mbar_init(&sm.a_full, 1); mbar_init(&sm.a_free, 1);
        for (int b = 0; b < 2; b++) { mbar_init(&sm.acc_full[b], 1); mbar_init(&sm.acc_empty[b], 8); }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&sm.tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    const int N = p.N;
    if (warp == 0) {
        if (lane == 0 && n_my > 0) {
            const uint32_t wbytes = (uint32_t)(p.ks * N * 16), abytes = (uint32_t)(p.ks * SLAB);
            mbar_arrive_expect_tx(&sm.w_bar, wbytes);
            for (int q = 0; q < 4; q++) bulk_g2s(sm.W + q * (wbytes / 4), p.w + q * (wbytes / 4), wbytes / 4, &sm.w_bar);
            mbar_wait(&sm.w_bar, 0);
            const uint32_t idesc = make_idesc_bf16(ROWS, N);
            const uint32_t a_base = smem_u32(sm.A), w_base = smem_u32(sm.W);
            for (int i = 0; i < n_my; i++) {
                const int tile = (int)blockIdx.x + i * (int)gridDim.x, b = i & 1;
                if (i > 0) mbar_wait(&sm.a_free, (uint32_t)((i - 1) & 1));
                mbar_arrive_expect_tx(&sm.a_full, abytes);
                bulk_g2s(sm.A, p.in + (int64_t)tile * p.in_stride, abytes, &sm.a_full);
                mbar_wait(&sm.a_full, (uint32_t)(i & 1));
                if (i >= 2) mbar_wait(&sm.acc_empty[b], (uint32_t)(((i >> 1) - 1) & 1));
                tc_fence_after();
                for (int ks = 0; ks < p.ks / 2; ks++) {
                    const uint64_t ad = make_smem_desc(a_base + (uint32_t)(ks * 2 * SLAB), SLAB, 128);
                    const uint64_t bd = make_smem_desc(w_base + (uint32_t)(ks * 2 * N * 16), (uint32_t)(N * 16), 128);
                    mma_bf16(tmem + (uint32_t)(b * HID), ad, bd, idesc, (uint32_t)(ks > 0));
                }
                mma_commit(&sm.a_free);
                mma_commit(&sm.acc_full[b]);
            }
        }
    } else if (warp >= 4) {
        const int row = (warp & 3) * 32 + lane;
        const int split = N > 128 ? 128 : 64;                    // N = 256 / 224 / 128: 128 + 128, 128 + 96, 64 + 64 columns
        const int cbeg = warp >= 8 ? split : 0, cend = warp >= 8 ? N : split;
        for (int i = 0; i < n_my; i++) {
            const int tile = (int)blockIdx.x + i * (int)gridDim.x, b = i & 1;
            mbar_wait(&sm.acc_full[b], (uint32_t)((i >> 1) & 1));
            tc_fence_after();
            const uint32_t tacc = tmem + (uint32_t)(b * HID) + ((uint32_t)((warp & 3) * 32) << 16);
            const uint4* mrow = p.mask ? reinterpret_cast<const uint4*>(p.mask + (int64_t)tile * p.mask_stride + row * 16) : nullptr;
            uint4* orow = p.out_bf ? reinterpret_cast<uint4*>(p.out_bf + (int64_t)tile * p.out_stride + row * 16) : nullptr;
            float* frow = p.out_f32 ? p.out_f32 + ((int64_t)tile * ROWS + row) * p.ld_f32 : nullptr;
#pragma unroll 1
            for (int c0 = cbeg; c0 < cend; c0 += 32) {
                float v[32];
                tmem_ld32(tacc + (uint32_t)c0, v);
                tmem_ld_wait();
                if (mrow) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        float h[8];
                        unpack8(__ldg(mrow + (c0 / 8 + j) * (SLAB / 16)), h);
#pragma unroll
                        for (int e = 0; e < 8; e++) v[8 * j + e] *= (h[e] > 0.f ? 1.f : p.slope);
                    }
                }
                if (orow) {
#pragma unroll
                    for (int j = 0; j < 4; j++) orow[(c0 / 8 + j) * (SLAB / 16)] = pack8(v + 8 * j);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; j++) *reinterpret_cast<float4*>(frow + c0 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.acc_empty[b]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------- wgrad
struct WgJob {
    const uint8_t* d; int64_t d_stride; int d_slab0;          // gradient tiles: base, tile stride, first k-slab of this job's 128 output features
    const uint8_t* x; int64_t x_stride; int x_slab0, xslabs;  // the layer's input tiles
    float* dW; int in_dim, out0;                              // (out, in_dim) fp32, accumulated into; first output feature of the job
    int n_tiles, spt;                                         // tile capacity, samples per tile
    float* db; int bias_col;                                  // bias gradient (out) and the accumulator column that holds it, see below
};
constexpr int WG_MAX_JOBS = 12;
struct WgradP { WgJob j[WG_MAX_JOBS]; int n_jobs; const int* S_dev; };
struct WgSmem {
    uint8_t D[2][16 * SLAB];        // one 128-column half of a delta tile: M = 128 output features, K = 128 rows
    uint8_t X[2][36 * SLAB];        // the layer's input tile: N = up to 288 input features, K = 128 rows
    uint64_t full[2], empty[2], done;
    uint32_t tmem_base;
};

// Bias gradients ride in the same GEMM: db_L = column sums of delta_L = delta_L^T . 1, i.e. the column of dW_L that belongs to a constant-1
// input.  The layer-1 / layer-3 / colour-layer-1 input tiles already carry that column (284 / 263 / 280: the bias column of the forward
// GEMMs, resp. a padding column the forward sets to 1 against a zero weight); for the 256- and 128-wide inputs a persistent ones slab sits
// behind the tile in shared memory and a 16-column tail MMA accumulates it.  (Round 1 ran a separate colsum_kernel over all seven
// gradient tiles: 0.135 of a 1.55 ms training step.)
__global__ void __launch_bounds__(256, 1) wgrad_tc_kernel(const WgradP p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    WgSmem& sm = *reinterpret_cast<WgSmem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const WgJob& jb = p.j[(int)blockIdx.x % p.n_jobs];
    const int part = (int)blockIdx.x / p.n_jobs, parts = (int)gridDim.x / p.n_jobs;
    const int n_tiles = dyn_tiles(p.S_dev, jb.n_tiles, jb.spt);
    const int n_my = n_tiles > part ? (n_tiles - part + parts - 1) / parts : 0;
    const int xslabs = jb.xslabs;
    const bool ones_tail = xslabs <= 32;                       // the input tile has no constant-1 column of its own
    if (tid == 0) {
        for (int s = 0; s < 2; s++) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
        mbar_init(&sm.done, 1);
        fence_barrier_init();
    }
    if (ones_tail) {       // slabs xslabs, xslabs + 1 of both stages: column 0 = 1 for all 128 rows, the other 15 columns 0 (never overwritten: a CTA serves one job)
        for (int i = tid; i < 2 * 2 * ROWS; i += 256) {
            const int s = i / (2 * ROWS), r = i % (2 * ROWS);
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (r < ROWS) v.x = 0x00003F80u;                   // bf16 1.0 in element 0 of the row's 16 bytes
            reinterpret_cast<uint4*>(sm.X[s] + (size_t)xslabs * SLAB)[r] = v;
        }
        fence_proxy_async();
    }
    if (warp == 0) tmem_alloc(&sm.tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    const int tail_slab0 = ones_tail ? xslabs : 32, tail_n = ones_tail ? 16 : 32;
    if (warp == 0) {
        if (lane == 0) {            // producer
            for (int i = 0; i < n_my; i++) {
                const int tile = part + i * parts, s = i & 1;
                if (i >= 2) mbar_wait(&sm.empty[s], (uint32_t)(((i >> 1) - 1) & 1));
                const uint32_t xbytes = (uint32_t)(xslabs * SLAB);
                mbar_arrive_expect_tx(&sm.full[s], 16u * SLAB + xbytes);
                bulk_g2s(sm.D[s], jb.d + (int64_t)tile * jb.d_stride + (int64_t)jb.d_slab0 * SLAB, 16u * SLAB, &sm.full[s]);
                bulk_g2s(sm.X[s], jb.x + (int64_t)tile * jb.x_stride + (int64_t)jb.x_slab0 * SLAB, xbytes, &sm.full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && n_my > 0) {   // MMA issuer: both operands MN-major (k = tile rows): LBO = 128 (next 8 rows), SBO = SLAB (next 8 features)
            const uint32_t mn = (1u << 15) | (1u << 16);
            const uint32_t idesc_main = make_idesc_bf16(ROWS, xslabs >= 32 ? 256 : xslabs * 8) | mn, idesc_tail = make_idesc_bf16(ROWS, tail_n) | mn;
            for (int i = 0; i < n_my; i++) {
                const int s = i & 1;
                mbar_wait(&sm.full[s], (uint32_t)((i >> 1) & 1));
                tc_fence_after();
                const uint32_t d_base = smem_u32(sm.D[s]), x_base = smem_u32(sm.X[s]);
                for (int ks = 0; ks < ROWS / 16; ks++) {
                    const uint64_t ad = make_smem_desc(d_base + (uint32_t)(ks * 256), 128, SLAB);
                    const uint64_t bd = make_smem_desc(x_base + (uint32_t)(ks * 256), 128, SLAB);
                    mma_bf16(tmem, ad, bd, idesc_main, (uint32_t)((i | ks) > 0));
                    const uint64_t bt = make_smem_desc(x_base + (uint32_t)(tail_slab0 * SLAB + ks * 256), 128, SLAB);
                    mma_bf16(tmem + (uint32_t)(tail_slab0 * 8), ad, bt, idesc_tail, (uint32_t)((i | ks) > 0));
                }
                mma_commit(&sm.empty[s]);
            }
            mma_commit(&sm.done);
        }
    } else if (warp >= 4 && n_my > 0) {
        mbar_wait(&sm.done, 0);
        tc_fence_after();
        const int o = jb.out0 + (warp & 3) * 32 + lane;          // output feature = accumulator lane
        const uint32_t tacc = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        float* dst = jb.dW + (int64_t)o * jb.in_dim;
        const int in_dim = jb.in_dim, bias_col = jb.bias_col;
        const int n_cols = tail_slab0 * 8 + tail_n;
#pragma unroll 1
        for (int c0 = 0; c0 < n_cols; c0 += 32) {
            float v[32];
            tmem_ld32(tacc + (uint32_t)c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j++) {
                if (c0 + j < in_dim) { if (jb.dW) atomicAdd(dst + c0 + j, v[j]); }
                else if (c0 + j == bias_col && jb.db) atomicAdd(jb.db + o, v[j]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------- rgb head backward
// rgb = sigmoid(Wc4 c3 + bc4) * 1.002 - 0.001 (SM:358-359): d raw, delta_c3 = lrelu'(c3) * (Wc4^T d raw) per sample; d Wc4, d bc4.
struct HeadBwd {
    const uint8_t* csave; const int* sample_ids; const float *d_rgb, *rgb, *wc4;
    int S, n_tiles; float slope; const int* S_dev;
    uint8_t* d3; float *dwc4, *dbc4;
};
__global__ void __launch_bounds__(128) color_head_bwd_kernel(const HeadBwd p) {
    __shared__ float s_w4[3][HC];
    __shared__ float s_dz[3][ROWS];
    const int row = threadIdx.x;
#pragma unroll
    for (int j = 0; j < 3; j++) s_w4[j][row] = p.wc4[j * HC + row];
    float gw[3] = {0.f, 0.f, 0.f}, gb = 0.f;
    const int S = dyn_count(p.S_dev, p.S), n_tiles = dyn_tiles(p.S_dev, p.n_tiles, ROWS);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();
        const int si = tile * ROWS + row;
        float dz[3] = {0.f, 0.f, 0.f};
        if (si < S) {
            const int slot = p.sample_ids[si];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const float sg = (p.rgb[3 * (int64_t)slot + j] + 0.001f) * (1.f / 1.002f);
                dz[j] = p.d_rgb[3 * (int64_t)slot + j] * 1.002f * sg * (1.f - sg);
            }
        }
        const uint4* c3 = reinterpret_cast<const uint4*>(p.csave + (int64_t)tile * CSAVE_TILE_BYTES + (int64_t)CSAVE_C3 * SLAB + row * 16);
        uint4* out = reinterpret_cast<uint4*>(p.d3 + (int64_t)tile * CDELTA_TILE_BYTES + row * 16);
#pragma unroll 4
        for (int j = 0; j < HC / 8; j++) {
            float h[8], g[8];
            unpack8(c3[j * (SLAB / 16)], h);
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int c = 8 * j + e;
                g[e] = (dz[0] * s_w4[0][c] + dz[1] * s_w4[1][c] + dz[2] * s_w4[2][c]) * (h[e] > 0.f ? 1.f : p.slope);
            }
            out[j * (SLAB / 16)] = pack8(g);
        }
#pragma unroll
        for (int j = 0; j < 3; j++) s_dz[j][row] = dz[j];
        __syncthreads();
        // d Wc4[j][c] += sum_rows dz_j c3[row][c]: thread t owns column t
        const __nv_bfloat16* col = reinterpret_cast<const __nv_bfloat16*>(p.csave + (int64_t)tile * CSAVE_TILE_BYTES + (int64_t)CSAVE_C3 * SLAB) +
                                   (row >> 3) * (SLAB / 2) + (row & 7);
        for (int r = 0; r < ROWS; r++) {
            const float h = __bfloat162float(col[r * 8]);
            gw[0] = fmaf(s_dz[0][r], h, gw[0]); gw[1] = fmaf(s_dz[1][r], h, gw[1]); gw[2] = fmaf(s_dz[2][r], h, gw[2]);
            if (row < 3) gb += s_dz[row][r];
        }
    }
    if (p.dwc4) {
#pragma unroll
        for (int j = 0; j < 3; j++) atomicAdd(p.dwc4 + j * HC + row, gw[j]);
        if (row < 3 && p.dbc4) atomicAdd(p.dbc4 + row, gb);
    }
}

// ---------------------------------------------------------------------------------------------- scatter to the points
struct ScatterP {
    Cam cam;
    const int *sample_pidx, *sample_ids; const float* dirs;
    const float *embed, *conf;
    const float* dx0;               // (rows, 224) fp32: gradient of [feat 32 | PE(feat) 192]
    const uint8_t* d3;              // delta_3 tiles: gradient of mlp_head layer-0 pre-activations
    const float* w3;                // mlp_head.layers.0 weight (256, 263): columns 256..262 multiply the 7 extras
    const float *dw_rows, *save_w;
    int S, SR, K, KP, n_tiles, weight_conf; const int* S_dev;
    float *g_embed, *g_color, *g_dir, *g_conf;
};
// One WARP per neighbour row (round 1: one thread per row, whose 896 B of dx0 and 512 B of delta_3 were read uncoalesced and whose
// 1792-FMA extras product ran serially: 0.19 ms per 136 k rows, 38 long-scoreboard stalls per issue).  Lane d owns embedding dimension d
// (its 7 dx0 values, one coalesced 128-byte red.add per row) and the 8 delta_3 columns of k-slab d (the 56 weights of the extras
// product stay in its registers); the 7 extras gradients are warp-reduced.
// The kernel is latency-bound (a chain of four dependent global loads per row: sample id -> point id -> embedding / gradients): what
// counts is warps in flight, so the 56 extras weights of a lane live in shared memory ([x][e][lane]: conflict-free) instead of
// registers (100 -> ~48 registers, 2 -> 5 blocks per SM).
__global__ void __launch_bounds__(256, 5) scatter_kernel(const ScatterP p) {
    __shared__ float s_w3[7 * 8 * 32];       // mlp_head.layers.0 weight, columns 256..262 (the 7 extras): [x][e][lane] = W3[8 * lane + e][256 + x]
    const int lane = threadIdx.x & 31;
    const int spt = ROWS / p.KP;
    const int S = dyn_count(p.S_dev, p.S), n_tiles = dyn_tiles(p.S_dev, p.n_tiles, spt);
    for (int i = threadIdx.x; i < 7 * 8 * 32; i += 256) {
        const int x = i / 256, e = (i / 32) % 8, l = i % 32;
        s_w3[i] = __ldg(p.w3 + (int64_t)(8 * l + e) * 263 + 256 + x);
    }
    __syncthreads();
    const int64_t n_rows = (int64_t)n_tiles * ROWS;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t grow = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; grow < n_rows; grow += warps) {
        const int tile = (int)(grow / ROWS), row = (int)(grow % ROWS);
        const int si = tile * spt + row / p.KP, k = row % p.KP;
        if (si >= S || k >= p.K) continue;                      // warp-uniform
        const int slot = __ldg(p.sample_ids + si);
        const int pt = __ldg(p.sample_pidx + (int64_t)slot * p.K + k);
        if (pt < 0) continue;
        if (p.g_embed) {
            const float* dx = p.dx0 + grow * NX0;
            float acc = __ldg(dx + lane);
            const float e = __ldg(p.embed + (int64_t)pt * 32 + lane);
            float sn, cs;
            __sincosf(e, &sn, &cs);                           // as the forward encoder: one sincos + double-angle recurrences
            const float2 q0 = __ldg(reinterpret_cast<const float2*>(dx + 32 + 6 * lane));
            const float2 q1 = __ldg(reinterpret_cast<const float2*>(dx + 32 + 6 * lane + 2));
            const float2 q2 = __ldg(reinterpret_cast<const float2*>(dx + 32 + 6 * lane + 4));
            const float ds[3] = {q0.x, q1.x, q2.x}, dc[3] = {q0.y, q1.y, q2.y};
#pragma unroll
            for (int f = 0; f < 3; f++) {                      // d sin(2^f x) = 2^f cos, d cos(2^f x) = -2^f sin (SU:61-67 layout)
                const float sc = (float)(1 << f);
                acc = fmaf(ds[f], sc * cs, acc);
                acc = fmaf(dc[f], -sc * sn, acc);
                const float s2 = 2.f * sn * cs, c2 = 1.f - 2.f * sn * sn;
                sn = s2; cs = c2;
            }
            atomicAdd(p.g_embed + (int64_t)pt * 32 + lane, acc);
        }
        if (p.g_color || p.g_dir) {
            float f[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(p.d3 + (int64_t)tile * DELTA_TILE_BYTES + row * 16) + lane * (SLAB / 16)), f);
            float de[7];
#pragma unroll
            for (int x = 0; x < 7; x++) {
                float a = 0.f;
#pragma unroll
                for (int e = 0; e < 8; e++) a = fmaf(f[e], s_w3[(x * 8 + e) * 32 + lane], a);
#pragma unroll
                for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                de[x] = a;
            }
            if (p.g_color && lane < 3) atomicAdd(p.g_color + 3 * (int64_t)pt + lane, lane == 0 ? de[0] : (lane == 1 ? de[1] : de[2]));
            if (p.g_dir && lane < 3) {
                // dr = dir . Rn ; e[3+j] = dr_j - v_j ; e[6] = <dr, v>  ->  d dr_j = de[3+j] + de[6] v_j ; d dir_i = sum_j d dr_j Rw2c[j][i]
                const int ray = slot / p.SR;
                const float rd[3] = {p.dirs[3 * ray], p.dirs[3 * ray + 1], p.dirs[3 * ray + 2]};
                float v[3];
                rot_w2c(p.cam, rd, v);
                float gi = 0.f;
#pragma unroll
                for (int j = 0; j < 3; j++) gi = fmaf(de[3 + j] + de[6] * v[j], p.cam.Rw[3 * j + lane], gi);
                atomicAdd(p.g_dir + 3 * (int64_t)pt + lane, gi);
            }
        }
        if (p.g_conf && p.weight_conf && lane == 0) {   // w = wn * clamp(conf): d conf = d w * wn (straight-through clamp, PA:740-742)
            const float cc = fminf(fmaxf(p.conf[pt], 1e-4f), 1.f);
            atomicAdd(p.g_conf + pt, p.dw_rows[grow] * (p.save_w[grow] / cc));
        }
    }
}

struct TrainWs {
    uint8_t* F; uint8_t* save; float *save_w, *save_raw; uint8_t* csave; uint8_t* d[4]; uint8_t* dc[3]; float* dF; float* dx0; float* dw_rows;
    uint8_t* wbwd;
    int64_t n_tiles, n_ctiles, total;
};
TrainWs carve_train(void* base, int64_t S, int K) {
    const int KP = K <= 8 ? 8 : (K <= 16 ? 16 : 32);
    // the forward kernel works on CTA pairs: with an odd tile count the peer CTA still encodes (and saves) one all-masked tile
    const int64_t spt = ROWS / KP, n_tiles = ((S + spt - 1) / spt + 1) / 2 * 2, rows = n_tiles * ROWS;
    uint8_t* p = (uint8_t*)base;
    TrainWs w;
    auto take = [&](int64_t bytes) { uint8_t* r = p; p += align_up(bytes, 1024); return r; };
    w.F = take(((S + ROWS - 1) / ROWS + 1) / 2 * 2 * F_TILE_BYTES);
    w.save = take(n_tiles * SAVE_TILE_BYTES);
    w.save_w = (float*)take(rows * 4);
    w.save_raw = (float*)take(rows * 4);
    // colour tiles of 128 samples (the colour kernel works on CTA pairs too: even count)
    const int64_t n_ctiles = ((S + ROWS - 1) / ROWS + 1) / 2 * 2;
    w.n_tiles = n_tiles; w.n_ctiles = n_ctiles;
    w.csave = take(n_ctiles * CSAVE_TILE_BYTES);
    for (int i = 0; i < 3; i++) w.dc[i] = take(n_ctiles * CDELTA_TILE_BYTES);
    w.dF = (float*)take(n_ctiles * ROWS * HID * 4);
    for (int i = 0; i < 4; i++) w.d[i] = take(n_tiles * DELTA_TILE_BYTES);
    w.dx0 = (float*)take(rows * NX0 * 4);
    w.dw_rows = (float*)take(rows * 4);
    w.wbwd = take(WBWD_BYTES);
    w.total = p - (uint8_t*)base;
    return w;
}
}  // namespace
}  // namespace pnerf

using namespace pnerf;

extern "C" int64_t pnerf_field_tc_train_workspace_bytes(int64_t n_samples, int K) { return carve_train(nullptr, n_samples, K).total + 1024; }

extern "C" int pnerf_field_forward_tc_train(const pnerf_points* pts, const pnerf_camera* cam, const pnerf_mlp* mlp, const void* wpack,
                                            const pnerf_mode* mode, const float* dirs, const float* sample_loc, const int* sample_pidx,
                                            const int* sample_ids, int S, const int* n_samples_dev, int SR, int K, float* sigma, float* rgb,
                                            void* workspace, int64_t workspace_bytes, void* stream) {
    if (!pts || !cam || !mlp || !wpack || !mode || S < 0 || K <= 0 || K > 32 || SR <= 0) return PNERF_ERR_ARG;
    if (S == 0) return PNERF_OK;
    if (!workspace || ((uintptr_t)workspace & 255) || workspace_bytes < pnerf_field_tc_train_workspace_bytes(S, K)) return PNERF_ERR_WORKSPACE;
    if (!(mode->lrelu_slope > 0.f && mode->lrelu_slope < 1.f)) return PNERF_ERR_ARG;
    TrainWs w = carve_train(workspace, S, K);
    return field_tc_launch(pts, cam, mlp, wpack, mode, dirs, sample_loc, sample_pidx, sample_ids, S, n_samples_dev, SR, K, sigma, rgb, w.F, w.save,
                           w.save_w, w.save_raw, true, w.csave, (cudaStream_t)stream);
}

extern "C" int pnerf_field_backward_tc(const pnerf_points* pts, const pnerf_camera* cam, const pnerf_mlp* mlp, const pnerf_mode* mode,
                                       const float* dirs, const float* sample_loc, const int* sample_pidx, const int* sample_ids, int S,
                                       const int* S_dev, int SR, int K, const float* d_sigma, const float* d_rgb, const float* rgb,
                                       float* g_embed, float* g_color, float* g_dir, float* g_conf, const pnerf_mlp_grad* gm,
                                       void* workspace, int64_t workspace_bytes, void* points_done_event, void* stream) {
    if (!pts || !cam || !mlp || !mode || !gm || S < 0 || K <= 0 || K > 32 || SR <= 0 || !d_sigma || !d_rgb || !rgb) return PNERF_ERR_ARG;
    if (S == 0) return PNERF_OK;
    if (!workspace || ((uintptr_t)workspace & 255) || workspace_bytes < pnerf_field_tc_train_workspace_bytes(S, K)) return PNERF_ERR_WORKSPACE;
    (void)sample_loc;
    cudaStream_t st = (cudaStream_t)stream;
    TrainWs w = carve_train(workspace, S, K);
    const int KP = K <= 8 ? 8 : (K <= 16 ? 16 : 32);
    const int n_tiles = (int)((S + ROWS / KP - 1) / (ROWS / KP));
    const int n_ctiles = (int)((S + ROWS - 1) / ROWS);
    const float slope = mode->lrelu_slope;
    // transposed bf16 weights: Bt(n = input feature, k = output feature)
    PackTs jobs;
    jobs.j[0] = {mlp->w4, 256, 256, WB4, 256}; jobs.j[1] = {mlp->w3, 263, 256, WB3, 256}; jobs.j[2] = {mlp->w2, 256, 256, WB2, 256};
    jobs.j[3] = {mlp->w1, 284, NX0, WB1, 256};
    jobs.j[4] = {mlp->wc3, 128, 128, WBC3, 128}; jobs.j[5] = {mlp->wc2, 128, 128, WBC2, 128}; jobs.j[6] = {mlp->wc1, 280, 256, WBC1, 128};
    pack_bwd_kernel<<<dim3(64, 7), 256, 0, st>>>(jobs, w.wbwd);
    PNERF_LAUNCH_CHECK();
    const size_t gsm = sizeof(GemmSmem);
    PNERF_CUDA(cudaFuncSetAttribute(tile_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm));
    auto gemm = [&](const uint8_t* in, int64_t in_stride, int ks, const uint8_t* mask, int64_t mask_stride, uint8_t* out_bf,
                    int64_t out_stride, float* out_f32, int ld, int64_t wb, int N, int tiles, int spt) -> int {
        GemmP g;
        g.S_dev = S_dev; g.spt = spt;
        g.in = in; g.in_stride = in_stride; g.ks = ks; g.mask = mask; g.mask_stride = mask_stride;
        g.out_bf = out_bf; g.out_stride = out_stride; g.out_f32 = out_f32; g.ld_f32 = ld;
        g.w = w.wbwd + wb; g.N = N; g.n_tiles = tiles; g.slope = slope;
        tile_gemm_kernel<<<tiles < kSMs ? tiles : kSMs, 384, gsm, st>>>(g);
        PNERF_LAUNCH_CHECK();
        return PNERF_OK;
    };
    int rc;
    // ---- colour network: rgb head -> delta_c3 -> delta_c2 -> delta_c1 -> dF_s
    HeadBwd hb;
    hb.csave = w.csave; hb.sample_ids = sample_ids; hb.d_rgb = d_rgb; hb.rgb = rgb; hb.wc4 = mlp->wc4; hb.S = S; hb.n_tiles = n_ctiles;
    hb.slope = slope; hb.S_dev = S_dev; hb.d3 = w.dc[2]; hb.dwc4 = gm->wc4; hb.dbc4 = gm->bc4;
    color_head_bwd_kernel<<<n_ctiles < kSMs * 8 ? n_ctiles : kSMs * 8, 128, 0, st>>>(hb);
    PNERF_LAUNCH_CHECK();
    if ((rc = gemm(w.dc[2], CDELTA_TILE_BYTES, 16, w.csave + (int64_t)CSAVE_C2 * SLAB, CSAVE_TILE_BYTES, w.dc[1], CDELTA_TILE_BYTES, nullptr, 0,
                   WBC3, 128, n_ctiles, ROWS))) return rc;                                   // delta_c2 = lrelu'(c2) * (delta_c3 Wc3)
    if ((rc = gemm(w.dc[1], CDELTA_TILE_BYTES, 16, w.csave + (int64_t)CSAVE_C1 * SLAB, CSAVE_TILE_BYTES, w.dc[0], CDELTA_TILE_BYTES, nullptr, 0,
                   WBC2, 128, n_ctiles, ROWS))) return rc;                                   // delta_c1 = lrelu'(c1) * (delta_c2 Wc2)
    if ((rc = gemm(w.dc[0], CDELTA_TILE_BYTES, 16, nullptr, 0, nullptr, 0, w.dF, HID, WBC1, 256, n_ctiles, ROWS))) return rc;   // dF_s = delta_c1 Wc1[:, :256]
    // ---- aggregation + density head backward -> delta_4
    AggBwd a;
    a.save = w.save; a.save_w = w.save_w; a.save_raw = w.save_raw; a.sample_ids = sample_ids; a.d_sigma = d_sigma; a.dF = w.dF; a.ldF = HID;
    a.wa = mlp->wa; a.S = S; a.KP = KP; a.n_tiles = n_tiles; a.softplus = mode->density_softplus; a.slope = slope;
    a.S_dev = S_dev; a.d4 = w.d[3]; a.dwa = gm->wa; a.dba = gm->ba; a.dw_rows = w.dw_rows;
    agg_bwd_kernel<<<n_tiles < kSMs * 8 ? n_tiles : kSMs * 8, 128, 0, st>>>(a);
    PNERF_LAUNCH_CHECK();
    // ---- dgrad chain of mlp_head / mlp_base
    auto dgrad = [&](const uint8_t* in, int mask_slab, uint8_t* out_bf, float* out_f32, int64_t wb, int N) -> int {
        return gemm(in, DELTA_TILE_BYTES, 32, mask_slab >= 0 ? w.save + (int64_t)mask_slab * SLAB : nullptr, SAVE_TILE_BYTES, out_bf,
                    DELTA_TILE_BYTES, out_f32, NX0, wb, N, n_tiles, ROWS / KP);
    };
    if ((rc = dgrad(w.d[3], SAVE_H3, w.d[2], nullptr, WB4, 256))) return rc;      // delta_3 = lrelu'(h3) * (delta_4 W4)
    if ((rc = dgrad(w.d[2], SAVE_X3, w.d[1], nullptr, WB3, 256))) return rc;      // delta_2 = lrelu'(h2) * (delta_3 W3[:, :256])
    if ((rc = dgrad(w.d[1], SAVE_H1, w.d[0], nullptr, WB2, 256))) return rc;      // delta_1 = lrelu'(h1) * (delta_2 W2)
    if (g_embed && (rc = dgrad(w.d[0], -1, nullptr, w.dx0, WB1, NX0))) return rc; // d x0[:, :224] = delta_1 W1[:, :224]
    // ---- point gradients first: they are the large collective of a data-parallel step (156 B per point), which can then start
    // (on another stream, after `points_done_event`) under the weight-gradient GEMMs below
    if (g_embed || g_color || g_dir || (g_conf && mode->weight_conf)) {
        ScatterP s;
        s.cam = make_cam(pts, cam);
        s.sample_pidx = sample_pidx; s.sample_ids = sample_ids; s.dirs = dirs; s.embed = pts->embed; s.conf = pts->conf;
        s.dx0 = w.dx0; s.d3 = w.d[2]; s.w3 = mlp->w3; s.dw_rows = w.dw_rows; s.save_w = w.save_w;
        s.S = S; s.S_dev = S_dev; s.SR = SR; s.K = K; s.KP = KP; s.n_tiles = n_tiles; s.weight_conf = mode->weight_conf;
        s.g_embed = g_embed; s.g_color = g_color; s.g_dir = g_dir; s.g_conf = g_conf;
        scatter_kernel<<<n_tiles * 16 < kSMs * 8 ? n_tiles * 16 : kSMs * 8, 256, 0, st>>>(s);
        PNERF_LAUNCH_CHECK();
    }
    if (points_done_event) PNERF_CUDA(cudaEventRecord((cudaEvent_t)points_done_event, st));
    // ---- wgrad + bias gradients: one launch each for all seven layers
    {
        WgradP g;
        g.S_dev = S_dev;
        int nj = 0;
        float* dWf[4] = {gm->w1, gm->w2, gm->w3, gm->w4};
        float* dbf[4] = {gm->b1, gm->b2, gm->b3, gm->b4};
        const int in_dim[4] = {284, 256, 263, 256}, xoff[4] = {SAVE_X0, SAVE_H1, SAVE_X3, SAVE_H3}, xsl[4] = {36, 32, 36, 32};
        const int bcol[4] = {284, 256, 263, 256};              // the constant-1 column of the layer's input (X0 / X3: in the tile; H1 / H3: the ones slab)
        for (int L = 0; L < 4; L++)
            for (int h = 0; h < 2; h++)
                if (dWf[L] || dbf[L]) g.j[nj++] = {w.d[L], DELTA_TILE_BYTES, 16 * h, w.save, SAVE_TILE_BYTES, xoff[L], xsl[L], dWf[L], in_dim[L], 128 * h, n_tiles, ROWS / KP,
                                         dbf[L], bcol[L]};
        float* dWc[3] = {gm->wc1, gm->wc2, gm->wc3};
        float* dbc[3] = {gm->bc1, gm->bc2, gm->bc3};
        const int cin[3] = {280, 128, 128}, coff[3] = {CSAVE_C0, CSAVE_C1, CSAVE_C2}, csl[3] = {36, 16, 16}, cbcol[3] = {280, 128, 128};
        for (int L = 0; L < 3; L++)
            if (dWc[L] || dbc[L]) g.j[nj++] = {w.dc[L], CDELTA_TILE_BYTES, 0, w.csave, CSAVE_TILE_BYTES, coff[L], csl[L], dWc[L], cin[L], 0, n_ctiles, ROWS, dbc[L], cbcol[L]};
        g.n_jobs = nj;
        if (nj > 0) {
            const size_t wsm = sizeof(WgSmem);
            PNERF_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsm));
            wgrad_tc_kernel<<<nj * (kSMs / nj), 256, wsm, st>>>(g);
            PNERF_LAUNCH_CHECK();
        }
    }
    return PNERF_OK;
}
