// Tensor-core implementation of the field networks (rows P, GA, W, E, M1, A of SURVEY.md 8a): gather +
// positional encodings + mlp_base + mlp_head + density head + K-aggregation as ONE persistent kernel on
// tcgen05 / TMEM, and mlp_color + rgb head (row M2) as a second one.  Replaces SU:190-209 and SM:270-366
// (twin PA:486-662,745-830); the reference runs the same math as ~40 separate torch/cuBLAS launches with every
// (M,256) fp32 activation round-tripping HBM.
//
// field_tc_kernel -- one CTA per SM, 448 threads, two tile slots in ping-pong:
//   tile      = 128 rows = (128/KP) consecutive valid samples x KP neighbour slots (KP = 8, 16 or 32 >= K)
//   warps 0-3 : encoder.  Thread = row: gathers the point (xyz, 32-d embedding, colour, dir, conf), computes the
//               relative position in world and perspective space, the inverse-distance weight, the 284-wide encoded
//               input (double-angle recurrences from one sincos per input) and writes it as the bf16 A operand of
//               layer 1 straight into shared memory (K-slab layout, see umma.cuh).  Nothing encoded touches HBM.
//   warps 4-7 / 8-11 : epilogue group of slot 0 / 1.  Thread = row = TMEM lane: tcgen05.ld the fp32 accumulator,
//               bias + LeakyReLU, bf16 pack, write the next layer's A operand in place; after layer 4 the density
//               head (in-thread dot), the weight w_k and the sum over the KP neighbour lanes (register butterfly).
//   warp 12   : weight producer: streams the four 256-wide layers (bf16, pre-packed K-slabs, L2 resident) as
//               16 KB chunks through a 4-stage ring with cp.async.bulk + mbarrier complete_tx.
//   warp 13   : MMA issuer: one thread issues tcgen05.mma (M=128, N=256, K=16) for slot 0 / slot 1 alternately,
//               so one slot's epilogue overlaps the other slot's MMAs; tcgen05.commit releases ring stages
//               and publishes accumulators.
//   TMEM      : 512 columns = 2 slots x (128 lanes x 256 fp32 columns).
//   HBM       : in 168 B per valid row (gather) + indices; out 4 B sigma + 512 B F_s (bf16) per sample.
#include "pnerf_common.cuh"
#include "umma.cuh"

namespace pnerf {
namespace {
using namespace umma;

constexpr int HID = 256;
constexpr int ROWS = 128;
constexpr int SLAB = ROWS * 16;                  // bytes of one 8-wide k-slab of a 128-row operand
constexpr int KIN_PAD = 288;                     // 284 (layer 1) and 263 (layer 3) padded to 9 chunks of 32
constexpr int A_BYTES = (KIN_PAD / 8) * SLAB;    // 73728
constexpr int CHUNK_K = 32;
constexpr int CHUNK_BYTES = CHUNK_K * HID * 2;   // 16384: 4 slabs of a 256-row operand
constexpr int NST = 4;
constexpr int N_CHUNKS = 34;                     // 9 + 8 + 9 + 8
constexpr int NT = 448;
constexpr int MAX_SPT = 16;                      // samples per tile at KP = 8

// colour network
constexpr int HC = 128;
constexpr int C1_BYTES = (KIN_PAD / 8) * HC * 16;   // 73728: Wc1 128 x 288
constexpr int C2_BYTES = (HC / 8) * HC * 16;        // 32768: Wc2 / Wc3 128 x 128
constexpr int WPACK_FIELD_BYTES = N_CHUNKS * CHUNK_BYTES;                  // 557056
constexpr int WPACK_BYTES = WPACK_FIELD_BYTES + C1_BYTES + 2 * C2_BYTES;   // 696320

struct Cam { float o[3]; float Rc[9]; float Rw[9]; };

struct Meta {                       // per (slot, tile parity): written by the encoder, read by the slot's epilogue group
    float w[ROWS];                  // aggregation weight of the row (0 for masked rows)
    uint4 extras[ROWS];             // bf16 x 8: colour 3, dir_r - v 3, <dir_r, v> 1, 0   (layer 3 inputs 256..263)
    int slot_id[MAX_SPT];           // output slot (r*SR+s) of each sample of the tile, -1 past the end
};

struct Smem {
    uint8_t A[2][A_BYTES];
    uint8_t W[NST][CHUNK_BYTES];
    float bias[4][HID];
    float wa[HID];
    Meta meta[2][2];
    uint64_t w_full[NST], w_empty[NST];
    uint64_t a_ready[2], acc_full[2], acc_empty[2], a_free[2];
    uint32_t tmem_base;
};

struct FieldParams {
    const float *xyz, *embed, *color, *dir, *conf;
    const float *dirs, *sample_loc;
    const int *sample_pidx, *sample_ids;
    const uint8_t* wpack;
    const float *b1, *b2, *b3, *b4, *wa, *ba;
    Cam cam;
    int S, SR, K, n_tiles;
    float slope;
    int softplus, weight_conf;
    float* sigma;                   // (R*SR) by slot
    __nv_bfloat16* F;               // (S, 256) aggregated features, by compact sample index
};

__device__ __forceinline__ uint4 pack8(const float* v) {
    uint4 r;
    r.x = pack_bf16(v[0], v[1]); r.y = pack_bf16(v[2], v[3]); r.z = pack_bf16(v[4], v[5]); r.w = pack_bf16(v[6], v[7]);
    return r;
}

__device__ __forceinline__ void to_pers(const Cam& c, float x, float y, float z, float& px, float& py, float& pz) {
    const float sx = x - c.o[0], sy = y - c.o[1], sz = z - c.o[2];
    const float cx = sx * c.Rc[0] + sy * c.Rc[3] + sz * c.Rc[6];
    const float cy = sx * c.Rc[1] + sy * c.Rc[4] + sz * c.Rc[7];
    const float cz = sx * c.Rc[2] + sy * c.Rc[5] + sz * c.Rc[8];
    px = cx / cz; py = cy / cz; pz = cz;
}
__device__ __forceinline__ void rot_w2c(const Cam& c, const float* u, float* out) {
#pragma unroll
    for (int j = 0; j < 3; j++) out[j] = u[0] * c.Rw[3 * j] + u[1] * c.Rw[3 * j + 1] + u[2] * c.Rw[3 * j + 2];
}

// [sin(x 2^f), cos(x 2^f)] for f = 0..F-1 interleaved (SU:61-67), one sincos + double-angle recurrences
template <int F>
__device__ __forceinline__ void pe(float x, float* out) {
    float s, c;
    __sincosf(x, &s, &c);
    out[0] = s; out[1] = c;
#pragma unroll
    for (int f = 1; f < F; f++) {
        const float s2 = 2.f * s * c, c2 = 1.f - 2.f * s * s;
        s = s2; c = c2;
        out[2 * f] = s; out[2 * f + 1] = c;
    }
}

// ---------------------------------------------------------------------------------------------- encoder
template <int KP>
__device__ __forceinline__ void encode_tile(const FieldParams& p, int tile, uint8_t* Abuf, Meta& meta, int row) {
    constexpr int SPT = ROWS / KP;
    const int si = tile * SPT + row / KP;
    const int k = row % KP;
    int slot = -1, pidx = -1;
    if (si < p.S) {
        slot = __ldg(p.sample_ids + si);
        if (k < p.K) pidx = __ldg(p.sample_pidx + (int64_t)slot * p.K + k);
    }
    if (k == 0) meta.slot_id[row / KP] = slot;
    uint4* Arow = reinterpret_cast<uint4*>(Abuf + row * 16);   // slab j of this row = Arow[j * (SLAB/16)]
    constexpr int SJ = SLAB / 16;
    float wraw = 0.f, cc = 1.f;
    float ex[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (pidx >= 0) {
        const float sx = __ldg(p.sample_loc + 3 * (int64_t)slot), sy = __ldg(p.sample_loc + 3 * (int64_t)slot + 1),
                    sz = __ldg(p.sample_loc + 3 * (int64_t)slot + 2);
        const int ray = slot / p.SR;
        const float rd[3] = {__ldg(p.dirs + 3 * (int64_t)ray), __ldg(p.dirs + 3 * (int64_t)ray + 1), __ldg(p.dirs + 3 * (int64_t)ray + 2)};
        const float X = __ldg(p.xyz + 3 * (int64_t)pidx), Y = __ldg(p.xyz + 3 * (int64_t)pidx + 1), Z = __ldg(p.xyz + 3 * (int64_t)pidx + 2);
        const float4* e4 = reinterpret_cast<const float4*>(p.embed + (int64_t)pidx * 32);
        float e[32];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float4 t = __ldg(e4 + j);
            e[4 * j] = t.x; e[4 * j + 1] = t.y; e[4 * j + 2] = t.z; e[4 * j + 3] = t.w;
        }
        const float col[3] = {__ldg(p.color + 3 * (int64_t)pidx), __ldg(p.color + 3 * (int64_t)pidx + 1), __ldg(p.color + 3 * (int64_t)pidx + 2)};
        const float dd[3] = {__ldg(p.dir + 3 * (int64_t)pidx), __ldg(p.dir + 3 * (int64_t)pidx + 1), __ldg(p.dir + 3 * (int64_t)pidx + 2)};
        if (p.weight_conf) cc = fminf(fmaxf(__ldg(p.conf + pidx), 1e-4f), 1.f);                // PA:740-742
        // geometry (SM:273-281)
        float spx, spy, spz, ppx, ppy, ppz;
        to_pers(p.cam, sx, sy, sz, spx, spy, spz);
        to_pers(p.cam, X, Y, Z, ppx, ppy, ppz);
        float d[6];
        d[0] = X - sx; d[1] = Y - sy; d[2] = Z - sz;
        d[3] = ppx * ppz - spx * spz; d[4] = ppy * ppz - spy * spz; d[5] = ppz - spz;
        wraw = 1.f / fmaxf(sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), 1e-6f);              // SM:471-474
        float d3[3];
        rot_w2c(p.cam, d, d3);                                                                  // SM:312
        d[0] = d3[0]; d[1] = d3[1]; d[2] = d3[2];
        float v[3], dr[3];
        rot_w2c(p.cam, rd, v);                                                                  // SM:303-304
        rot_w2c(p.cam, dd, dr);                                                                 // SM:330
        ex[0] = col[0]; ex[1] = col[1]; ex[2] = col[2];
        ex[3] = dr[0] - v[0]; ex[4] = dr[1] - v[1]; ex[5] = dr[2] - v[2];
        ex[6] = dr[0] * v[0] + dr[1] * v[1] + dr[2] * v[2];                                     // SM:334
        // layer-1 input [feat 32 | PE(feat, F=3) 192 | PE(dists6, F=5) 60 | 0 x 4]
#pragma unroll
        for (int j = 0; j < 4; j++) Arow[j * SJ] = pack8(e + 8 * j);
#pragma unroll
        for (int g = 0; g < 8; g++) {            // 4 embedding dims -> 24 values -> slabs 4+3g .. 4+3g+2
            float t[24];
#pragma unroll
            for (int c = 0; c < 4; c++) pe<3>(e[4 * g + c], t + 6 * c);
#pragma unroll
            for (int j = 0; j < 3; j++) Arow[(4 + 3 * g + j) * SJ] = pack8(t + 8 * j);
        }
        {
            float t[64];
#pragma unroll
            for (int c = 0; c < 6; c++) pe<5>(d[c], t + 10 * c);
            t[60] = t[61] = t[62] = t[63] = 0.f;
#pragma unroll
            for (int j = 0; j < 8; j++) Arow[(28 + j) * SJ] = pack8(t + 8 * j);
        }
    } else {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int j = 0; j < KIN_PAD / 8; j++) Arow[j * SJ] = z;
    }
    float wsum = wraw;                            // SM:286: normalise over the sample's neighbours
#pragma unroll
    for (int o = KP / 2; o; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    float w = wraw / fmaxf(wsum, 1e-8f);
    if (p.weight_conf) w *= cc;                   // PA:826 (original flow only)
    meta.w[row] = pidx >= 0 ? w : 0.f;
    meta.extras[row] = pack8(ex);
}

// ---------------------------------------------------------------------------------------------- epilogues
__device__ __forceinline__ void epilogue_store(uint32_t tacc_lane, const float* __restrict__ bias, float slope, uint8_t* Abuf,
                                               int row) {
    uint4* Arow = reinterpret_cast<uint4*>(Abuf + row * 16);
    constexpr int SJ = SLAB / 16;
#pragma unroll 1
    for (int c0 = 0; c0 < HID; c0 += 32) {
        float v[32];
        tmem_ld32(tacc_lane + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(bias + c0 + j);
            float x;
            x = v[j] + b.x; v[j] = fmaxf(x, x * slope);
            x = v[j + 1] + b.y; v[j + 1] = fmaxf(x, x * slope);
            x = v[j + 2] + b.z; v[j + 2] = fmaxf(x, x * slope);
            x = v[j + 3] + b.w; v[j + 3] = fmaxf(x, x * slope);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) Arow[(c0 / 8 + j) * SJ] = pack8(v + 8 * j);
    }
}

// sum over the KP lanes of a neighbour group of 32 per-lane values; afterwards lane gl (position in its group)
// holds, in a[0 .. 32/KP), the sums of values gl*(32/KP) + j.
template <int D, int N>
__device__ __forceinline__ void bfly_step(float* a, int lane) {
    const bool up = (lane & D) != 0;
#pragma unroll
    for (int j = 0; j < N / 2; j++) {
        const float send = up ? a[j] : a[j + N / 2];
        const float keep = up ? a[j + N / 2] : a[j];
        a[j] = keep + __shfl_xor_sync(0xffffffffu, send, D);
    }
}
template <int KP>
__device__ __forceinline__ void butterfly(float* a, int lane) {
    if (KP == 32) { bfly_step<16, 32>(a, lane); bfly_step<8, 16>(a, lane); bfly_step<4, 8>(a, lane); bfly_step<2, 4>(a, lane); bfly_step<1, 2>(a, lane); }
    if (KP == 16) { bfly_step<8, 32>(a, lane); bfly_step<4, 16>(a, lane); bfly_step<2, 8>(a, lane); bfly_step<1, 4>(a, lane); }
    if (KP == 8) { bfly_step<4, 32>(a, lane); bfly_step<2, 16>(a, lane); bfly_step<1, 8>(a, lane); }
}

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }

template <int KP>
__device__ __forceinline__ void epilogue_aggregate(const FieldParams& p, uint32_t tacc_lane, const float* __restrict__ bias,
                                                   const float* __restrict__ wa, const Meta& meta, int tile, int row) {
    constexpr int SPT = ROWS / KP, VPL = 32 / KP;
    const int lane = threadIdx.x & 31, gl = lane % KP;
    const int sl = row / KP;
    const int si = tile * SPT + sl;
    const float w = meta.w[row];
    const int slot = meta.slot_id[sl];
    float dot = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < HID; c0 += 32) {
        float v[32];
        tmem_ld32(tacc_lane + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(bias + c0 + j);
            const float4 a = *reinterpret_cast<const float4*>(wa + c0 + j);
            float x;
            x = v[j] + b.x; x = fmaxf(x, x * p.slope); dot = fmaf(x, a.x, dot); v[j] = x * w;
            x = v[j + 1] + b.y; x = fmaxf(x, x * p.slope); dot = fmaf(x, a.y, dot); v[j + 1] = x * w;
            x = v[j + 2] + b.z; x = fmaxf(x, x * p.slope); dot = fmaf(x, a.z, dot); v[j + 2] = x * w;
            x = v[j + 3] + b.w; x = fmaxf(x, x * p.slope); dot = fmaf(x, a.w, dot); v[j + 3] = x * w;
        }
        butterfly<KP>(v, lane);
        if (slot >= 0) {
            __nv_bfloat16* dst = p.F + (int64_t)si * HID + c0 + gl * VPL;
            if (VPL == 4) {
                uint2 o; o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]);
                *reinterpret_cast<uint2*>(dst) = o;
            } else if (VPL == 2) {
                *reinterpret_cast<uint32_t*>(dst) = pack_bf16(v[0], v[1]);
            } else {
                dst[0] = __float2bfloat16(v[0]);
            }
        }
    }
    const float raw = dot + __ldg(p.ba);
    const float a = p.softplus ? softplus_f(raw - 1.f) : fmaxf(raw, 0.f);   // PA:260-265 / SM:221
    float sg = w * a;
#pragma unroll
    for (int o = KP / 2; o; o >>= 1) sg += __shfl_xor_sync(0xffffffffu, sg, o);
    if (gl == 0 && slot >= 0) p.sigma[slot] = sg;                           // SM:344
}

__host__ __device__ constexpr int layer_chunks(int L) { return (L & 1) ? 8 : 9; }
__host__ __device__ constexpr int layer_chunk0(int L) { return L == 0 ? 0 : (L == 1 ? 9 : (L == 2 ? 17 : 26)); }

template <int KP>
__global__ void __launch_bounds__(NT, 1) field_tc_kernel(const FieldParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_my = p.n_tiles > (int)blockIdx.x ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (tid == 0) {
        for (int i = 0; i < NST; i++) { mbar_init(&sm.w_full[i], 1); mbar_init(&sm.w_empty[i], 1); }
        for (int s = 0; s < 2; s++) {
            mbar_init(&sm.a_ready[s], 128); mbar_init(&sm.acc_full[s], 1); mbar_init(&sm.acc_empty[s], 128); mbar_init(&sm.a_free[s], 1);
        }
        fence_barrier_init();
    }
    if (warp == 13) tmem_alloc(&sm.tmem_base, 512);
    for (int i = tid; i < HID; i += NT) {
        sm.bias[0][i] = p.b1[i]; sm.bias[1][i] = p.b2[i]; sm.bias[2][i] = p.b3[i]; sm.bias[3][i] = p.b4[i];
        sm.wa[i] = p.wa[i];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (warp < 4) {
        // ===================================================== encoder
        uint32_t ph[2] = {0, 0};
        for (int j = 0; j < n_my; j++) {
            const int s = j & 1;
            if (j >= 2) { mbar_wait(&sm.a_free[s], ph[s]); ph[s] ^= 1; }
            encode_tile<KP>(p, (int)blockIdx.x + j * (int)gridDim.x, sm.A[s], sm.meta[s][(j >> 1) & 1], tid);
            fence_proxy_async();
            mbar_arrive(&sm.a_ready[s]);
        }
    } else if (warp < 12) {
        // ===================================================== epilogue group of slot s
        const int s = (warp - 4) >> 2;
        const int row = tid - 128 - s * 128;
        const uint32_t tacc_lane = tmem + (uint32_t)(s * HID) + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t ph = 0;
        for (int j = s; j < n_my; j += 2) {
            const int tile = (int)blockIdx.x + j * (int)gridDim.x;
            Meta& meta = sm.meta[s][(j >> 1) & 1];
#pragma unroll 1
            for (int L = 0; L < 3; L++) {
                mbar_wait(&sm.acc_full[s], ph); ph ^= 1;
                tc_fence_after();
                epilogue_store(tacc_lane, sm.bias[L], p.slope, sm.A[s], row);
                if (L == 1) {   // layer-3 input columns 256..287: the 7 per-row extras, then zeros
                    uint4* Arow = reinterpret_cast<uint4*>(sm.A[s] + row * 16);
                    Arow[32 * (SLAB / 16)] = meta.extras[row];
                    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                    Arow[33 * (SLAB / 16)] = z; Arow[34 * (SLAB / 16)] = z; Arow[35 * (SLAB / 16)] = z;
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(&sm.a_ready[s]);
            }
            mbar_wait(&sm.acc_full[s], ph); ph ^= 1;
            tc_fence_after();
            epilogue_aggregate<KP>(p, tacc_lane, sm.bias[3], sm.wa, meta, tile, row);
            tc_fence_before();
            mbar_arrive(&sm.acc_empty[s]);
        }
    } else if (warp == 12) {
        // ===================================================== weight producer
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int j0 = 0; j0 < n_my; j0 += 2)
                for (int L = 0; L < 4; L++)
                    for (int s = 0; s < 2 && j0 + s < n_my; s++)
                        for (int c = 0; c < layer_chunks(L); c++) {
                            mbar_wait(&sm.w_empty[stage], phase ^ 1);
                            mbar_arrive_expect_tx(&sm.w_full[stage], CHUNK_BYTES);
                            bulk_g2s(sm.W[stage], p.wpack + (size_t)(layer_chunk0(L) + c) * CHUNK_BYTES, CHUNK_BYTES, &sm.w_full[stage]);
                            if (++stage == NST) { stage = 0; phase ^= 1; }
                        }
        }
    } else {
        // ===================================================== MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(ROWS, HID);
            uint32_t stage = 0, phase = 0, ar[2] = {0, 0}, ae[2] = {0, 0};
            for (int j0 = 0; j0 < n_my; j0 += 2)
                for (int L = 0; L < 4; L++)
                    for (int s = 0; s < 2 && j0 + s < n_my; s++) {
                        mbar_wait(&sm.a_ready[s], ar[s]); ar[s] ^= 1;
                        if (L == 0 && j0 >= 2) { mbar_wait(&sm.acc_empty[s], ae[s]); ae[s] ^= 1; }
                        tc_fence_after();
                        const uint32_t tacc = tmem + (uint32_t)(s * HID);
                        const uint32_t a_base = smem_u32(sm.A[s]);
                        for (int c = 0; c < layer_chunks(L); c++) {
                            mbar_wait(&sm.w_full[stage], phase);
                            tc_fence_after();
                            const uint32_t b_base = smem_u32(sm.W[stage]);
#pragma unroll
                            for (int kk = 0; kk < 2; kk++) {
                                const uint64_t ad = make_smem_desc(a_base + (uint32_t)((c * 4 + kk * 2) * SLAB), SLAB, 128);
                                const uint64_t bd = make_smem_desc(b_base + (uint32_t)(kk * 2 * HID * 16), HID * 16, 128);
                                mma_bf16(tacc, ad, bd, idesc, (uint32_t)((c | kk) > 0));
                            }
                            mma_commit(&sm.w_empty[stage]);
                            if (++stage == NST) { stage = 0; phase ^= 1; }
                        }
                        mma_commit(&sm.acc_full[s]);
                        if (L == 3) mma_commit(&sm.a_free[s]);
                    }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 13) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------- colour network
// mlp_color 280 -> 128 -> 128 -> 128 (LeakyReLU) + rgb head 128 -> 3 (sigmoid, *1.002 - 0.001), SM:355-359.
// One CTA = 128 threads = 128 samples per tile; all three weight matrices stay resident in shared memory
// (139 KB), the A operand is [F_s 256 | PE(v) 24 | 0 x 8] built from the bf16 features of field_tc_kernel.
struct ColorParams {
    const __nv_bfloat16* F;
    const int* sample_ids;
    const float* dirs;
    const uint8_t* wpack_c;        // Wc1 | Wc2 | Wc3, K-slab packed
    const float *bc1, *bc2, *bc3, *wc4, *bc4;
    Cam cam;
    int S, SR, n_tiles;
    float slope;
    float* rgb;                    // (R*SR,3) by slot
    __nv_bfloat16* C3;             // optional (S,128) dump of the last hidden layer (training)
};

struct SmemC {
    uint8_t A[A_BYTES];
    uint8_t W1[C1_BYTES];
    uint8_t W2[C2_BYTES];
    uint8_t W3[C2_BYTES];
    float bias[3][HC];
    float w4[3][HC];
    uint64_t bar_w, bar_mma;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(128, 1) color_tc_kernel(const ColorParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    SmemC& sm = *reinterpret_cast<SmemC*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&sm.bar_w, 1); mbar_init(&sm.bar_mma, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&sm.tmem_base, 128);
    for (int i = tid; i < HC; i += 128) {
        sm.bias[0][i] = p.bc1[i]; sm.bias[1][i] = p.bc2[i]; sm.bias[2][i] = p.bc3[i];
        sm.w4[0][i] = p.wc4[i]; sm.w4[1][i] = p.wc4[HC + i]; sm.w4[2][i] = p.wc4[2 * HC + i];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    const uint32_t tacc_lane = tmem + ((uint32_t)(warp * 32) << 16);
    if (tid == 0) {   // weights: one bulk copy each, once per CTA
        mbar_arrive_expect_tx(&sm.bar_w, C1_BYTES + 2 * C2_BYTES);
        bulk_g2s(sm.W1, p.wpack_c, C1_BYTES, &sm.bar_w);
        bulk_g2s(sm.W2, p.wpack_c + C1_BYTES, C2_BYTES, &sm.bar_w);
        bulk_g2s(sm.W3, p.wpack_c + C1_BYTES + C2_BYTES, C2_BYTES, &sm.bar_w);
    }
    const uint32_t idesc = make_idesc_bf16(ROWS, HC);
    uint32_t mma_phase = 0;
    bool w_ready = false;
    uint4* Arow = reinterpret_cast<uint4*>(sm.A + tid * 16);
    constexpr int SJ = SLAB / 16;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int si = tile * ROWS + tid;
        int slot = -1;
        if (si < p.S) {
            slot = __ldg(p.sample_ids + si);
            const uint4* f4 = reinterpret_cast<const uint4*>(p.F + (int64_t)si * HID);
#pragma unroll 8
            for (int j = 0; j < 32; j++) Arow[j * SJ] = __ldg(f4 + j);
            const int ray = slot / p.SR;
            const float rd[3] = {__ldg(p.dirs + 3 * (int64_t)ray), __ldg(p.dirs + 3 * (int64_t)ray + 1), __ldg(p.dirs + 3 * (int64_t)ray + 2)};
            float v[3];
            rot_w2c(p.cam, rd, v);
            float t[32];              // ori=True layout minus the raw copy: [sin (d-major, f-minor) 12 | cos 12] (SM:305-306)
#pragma unroll
            for (int d = 0; d < 3; d++) {
                float q[8];
                pe<4>(v[d], q);
#pragma unroll
                for (int f = 0; f < 4; f++) { t[d * 4 + f] = q[2 * f]; t[12 + d * 4 + f] = q[2 * f + 1]; }
            }
#pragma unroll
            for (int j = 24; j < 32; j++) t[j] = 0.f;
#pragma unroll
            for (int j = 0; j < 4; j++) Arow[(32 + j) * SJ] = pack8(t + 8 * j);
        } else {
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 4
            for (int j = 0; j < 36; j++) Arow[j * SJ] = z;
        }
#pragma unroll 1
        for (int L = 0; L < 3; L++) {
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                if (!w_ready) mbar_wait(&sm.bar_w, 0);
                tc_fence_after();
                const uint32_t a_base = smem_u32(sm.A);
                const uint32_t b_base = smem_u32(L == 0 ? sm.W1 : (L == 1 ? sm.W2 : sm.W3));
                const int nk = L == 0 ? KIN_PAD / 16 : HC / 16;
                for (int ks = 0; ks < nk; ks++) {
                    const uint64_t ad = make_smem_desc(a_base + (uint32_t)(ks * 2 * SLAB), SLAB, 128);
                    const uint64_t bd = make_smem_desc(b_base + (uint32_t)(ks * 2 * HC * 16), HC * 16, 128);
                    mma_bf16(tmem, ad, bd, idesc, (uint32_t)(ks > 0));
                }
                mma_commit(&sm.bar_mma);
            }
            w_ready = true;
            mbar_wait(&sm.bar_mma, mma_phase); mma_phase ^= 1;
            tc_fence_after();
            float r[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
            for (int c0 = 0; c0 < HC; c0 += 32) {
                float v[32];
                tmem_ld32(tacc_lane + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const float x = v[j] + sm.bias[L][c0 + j];
                    v[j] = fmaxf(x, x * p.slope);
                }
                if (L < 2) {
#pragma unroll
                    for (int j = 0; j < 4; j++) Arow[(c0 / 8 + j) * SJ] = pack8(v + 8 * j);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        r[0] = fmaf(v[j], sm.w4[0][c0 + j], r[0]); r[1] = fmaf(v[j], sm.w4[1][c0 + j], r[1]);
                        r[2] = fmaf(v[j], sm.w4[2][c0 + j], r[2]);
                    }
                    if (p.C3 && slot >= 0) {
                        uint4* dst = reinterpret_cast<uint4*>(p.C3 + (int64_t)si * HC + c0);
#pragma unroll
                        for (int j = 0; j < 4; j++) dst[j] = pack8(v + 8 * j);
                    }
                }
            }
            if (L == 2 && slot >= 0) {
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const float x = r[j] + __ldg(p.bc4 + j);
                    p.rgb[3 * (int64_t)slot + j] = (1.f / (1.f + __expf(-x))) * (1.f + 2.f * 0.001f) - 0.001f;   // SM:359
                }
            }
        }
        tc_fence_before();
        __syncthreads();     // the next tile overwrites A and the accumulator
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

// ---------------------------------------------------------------------------------------------- weight packing
// fp32 nn.Linear weights (out,in) -> bf16 K-slab layout [k/8][out][8], K zero-padded; 34 field chunks then Wc1, Wc2, Wc3.
struct PackJob { const float* w; int out, in, kpad; int64_t dst_off; };
struct PackJobs { PackJob j[7]; };

__global__ void __launch_bounds__(256) pack_weights_kernel(PackJobs jobs, uint8_t* __restrict__ dst) {
    const PackJob jb = jobs.j[blockIdx.y];
    const int total = jb.out * jb.kpad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int k = i % jb.kpad, n = i / jb.kpad;
        const float v = k < jb.in ? jb.w[(int64_t)n * jb.in + k] : 0.f;
        __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst + jb.dst_off);
        d[((int64_t)(k >> 3) * jb.out + n) * 8 + (k & 7)] = __float2bfloat16(v);
    }
}

Cam make_cam(const pnerf_points* pts, const pnerf_camera* cam) {
    Cam c;
    for (int i = 0; i < 3; i++) c.o[i] = cam->origin[i];
    for (int i = 0; i < 9; i++) { c.Rc[i] = cam->R_c2w[i]; c.Rw[i] = pts->Rw2c[i]; }
    return c;
}

}  // namespace
}  // namespace pnerf

using namespace pnerf;

extern "C" int64_t pnerf_tc_wpack_bytes(void) { return WPACK_BYTES; }

extern "C" int pnerf_tc_pack_weights(const pnerf_mlp* mlp, void* wpack, void* stream) {
    if (!mlp || !wpack) return PNERF_ERR_ARG;
    PackJobs jobs;
    jobs.j[0] = {mlp->w1, 256, 284, 288, 0};
    jobs.j[1] = {mlp->w2, 256, 256, 256, (int64_t)9 * CHUNK_BYTES};
    jobs.j[2] = {mlp->w3, 256, 263, 288, (int64_t)17 * CHUNK_BYTES};
    jobs.j[3] = {mlp->w4, 256, 256, 256, (int64_t)26 * CHUNK_BYTES};
    jobs.j[4] = {mlp->wc1, 128, 280, 288, (int64_t)WPACK_FIELD_BYTES};
    jobs.j[5] = {mlp->wc2, 128, 128, 128, (int64_t)WPACK_FIELD_BYTES + C1_BYTES};
    jobs.j[6] = {mlp->wc3, 128, 128, 128, (int64_t)WPACK_FIELD_BYTES + C1_BYTES + C2_BYTES};
    for (int i = 0; i < 7; i++) if (!jobs.j[i].w) return PNERF_ERR_ARG;
    pack_weights_kernel<<<dim3(64, 7), 256, 0, (cudaStream_t)stream>>>(jobs, (uint8_t*)wpack);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}

extern "C" int64_t pnerf_field_tc_workspace_bytes(int64_t n_samples) { return align_up(n_samples * HID * 2, 256) + 256; }

extern "C" int pnerf_field_forward_tc(const pnerf_points* pts, const pnerf_camera* cam, const pnerf_mlp* mlp, const void* wpack,
                                      const pnerf_mode* mode, const float* dirs, const float* sample_loc, const int* sample_pidx,
                                      const int* sample_ids, int S, int SR, int K, float* sigma, float* rgb, void* workspace,
                                      int64_t workspace_bytes, void* stream) {
    if (!pts || !cam || !mlp || !wpack || !mode || S < 0 || K <= 0 || K > 32 || SR <= 0) return PNERF_ERR_ARG;
    if (S == 0) return PNERF_OK;
    if (!workspace || workspace_bytes < pnerf_field_tc_workspace_bytes(S)) return PNERF_ERR_WORKSPACE;
    if (!(mode->lrelu_slope > 0.f && mode->lrelu_slope < 1.f)) return PNERF_ERR_ARG;   // lrelu(x) = max(x, slope x)
    cudaStream_t st = (cudaStream_t)stream;
    const int KP = K <= 8 ? 8 : (K <= 16 ? 16 : 32);
    FieldParams p;
    p.xyz = pts->xyz; p.embed = pts->embed; p.color = pts->color; p.dir = pts->dir; p.conf = pts->conf;
    p.dirs = dirs; p.sample_loc = sample_loc; p.sample_pidx = sample_pidx; p.sample_ids = sample_ids;
    p.wpack = (const uint8_t*)wpack;
    p.b1 = mlp->b1; p.b2 = mlp->b2; p.b3 = mlp->b3; p.b4 = mlp->b4; p.wa = mlp->wa; p.ba = mlp->ba;
    p.cam = make_cam(pts, cam);
    p.S = S; p.SR = SR; p.K = K;
    const int spt = ROWS / KP;
    p.n_tiles = (S + spt - 1) / spt;
    p.slope = mode->lrelu_slope; p.softplus = mode->density_softplus; p.weight_conf = mode->weight_conf;
    p.sigma = sigma; p.F = (__nv_bfloat16*)workspace;
    const int grid = p.n_tiles < kSMs ? p.n_tiles : kSMs;
    const size_t smem = sizeof(Smem);
    auto launch = [&](auto kern) -> int {
        PNERF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, NT, smem, st>>>(p);
        PNERF_LAUNCH_CHECK();
        return PNERF_OK;
    };
    int rc = KP == 8 ? launch(field_tc_kernel<8>) : (KP == 16 ? launch(field_tc_kernel<16>) : launch(field_tc_kernel<32>));
    if (rc) return rc;

    ColorParams c;
    c.F = p.F; c.sample_ids = sample_ids; c.dirs = dirs;
    c.wpack_c = (const uint8_t*)wpack + WPACK_FIELD_BYTES;
    c.bc1 = mlp->bc1; c.bc2 = mlp->bc2; c.bc3 = mlp->bc3; c.wc4 = mlp->wc4; c.bc4 = mlp->bc4;
    c.cam = p.cam; c.S = S; c.SR = SR; c.n_tiles = (S + ROWS - 1) / ROWS;
    c.slope = mode->lrelu_slope; c.rgb = rgb; c.C3 = nullptr;
    const int cgrid = c.n_tiles < kSMs ? c.n_tiles : kSMs;
    const size_t csmem = sizeof(SmemC);
    PNERF_CUDA(cudaFuncSetAttribute(color_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
    color_tc_kernel<<<cgrid, 128, csmem, st>>>(c);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}
