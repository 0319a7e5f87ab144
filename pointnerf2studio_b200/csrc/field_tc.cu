// Tensor-core implementation of the field networks (rows P, GA, W, E, M1, A of SURVEY.md 8a): gather +
// positional encodings + mlp_base + mlp_head + density head + K-aggregation as ONE persistent kernel on
// tcgen05 / TMEM, and mlp_color + rgb head (row M2) as a second one.  Replaces SU:190-209 and SM:270-366
// (twin PA:486-662,745-830); the reference runs the same math as ~40 separate torch/cuBLAS launches with every
// (M,256) fp32 activation round-tripping HBM.
//
// field_tc_kernel -- CTA PAIRS (cluster of 2 = the two SMs of a TPC, tcgen05 cta_group::2), one CTA per SM, 704 threads,
// two tile slots in ping-pong.  One MMA (M = 256, N = 256, K = 16) spans the pair: each CTA supplies its own 128 rows
// of A and only HALF of every weight chunk (its 128 of the 256 output features), so the L2 -> SM weight stream is halved per SM.
//   tile      = 128 rows = (128/KP) consecutive valid samples x KP neighbour slots (KP = 8, 16 or 32 >= K)
//   warps 0-3 : encoder (thread = row; ENC_PARTS = 2 splits a row over two threads): neighbour / sample ids are loaded two
//               tiles ahead and the point rows prefetched into L2; gathers the point (xyz, 32-d embedding, colour, dir, conf),
//               computes the relative position in world and perspective space, the inverse-distance weight, the 284-wide
//               encoded input (double-angle recurrences from one sincos per input) plus the constant-1 column that carries
//               the bias, and writes it as the bf16 A operand of layer 1 straight into shared memory (K-slab layout, see
//               umma.cuh).  Nothing encoded touches HBM (except SAVE mode, which keeps every operand tile for the backward).
//   warps 4-11 / 12-19 : epilogue group of slot 0 / 1, two warps per 32 TMEM lanes (one per 128-column half).  Thread = row =
//               TMEM lane: tcgen05.ld the fp32 accumulator (the next 32 columns are requested before the current 32 are
//               processed), LeakyReLU, bf16 pack, write the next layer's A operand in place; after layer 4 the density head
//               (in-thread dot, halves combined through shared memory), the weight w_k and the sum over the KP neighbour
//               lanes (register butterfly); the accumulator is released right after its last load.
//   warp 20   : weight producer: two issuing lanes stream this CTA's half of the four 256-wide layers (bf16, pre-packed
//               K-slabs with the bias column, L2 resident) as 12 KB half-chunks, three groups of two chunks in flight,
//               cp.async.bulk + mbarrier complete_tx.
//   warp 21   : leader CTA: MMA issuer -- one thread issues tcgen05.mma.cta_group::2 for the two slots in the static 8-step
//               period of for_each_step (slots two layers apart: two layers of the other slot lie inside every tile boundary),
//               so one slot's epilogues and its tile boundary overlap the other slot's MMAs; ONE multicast
//               tcgen05.commit per weight group releases the ring stage, one per layer publishes the accumulator in both CTAs.
//               Peer CTA: relay -- forwards "my half of the group has landed" to the leader.  Barriers the issuer waits on live
//               in the leader (remote arrivals from the peer); every wait is an mbarrier.try_wait with a suspend-time hint.
//   TMEM      : 512 columns = 2 slots x (128 lanes x 256 fp32 columns).
//   HBM       : in 168 B per valid row (gather) + indices; out 4 B sigma + 512 B F_s (bf16, in the colour kernel's operand layout:
//               128-sample tiles of 32 k-slabs, tc_layout.cuh) per sample.
// Measured design choices (tools/tc_microbench.py, tools/tc_trace.py, tools/sweep_field_tc.sh) are listed in DESIGN.md section 4.
#include "pnerf_common.cuh"
#include "tc_layout.cuh"
#include "umma.cuh"

namespace pnerf {
namespace {
using namespace umma;
using namespace tcl;

constexpr int HID = 256;
constexpr int KIN_PAD = 288;                     // 284 (layer 1) and 263 (layer 3) padded to 9 chunks of 32
constexpr int A_BYTES = (KIN_PAD / 8) * SLAB;    // 73728
// Weight stream.  A bulk copy costs ~240 clk of engine time per SM whatever its size (measured, tools/tc_microbench.py:
// 4, 8 and 16 KB copies all run at one per 235-245 clk with two issuing lanes, one per ~420 clk with a single lane), so
// the stream is sized in few, large copies: a chunk is 48 k-columns of a layer, stored as two N-halves of 12 KB (one per
// CTA of the pair); the 256-wide layers end with a 16-column chunk.  Chunks travel in groups of two (one per issuing lane),
// a ring stage = a group, released by ONE tcgen05.commit (a commit costs ~120 clk of tensor-pipe time); three groups in flight.
constexpr int CHUNK_SLABS = 6;
constexpr int CHUNK_K = CHUNK_SLABS * 8;
constexpr int CHUNK_BYTES = CHUNK_K * HID * 2;   // 24576 for the pair
constexpr int HALF_BYTES = CHUNK_BYTES / 2;      // 12288: what one CTA loads per full chunk: 6 slabs x 128 output features
constexpr int NGRP = 3;
// Biases ride in the GEMMs: every layer's A operand carries a constant-1 column (x0: pad column 284; h1 / h3: an extra k-slab,
// K = 272; x3: pad column 263) and the packed weights carry the bias in that column, so no epilogue adds a bias (-40 of the
// ~143 instructions per 32 accumulator columns) at the price of one more MMA in the two 256-wide layers.
constexpr int BIAS_COL[4] = {284, 256, 263, 256};
// Warp budget.  Measured (render bench, field kernels): ENC_PARTS 1 / EPW 8 (704 threads, 80 registers): 16.6 ms; ENC_PARTS 2 /
// EPW 8 (832 threads, 72 registers, spills): 17.5 ms although a tile encodes in 4.5 k instead of 7.4 k clk; ENC_PARTS 2 / EPW 4:
// 19.3 ms.  More warps do not shorten the epilogues -- the roles contend for issue slots in bursts -- so instructions, not
// warps, are what the next version has to cut.
#ifndef PNERF_ENC_PARTS
#define PNERF_ENC_PARTS 1
#endif
#ifndef PNERF_EPW
#define PNERF_EPW 8
#endif
constexpr int ENC_PARTS = PNERF_ENC_PARTS;       // threads per row in the encoder: 1 (4 warps) or 2 (8 warps)
constexpr int EPW = PNERF_EPW;                           // epilogue warps per slot: 4 (a warp drains all 256 columns of its 32 lanes) or 8 (128 each)
constexpr int ENCW = 4 * ENC_PARTS;
constexpr int NT = (ENCW + 2 * EPW + 2) * 32;    // encoder + 2 x EPW epilogue + producer + issuer warps
constexpr int MAX_SPT = 64;                      // samples per tile at KP = 2

// colour network
constexpr int HC = 128;
constexpr int C1_BYTES = (KIN_PAD / 8) * HC * 16;   // 73728: Wc1 128 x 288
constexpr int C2_BYTES = (HC / 8) * HC * 16;        // 32768: Wc2 / Wc3 128 x 128
constexpr int WPACK_FIELD_BYTES = (36 + 34 + 34 + 34) * HID * 16;          // 565248
constexpr int WPACK_BYTES = WPACK_FIELD_BYTES + C1_BYTES + 2 * C2_BYTES;   // 688128

struct Cam { float o[3]; float Rc[9]; float Rw[9]; };
struct CamOR { float o[3]; float Rc[9]; };    // the per-step part (pnerf_camera.dev words 0..11)

struct Meta {                       // per (slot, tile parity): written by the encoder, read by the slot's aggregation epilogue
    float w[ROWS];                  // aggregation weight of the row (0 for masked rows)
    int slot_id[MAX_SPT];           // output slot (r*SR+s) of each sample of the tile, -1 past the end
};
struct SlotScratch {                // per slot: lifetimes end before the slot's next tile needs them
    uint4 extras[ROWS];             // encoder -> layer-2 epilogue.  bf16 x 8: colour 3, dir_r - v 3, <dir_r, v> 1, 0 (layer 3 inputs 256..263)
    float dot_hi[ROWS];             // density-head partial dot of columns 128..255 (upper-half epilogue warp -> lower-half warp)
};

struct Smem {
    uint8_t A[2][A_BYTES];
    uint8_t W[NGRP][2][HALF_BYTES];     // three groups in flight, a group = two 48-k chunks (this CTA's N-half)
    Meta meta[2][2];
    SlotScratch scratch[2];
    uint64_t w_full[NGRP], w_empty[NGRP], w_peer[NGRP];
    uint64_t a_ready[2], acc_full[2], acc_empty[2], a_free[2];
    uint32_t tmem_base;
    CamOR cam;                          // camera origin / rotation: the launch parameters, or the device-side step constants (graph replay)
};

static_assert(sizeof(Smem) <= 232448, "field_tc_kernel shared memory exceeds the 227 KB per-CTA limit");

struct FieldParams {
    const float *xyz, *embed, *color, *dir, *conf;
    const float *dirs, *sample_loc;
    const int *sample_pidx, *sample_ids;
    const uint8_t* wpack;
    const float *wa, *ba;                    // density head (the four layer biases ride in the packed weights)
    Cam cam;
    const float* cam_dev;           // pnerf_camera.dev: per-step camera read at run time, or NULL
    int S, SR, K, n_tiles;          // S / n_tiles: host-side CAPACITY when S_dev is set (grid and workspace are sized from it)
    const int* S_dev;               // device-side number of valid samples (<= S), or NULL: the host does not know S and never syncs for it
    int si0;                        // position of this launch's first sample in F (sample lists bucketed by neighbour count)
    float slope;
    int softplus, weight_conf;
    float* sigma;                   // (R*SR) by slot
    uint8_t* F;                     // aggregated features by compact sample index, 128-sample tiles of 32 k-slabs (tc_layout.cuh)
    unsigned long long* trace;      // profiling hook (pnerf_tc_set_trace): per-warp event timelines of CTA 0, or NULL
    // training: every MMA operand of the forward pass is kept for the backward GEMMs, in the tile layout it had in shared
    // memory (so a backward kernel bulk-copies it straight back into an operand buffer):
    //   save + tile * SAVE_TILE_BYTES : [X0 36 slabs | H1 32 | X3 36 | H3 32 | H4 32], slab = 128 rows x 8 bf16 (2 KB)
    uint8_t* save;
    float *save_w, *save_raw;       // per row (tile * 128 + row): aggregation weight, density pre-activation
};


// Pipeline trace: lane 0 of a role warp of CTA 0 appends (clock64 << 8 | event) to its own 2048-entry lane of the buffer.
constexpr int TRACE_PER_WARP = 2048;
struct Tracer {
    unsigned long long* buf; int n;
    __device__ __forceinline__ Tracer(unsigned long long* b, int warp, int lane) : buf(b && blockIdx.x == 0 && lane == 0 ? b + warp * TRACE_PER_WARP : nullptr), n(1) {}
    __device__ __forceinline__ void ev(int id) { if (buf && n < TRACE_PER_WARP) { buf[n++] = ((unsigned long long)clock64() << 8) | (unsigned)id; buf[0] = n; } }
};

__device__ __forceinline__ uint4 pack8(const float* v) {
    uint4 r;
    r.x = pack_bf16(v[0], v[1]); r.y = pack_bf16(v[2], v[3]); r.z = pack_bf16(v[4], v[5]); r.w = pack_bf16(v[6], v[7]);
    return r;
}

template <class C_>
__device__ __forceinline__ void to_pers(const C_& c, float x, float y, float z, float& px, float& py, float& pz) {
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
// The gather is a three-level dependent chain (sample id -> neighbour index -> point row).  The encoder runs it two tiles
// ahead: ids of tile j+2 are loaded and the point rows of tile j+1 are prefetched into L2 while tile j is encoded.
__device__ __forceinline__ int dyn_count(const int* n_dev, int cap) { return n_dev ? min(__ldg(n_dev), cap) : cap; }

template <int KP>
__device__ __forceinline__ void load_ids(const FieldParams& p, int S, int tile, int row, int& slot, int& pidx) {
    constexpr int SPT = ROWS / KP;
    const int si = tile * SPT + row / KP;
    const int k = row % KP;
    slot = -1; pidx = -1;
    if (si < S) {
        slot = __ldg(p.sample_ids + si);
        if (k < p.K) pidx = __ldg(p.sample_pidx + (int64_t)slot * p.K + k);
    }
}
__device__ __forceinline__ void prefetch_l2(const void* a) { asm volatile("prefetch.global.L2 [%0];" ::"l"(a)); }
__device__ __forceinline__ void prefetch_point(const FieldParams& p, int pidx, int part) {
    if (pidx < 0) return;
    if (part == 0) {
        prefetch_l2(p.embed + (int64_t)pidx * 32);
        prefetch_l2(p.xyz + 3 * (int64_t)pidx);
    } else {
        prefetch_l2(p.color + 3 * (int64_t)pidx);
        prefetch_l2(p.dir + 3 * (int64_t)pidx);
    }
}

// A row-owning thread's view of a 128-row operand tile: slab j of the row lives at [j * SLAB + row * 16].  With SAVE the same
// 16 bytes also go to the tile's copy in global memory (a warp stores 512 contiguous bytes per slab).
template <bool SAVE>
struct RowSink {
    uint4* s; uint4* g;
    __device__ __forceinline__ void put(int j, const uint4& v) const {
        s[j * (SLAB / 16)] = v;
        if (SAVE) g[j * (SLAB / 16)] = v;
    }
};

// Two threads per row: PART 0 encodes embedding dims 0..15 (raw -> slabs 0,1; PE -> slabs 4..15) and PE(dists) (slabs 28..35),
// PART 1 encodes embedding dims 16..31 (slabs 2,3 and 16..27), the aggregation weight, the layer-3 extras and the tile metadata.
// Both recompute the (cheap) geometry.  The encoder sits on the critical path at every tile boundary -- a slot's A buffer is only
// free once its layer-4 MMAs are done -- so its latency, not its throughput, is what the split buys.
template <int KP, bool SAVE, int PART>
__device__ __forceinline__ void encode_tile(const FieldParams& p, const CamOR& camor, int slot, int pidx, uint8_t* Abuf, uint8_t* gsave, Meta& meta,
                                            SlotScratch& scr, int row) {
    const int k = row % KP;
    if (PART == 1 && k == 0) meta.slot_id[row / KP] = slot;
    const RowSink<SAVE> A{reinterpret_cast<uint4*>(Abuf + row * 16), reinterpret_cast<uint4*>(gsave + row * 16)};
    float wraw = 0.f, cc = 1.f;
    float ex[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (pidx >= 0) {
        const float sx = __ldg(p.sample_loc + 3 * (int64_t)slot), sy = __ldg(p.sample_loc + 3 * (int64_t)slot + 1),
                    sz = __ldg(p.sample_loc + 3 * (int64_t)slot + 2);
        const float X = __ldg(p.xyz + 3 * (int64_t)pidx), Y = __ldg(p.xyz + 3 * (int64_t)pidx + 1), Z = __ldg(p.xyz + 3 * (int64_t)pidx + 2);
        const float4* e4 = reinterpret_cast<const float4*>(p.embed + (int64_t)pidx * 32 + 16 * PART);
        float e[16];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float4 t = __ldg(e4 + j);
            e[4 * j] = t.x; e[4 * j + 1] = t.y; e[4 * j + 2] = t.z; e[4 * j + 3] = t.w;
        }
        float d[6];
        d[0] = X - sx; d[1] = Y - sy; d[2] = Z - sz;
        // this part's half of [feat 32 | PE(feat, F=3) 192]
#pragma unroll
        for (int j = 0; j < 2; j++) A.put(2 * PART + j, pack8(e + 8 * j));
#pragma unroll
        for (int g = 0; g < 4; g++) {            // 4 embedding dims -> 24 values -> 3 slabs
            float t[24];
#pragma unroll
            for (int c = 0; c < 4; c++) pe<3>(e[4 * g + c], t + 6 * c);
#pragma unroll
            for (int j = 0; j < 3; j++) A.put(4 + 12 * PART + 3 * g + j, pack8(t + 8 * j));
        }
        if (PART == 0) {
            // geometry (SM:273-281) and PE(dists6, F=5) 60 | 0 x 4
            float spx, spy, spz, ppx, ppy, ppz;
            if (p.cam_dev) {      // graph replay: this step's camera from shared memory (warp-uniform branch)
                to_pers(camor, sx, sy, sz, spx, spy, spz);
                to_pers(camor, X, Y, Z, ppx, ppy, ppz);
            } else {              // launch parameters: constant-bank operands, no loads
                to_pers(p.cam, sx, sy, sz, spx, spy, spz);
                to_pers(p.cam, X, Y, Z, ppx, ppy, ppz);
            }
            d[3] = ppx * ppz - spx * spz; d[4] = ppy * ppz - spy * spz; d[5] = ppz - spz;
            float d3[3];
            rot_w2c(p.cam, d, d3);                                                              // SM:312
            d[0] = d3[0]; d[1] = d3[1]; d[2] = d3[2];
            float t[64];
#pragma unroll
            for (int c = 0; c < 6; c++) pe<5>(d[c], t + 10 * c);
            t[60] = 1.f;                           // column 284: the constant that carries mlp_base.layers.0's bias
            t[61] = t[62] = t[63] = 0.f;
#pragma unroll
            for (int j = 0; j < 8; j++) A.put(28 + j, pack8(t + 8 * j));
        } else {
            wraw = 1.f / fmaxf(sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), 1e-6f);          // SM:471-474
            const int ray = slot / p.SR;
            const float rd[3] = {__ldg(p.dirs + 3 * (int64_t)ray), __ldg(p.dirs + 3 * (int64_t)ray + 1), __ldg(p.dirs + 3 * (int64_t)ray + 2)};
            const float col[3] = {__ldg(p.color + 3 * (int64_t)pidx), __ldg(p.color + 3 * (int64_t)pidx + 1), __ldg(p.color + 3 * (int64_t)pidx + 2)};
            const float dd[3] = {__ldg(p.dir + 3 * (int64_t)pidx), __ldg(p.dir + 3 * (int64_t)pidx + 1), __ldg(p.dir + 3 * (int64_t)pidx + 2)};
            if (p.weight_conf) cc = fminf(fmaxf(__ldg(p.conf + pidx), 1e-4f), 1.f);            // PA:740-742
            float v[3], dr[3];
            rot_w2c(p.cam, rd, v);                                                              // SM:303-304
            rot_w2c(p.cam, dd, dr);                                                             // SM:330
            ex[0] = col[0]; ex[1] = col[1]; ex[2] = col[2];
            ex[3] = dr[0] - v[0]; ex[4] = dr[1] - v[1]; ex[5] = dr[2] - v[2];
            ex[6] = dr[0] * v[0] + dr[1] * v[1] + dr[2] * v[2];                                 // SM:334
            ex[7] = 1.f;                                                                        // column 263: carries mlp_head.layers.0's bias
        }
    } else {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int j = 0; j < 2; j++) A.put(2 * PART + j, z);
#pragma unroll
        for (int j = 0; j < 12; j++) A.put(4 + 12 * PART + j, z);
        if (PART == 0) {
#pragma unroll
            for (int j = 28; j < KIN_PAD / 8; j++) A.put(j, z);
        }
    }
    if (PART == 1) {
        float wsum = wraw;                        // SM:286: normalise over the sample's neighbours
#pragma unroll
        for (int o = KP / 2; o; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
        float w = wraw / fmaxf(wsum, 1e-8f);
        if (p.weight_conf) w *= cc;               // PA:826 (original flow only)
        meta.w[row] = pidx >= 0 ? w : 0.f;
        scr.extras[row] = pack8(ex);
    }
}

// ---------------------------------------------------------------------------------------------- epilogues
// The kernel is co-bound by warp-instruction issue (ncu: 29 k warp-instructions per tile against 11 k tensor-pipe clocks), and
// 40 % of those were the FMUL / FMNMX of LeakyReLU.  Blackwell's packed fp32 arithmetic (mul / fma / add .f32x2: two IEEE fp32
// results per instruction, bit-identical to the scalar forms) halves the multiplies and adds of the epilogues.
#ifndef PNERF_F32X2
#define PNERF_F32X2 1
#endif
__device__ __forceinline__ uint64_t pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(uint64_t r, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ void mul2(float& a, float& b, float s) {            // (a, b) *= s
#if PNERF_F32X2
    uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack2(a, b)), "l"(pack2(s, s))); unpack2(r, a, b);
#else
    a *= s; b *= s;
#endif
}
__device__ __forceinline__ void add2(float& a, float& b, float c, float d) {   // (a, b) += (c, d)
#if PNERF_F32X2
    uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack2(a, b)), "l"(pack2(c, d))); unpack2(r, a, b);
#else
    a += c; b += d;
#endif
}
__device__ __forceinline__ void fma2(float& a, float& b, float x, float y, float u, float v) {   // (a, b) += (x * u, y * v)
#if PNERF_F32X2
    uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pack2(x, y)), "l"(pack2(u, v)), "l"(pack2(a, b))); unpack2(r, a, b);
#else
    a = fmaf(x, u, a); b = fmaf(y, v, b);
#endif
}
__device__ __forceinline__ void lrelu2(float& a, float& b, float slope) {      // LeakyReLU = max(x, slope * x), 0 < slope < 1
    float sa = a, sb = b;
    mul2(sa, sb, slope);
    a = fmaxf(a, sa); b = fmaxf(b, sb);
}
// Hidden-layer epilogue of one 128-column half: bias + LeakyReLU in fp32, bf16 pack, next layer's A operand in place.
// tcgen05.ld takes ~210 clk while the tensor pipe works on the other slot (60 clk idle; tools/tc_microbench.py), so the load
// of chunk i+1 is issued before chunk i is processed (tcgen05.wait::ld waits for every outstanding load, hence the order).
template <bool SAVE>
__device__ __forceinline__ void epilogue_chunk_store(float (&v)[32], float slope, const RowSink<SAVE>& A, int c0) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) lrelu2(v[j], v[j + 1], slope);       // the bias is already in the accumulator
#pragma unroll
    for (int j = 0; j < 4; j++) A.put(c0 / 8 + j, pack8(v + 8 * j));
}
template <bool SAVE>
__device__ __forceinline__ void epilogue_store(uint32_t tacc_lane, float slope, uint8_t* Abuf, uint8_t* gsave, int row, int cbeg) {
    const RowSink<SAVE> Arow{reinterpret_cast<uint4*>(Abuf + row * 16), reinterpret_cast<uint4*>(gsave + row * 16)};
    float va[32], vb[32];
    tmem_ld32(tacc_lane + cbeg, va);
    tmem_ld_wait();
    tmem_ld32(tacc_lane + cbeg + 32, vb);
    epilogue_chunk_store<SAVE>(va, slope, Arow, cbeg);
    tmem_ld_wait();
    tmem_ld32(tacc_lane + cbeg + 64, va);
    epilogue_chunk_store<SAVE>(vb, slope, Arow, cbeg + 32);
    tmem_ld_wait();
    tmem_ld32(tacc_lane + cbeg + 96, vb);
    epilogue_chunk_store<SAVE>(va, slope, Arow, cbeg + 64);
    tmem_ld_wait();
    epilogue_chunk_store<SAVE>(vb, slope, Arow, cbeg + 96);
}

// sum over the KP lanes of a neighbour group of 32 per-lane values; afterwards lane gl (position in its group)
// holds, in a[0 .. 32/KP), the sums of values gl*(32/KP) + j.
template <int D, int N>
__device__ __forceinline__ void bfly_step(float* a, int lane) {
    const bool up = (lane & D) != 0;
    if constexpr (N == 2) {
        const float send = up ? a[0] : a[1], keep = up ? a[1] : a[0];
        a[0] = keep + __shfl_xor_sync(0xffffffffu, send, D);
    } else {
#pragma unroll
        for (int j = 0; j + 1 < N / 2; j += 2) {
            const float s0 = up ? a[j] : a[j + N / 2], s1 = up ? a[j + 1] : a[j + 1 + N / 2];
            float k0 = up ? a[j + N / 2] : a[j], k1 = up ? a[j + 1 + N / 2] : a[j + 1];
            add2(k0, k1, __shfl_xor_sync(0xffffffffu, s0, D), __shfl_xor_sync(0xffffffffu, s1, D));
            a[j] = k0; a[j + 1] = k1;
        }
    }
}
template <int KP>
__device__ __forceinline__ void butterfly(float* a, int lane) {
    if (KP == 32) { bfly_step<16, 32>(a, lane); bfly_step<8, 16>(a, lane); bfly_step<4, 8>(a, lane); bfly_step<2, 4>(a, lane); bfly_step<1, 2>(a, lane); }
    if (KP == 16) { bfly_step<8, 32>(a, lane); bfly_step<4, 16>(a, lane); bfly_step<2, 8>(a, lane); bfly_step<1, 4>(a, lane); }
    if (KP == 8) { bfly_step<4, 32>(a, lane); bfly_step<2, 16>(a, lane); bfly_step<1, 8>(a, lane); }
    if (KP == 4) { bfly_step<2, 32>(a, lane); bfly_step<1, 16>(a, lane); }
    if (KP == 2) { bfly_step<1, 32>(a, lane); }
}

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }

template <int KP, bool SAVE>
__device__ __forceinline__ void aggregate_chunk(const FieldParams& p, float (&v)[32], const float* __restrict__ wa, float w, float& dot, float& dot1,
                                                int slot, int si, int c0, int lane, uint4* gh4) {
    constexpr int VPL = 32 / KP;
    const int gl = lane % KP;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(wa + c0 + j));
        lrelu2(v[j], v[j + 1], p.slope);
        lrelu2(v[j + 2], v[j + 3], p.slope);
        fma2(dot, dot1, v[j], v[j + 1], a.x, a.y);
        fma2(dot, dot1, v[j + 2], v[j + 3], a.z, a.w);
    }
    if (SAVE) {
#pragma unroll
        for (int j = 0; j < 4; j++) gh4[(c0 / 8 + j) * (SLAB / 16)] = pack8(v + 8 * j);
    }
#pragma unroll
    for (int j = 0; j < 32; j += 2) mul2(v[j], v[j + 1], w);
    butterfly<KP>(v, lane);
    if (slot >= 0) {
        const int c = c0 + gl * VPL;
        uint8_t* dst = p.F + (int64_t)(si / ROWS) * F_TILE_BYTES + (c >> 3) * SLAB + (si % ROWS) * 16 + (c & 7) * 2;
        if (VPL >= 8) {           // KP = 4 / 2: 8 / 16 columns per lane = one / two whole k-slab entries of the sample's row
#pragma unroll
            for (int q = 0; q < VPL / 8; q++) *reinterpret_cast<uint4*>(dst + q * SLAB) = pack8(v + 8 * q);
        } else if (VPL == 4) {
            uint2 o; o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]);
            *reinterpret_cast<uint2*>(dst) = o;
        } else if (VPL == 2) {
            *reinterpret_cast<uint32_t*>(dst) = pack_bf16(v[0], v[1]);
        } else {
            *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16(v[0]);
        }
    }
}

template <int KP, bool SAVE>
__device__ __forceinline__ void epilogue_aggregate(const FieldParams& p, uint32_t tacc_lane, const float* __restrict__ wa, Meta& meta,
                                                   SlotScratch& scr, int tile, int row, int half, int bar_id, uint64_t* acc_empty) {
    uint4* gh4 = SAVE ? reinterpret_cast<uint4*>(p.save + (int64_t)tile * SAVE_TILE_BYTES + (int64_t)SAVE_H4 * SLAB + row * 16) : nullptr;
    constexpr int SPT = ROWS / KP;
    const int lane = threadIdx.x & 31, gl = lane % KP;
    const int sl = row / KP;
    const int si = p.si0 + tile * SPT + sl;
    const float w = meta.w[row];
    const int slot = meta.slot_id[sl];
    float dot = 0.f, dot1 = 0.f;     // even / odd columns of the density head
#pragma unroll 1
    for (int cbeg = (EPW == 8 ? half * (HID / 2) : 0); cbeg < (EPW == 8 ? (half + 1) * (HID / 2) : HID); cbeg += HID / 2) {
        float va[32], vb[32];
        tmem_ld32(tacc_lane + cbeg, va);
        tmem_ld_wait();
        tmem_ld32(tacc_lane + cbeg + 32, vb);
        aggregate_chunk<KP, SAVE>(p, va, wa, w, dot, dot1, slot, si, cbeg, lane, gh4);
        tmem_ld_wait();
        tmem_ld32(tacc_lane + cbeg + 64, va);
        aggregate_chunk<KP, SAVE>(p, vb, wa, w, dot, dot1, slot, si, cbeg + 32, lane, gh4);
        tmem_ld_wait();
        tmem_ld32(tacc_lane + cbeg + 96, vb);
        aggregate_chunk<KP, SAVE>(p, va, wa, w, dot, dot1, slot, si, cbeg + 64, lane, gh4);
        tmem_ld_wait();
        if (cbeg + HID / 2 >= (EPW == 8 ? (half + 1) * (HID / 2) : HID)) {
            // this warp's last TMEM load has landed: the accumulator may be overwritten by the slot's next tile while the last
            // chunk is still being reduced and stored (the aggregation epilogue is what gates a tile boundary)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(acc_empty, 0);
        }
        aggregate_chunk<KP, SAVE>(p, vb, wa, w, dot, dot1, slot, si, cbeg + 96, lane, gh4);
    }
    dot += dot1;
    // combine the two column halves of the density head: the upper-half warp hands its partial dot to the lower-half warp
    float dot_other = 0.f;
    if (EPW == 8) {
        if (half) scr.dot_hi[row] = dot;
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        if (half) return;
        dot_other = scr.dot_hi[row];
    }
    const float raw = dot + dot_other + __ldg(p.ba);
    if (SAVE) { p.save_raw[(int64_t)tile * ROWS + row] = raw; p.save_w[(int64_t)tile * ROWS + row] = w; }
    const float a = p.softplus ? softplus_f(raw - 1.f) : fmaxf(raw, 0.f);   // PA:260-265 / SM:221
    float sg = w * a;
#pragma unroll
    for (int o = KP / 2; o; o >>= 1) sg += __shfl_xor_sync(0xffffffffu, sg, o);
    if (gl == 0 && slot >= 0) p.sigma[slot] = sg;                           // SM:344
}

// 8-wide k-slabs of the layer: K = 288 for layer 1 (284 inputs + the bias column), K = 272 for the others -- layer 3's 263 inputs + its bias
// column (264) fit 272 as well: one MMA less than the 288 it used to be padded to (1 of 70 per tile)
__host__ __device__ constexpr int layer_slabs(int L) { return L == 0 ? 36 : 34; }
__host__ __device__ constexpr int layer_chunks(int L) { return (void)L, 6; }            // 48-k chunks: 288 = 6 x 48, 272 = 5 x 48 + 32
__host__ __device__ constexpr int layer_byte0(int L) { return HID * 16 * (L == 0 ? 0 : (L == 1 ? 36 : (L == 2 ? 70 : 104))); }
__host__ __device__ constexpr int chunk_slabs(int L, int c) {
    return layer_slabs(L) - CHUNK_SLABS * c < CHUNK_SLABS ? layer_slabs(L) - CHUNK_SLABS * c : CHUNK_SLABS;
}

// The order in which (slot, layer) steps go through the tensor pipe, shared by the weight producer, the MMA issuer and the
// relay (the issue order is static because the weight stream has to be prefetched in that order).  Slot s works on this CTA's
// tiles j = 2*it + s; step q of a slot is layer q & 3 of its tile q >> 2.
// What the order has to hide (pipeline trace, clocks): a layer's MMAs T = 2.85 k, a hidden-layer epilogue E = 1.3 k, a tile
// boundary B = 6.5 k (aggregation epilogue drains the accumulator while the ONE encoder group writes the slot's next tile).
// Strict alternation s0 s1 s0 s1 ... puts one layer of the other slot between a slot's L3 and its next L0 and makes both
// boundaries coincide: the pipe idles B - T per boundary and the second encode queues behind the first (7.9 k of 33.2 k clk per
// pair of tiles; field kernels of the render bench 13.6 ms).  Instead the slots run two layers apart in the period
//      s0L0 s1L2 s0L1 s1L3 s0L2 s0L3 s1L0' s1L1'
// so that TWO layers of the other slot (+ its epilogue gap) lie inside every boundary and the two encodes never overlap; the
// price is one exposed epilogue E per slot and period (s0L2->s0L3, s1L0->s1L1).
template <class F>
__device__ __forceinline__ void for_each_step(int n_my, F&& fn) {
    const int n0 = 4 * ((n_my + 1) >> 1), n1 = 4 * (n_my >> 1);
    auto go = [&](int s, int q) { if (q >= 0 && q < (s ? n1 : n0)) fn(s, q & 3, q >> 2); };
    for (int k = 0; 4 * k < n0 + 4; k++) {
        go(0, 4 * k);
        go(1, 4 * k - 2);
        go(0, 4 * k + 1);
        go(1, 4 * k - 1);
        go(0, 4 * k + 2);
        go(0, 4 * k + 3);
        go(1, 4 * k);
        go(1, 4 * k + 1);
    }
}

// The MMA issuer is one thread on the critical path of every hand-off: it polls (mbarrier.test_wait) instead of parking in try_wait
// with a suspend hint like every other role (-1 % on the field kernels, A/B on one box: 12.08 -> 11.97 ms).
#ifndef PNERF_ISSUER_SPIN
#define PNERF_ISSUER_SPIN 1
#endif
__device__ __forceinline__ void mbar_wait_poll(uint64_t* bar, uint32_t parity) {
    for (uint32_t spins = 0;; ++spins) {
        uint32_t ok;
        asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
        if (spins > (1u << 28)) __trap();
    }
}
#if PNERF_ISSUER_SPIN
#define ISSUER_WAIT(bar, parity) mbar_wait_poll(bar, parity)
#else
#define ISSUER_WAIT(bar, parity) mbar_wait_cluster(bar, parity)
#endif
template <int KP, bool SAVE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT, 1) field_tc_kernel(const FieldParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();                      // 0 = leader (issues the MMAs)
    const int pair = (int)blockIdx.x >> 1, n_pairs = (int)gridDim.x >> 1;
    const int S = dyn_count(p.S_dev, p.S);
    const int n_tiles = (S + ROWS / KP - 1) / (ROWS / KP);
    const int n_super = (n_tiles + 1) >> 1;                       // 256-row super-tiles; this CTA takes tile 2*T + rank
    const int n_my = n_super > pair ? (n_super - pair + n_pairs - 1) / n_pairs : 0;
    auto tile_of = [&](int j) { return 2 * (pair + j * n_pairs) + (int)rank; };

    if (tid == 0) {
        for (int i = 0; i < NGRP; i++) { mbar_init(&sm.w_full[i], 1); mbar_init(&sm.w_empty[i], 1); mbar_init(&sm.w_peer[i], 1); }
        for (int s = 0; s < 2; s++) {
            // a_ready / acc_empty: one arrival per warp of the pair: 2 CTAs x 8 encoder warps, 2 CTAs x 8 epilogue warps
            mbar_init(&sm.a_ready[s], 16); mbar_init(&sm.acc_full[s], 1); mbar_init(&sm.acc_empty[s], 2 * EPW); mbar_init(&sm.a_free[s], 1);
        }
        fence_barrier_init();
    }
    if (warp == ENCW + 2 * EPW + 1) tmem_alloc2(&sm.tmem_base, 512);
    if (tid < 12) {
        float* dst = tid < 3 ? &sm.cam.o[tid] : &sm.cam.Rc[tid - 3];
        *dst = p.cam_dev ? __ldg(p.cam_dev + tid) : (tid < 3 ? p.cam.o[tid] : p.cam.Rc[tid - 3]);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // both CTAs' barriers exist before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    Tracer tr(p.trace, warp, lane);
    if (warp < ENCW) {
        // ===================================================== encoder: ENC_PARTS threads per row (warps 0-3 part 0, warps 4-7 part 1)
        const int part = warp >> 2, erow = tid & 127;
        uint32_t ph[2] = {0, 0};
        int slot0, pidx0, slot1 = -1, pidx1 = -1;
        load_ids<KP>(p, S, tile_of(0), erow, slot0, pidx0);
        if (n_my > 1) load_ids<KP>(p, S, tile_of(1), erow, slot1, pidx1);
        for (int j = 0; j < n_my; j++) {
            const int s = j & 1;
            int slot2 = -1, pidx2 = -1;
            if (j + 2 < n_my) load_ids<KP>(p, S, tile_of(j + 2), erow, slot2, pidx2);
            if (ENC_PARTS == 1) { prefetch_point(p, pidx1, 0); prefetch_point(p, pidx1, 1); }
            else prefetch_point(p, pidx1, part);
            if (j >= 2) { mbar_wait(&sm.a_free[s], ph[s]); ph[s] ^= 1; }
            tr.ev(1);
            uint8_t* gsave = SAVE ? p.save + (int64_t)tile_of(j) * SAVE_TILE_BYTES : nullptr;
            if (ENC_PARTS == 1 || part == 0) encode_tile<KP, SAVE, 0>(p, sm.cam, slot0, pidx0, sm.A[s], gsave, sm.meta[s][(j >> 1) & 1], sm.scratch[s], erow);
            if (ENC_PARTS == 1 || part == 1) encode_tile<KP, SAVE, 1>(p, sm.cam, slot0, pidx0, sm.A[s], gsave, sm.meta[s][(j >> 1) & 1], sm.scratch[s], erow);
            slot0 = slot1; pidx0 = pidx1; slot1 = slot2; pidx1 = pidx2;
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                for (int i = 0; i < 4 / ENCW + 1; i++) mbar_arrive_remote(&sm.a_ready[s], 0);   // the barrier counts 16 per pair: 2 per warp at ENCW = 4
            }
            tr.ev(2);
        }
    } else if (warp < ENCW + 2 * EPW) {
        // ===================================================== epilogue group of slot s: warp = lane quarter (warp & 3, the
        // TMEM lanes a warp may touch) [x column half when EPW == 8]
        const int s = (warp - ENCW) / EPW;
        const int half = EPW == 8 ? ((warp - ENCW) >> 2) & 1 : 0;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t tacc_lane = tmem + (uint32_t)(s * HID) + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t ph = 0;
        for (int j = s; j < n_my; j += 2) {
            const int tile = tile_of(j);
            Meta& meta = sm.meta[s][(j >> 1) & 1];
#pragma unroll 1
            for (int L = 0; L < 3; L++) {
                mbar_wait(&sm.acc_full[s], ph); ph ^= 1;
                tc_fence_after();
                tr.ev(10 + L);
                // layer L's output is the next layer's A operand: H1 (L=0), X3 = [H2 | extras] (L=1), H3 (L=2)
                uint8_t* gsave = SAVE ? p.save + (int64_t)tile * SAVE_TILE_BYTES + (int64_t)(L == 0 ? SAVE_H1 : (L == 1 ? SAVE_X3 : SAVE_H3)) * SLAB
                                      : nullptr;
                if (EPW == 8) {
                    epilogue_store<SAVE>(tacc_lane, p.slope, sm.A[s], gsave, row, half * (HID / 2));
                } else {
                    epilogue_store<SAVE>(tacc_lane, p.slope, sm.A[s], gsave, row, 0);
                    epilogue_store<SAVE>(tacc_lane, p.slope, sm.A[s], gsave, row, HID / 2);
                }
                if (half || EPW == 4) {
                    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                    if (L == 1) {   // layer-3 input columns 256..287: the 7 per-row extras and the constant 1 (bias column 263), then zeros
                        const RowSink<SAVE> A{reinterpret_cast<uint4*>(sm.A[s] + row * 16), reinterpret_cast<uint4*>(gsave + row * 16)};
                        A.put(32, sm.scratch[s].extras[row]);
                        A.put(33, z);                               // K = 272: columns 264..271
                        if (SAVE) { A.put(34, z); A.put(35, z); }   // the saved X3 tile keeps its 36-slab layout for the backward GEMMs
                    } else {        // layers 2 and 4 (K = 272): column 256 = 1 carries the bias, 257..271 = 0 (not part of the saved tile)
                        uint4* Arow = reinterpret_cast<uint4*>(sm.A[s] + row * 16);
                        Arow[32 * (SLAB / 16)] = make_uint4(0x00003F80u, 0u, 0u, 0u);
                        Arow[33 * (SLAB / 16)] = z;
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_remote(&sm.a_ready[s], 0);
                    if (EPW == 4) mbar_arrive_remote(&sm.a_ready[s], 0);   // the barrier counts 16 = the 8 encoder warps of each CTA
                }
                tr.ev(20 + L);
            }
            mbar_wait(&sm.acc_full[s], ph); ph ^= 1;
            tc_fence_after();
            tr.ev(13);
            epilogue_aggregate<KP, SAVE>(p, tacc_lane, p.wa, meta, sm.scratch[s], tile, row, half, 1 + s * 4 + (warp & 3),
                                         &sm.acc_empty[s]);
            tr.ev(23);
        }
    } else if (warp == ENCW + 2 * EPW) {
        // ===================================================== weight producer
        // Two issuing lanes (one lane alone sustains only a copy per ~420 clk): lane e loads chunk e of every group.
        if (lane < 2) {
            uint32_t g = 0;
            for_each_step(n_my, [&](int s, int L, int it) {
                (void)s; (void)it;
                for (int c0 = 0; c0 < layer_chunks(L); c0 += 2, g++) {
                    const uint32_t b = g % NGRP, phase = (g / NGRP) & 1;
                    const int nc = layer_chunks(L) - c0 < 2 ? 1 : 2;
                    mbar_wait(&sm.w_empty[b], phase ^ 1);
                    if (lane == 0) {
                        uint32_t total = 0;
                        for (int e = 0; e < nc; e++) total += (uint32_t)chunk_slabs(L, c0 + e) * 2048u;
                        mbar_arrive_expect_tx(&sm.w_full[b], total);
                    }
                    if (lane < nc) {
                        const int c = c0 + lane;
                        const uint32_t bytes = (uint32_t)chunk_slabs(L, c) * 2048u;            // this CTA's N-half of the chunk
                        bulk_g2s(sm.W[b][lane], p.wpack + layer_byte0(L) + (size_t)c * CHUNK_BYTES + rank * bytes, bytes, &sm.w_full[b]);
                    }
                }
            });
        }
    } else if (rank == 0) {
        // ===================================================== MMA issuer (leader CTA only)
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(2 * ROWS, HID);
            uint32_t g = 0, ar[2] = {0, 0}, ae[2] = {0, 0};
            for_each_step(n_my, [&](int s, int L, int it) {
                tr.ev(30 + 4 * s + L);
                ISSUER_WAIT(&sm.a_ready[s], ar[s]); ar[s] ^= 1;
                if (L == 0 && it >= 1) { ISSUER_WAIT(&sm.acc_empty[s], ae[s]); ae[s] ^= 1; }
                tc_fence_after();
                tr.ev(40 + 4 * s + L);
                const uint32_t tacc = tmem + (uint32_t)(s * HID);
                const uint32_t a_base = smem_u32(sm.A[s]);
                for (int c0 = 0; c0 < layer_chunks(L); c0 += 2, g++) {
                    const uint32_t b = g % NGRP, phase = (g / NGRP) & 1;
                    const int nc = layer_chunks(L) - c0 < 2 ? 1 : 2;
                    ISSUER_WAIT(&sm.w_full[b], phase);
                    ISSUER_WAIT(&sm.w_peer[b], phase);
                    tc_fence_after();
                    for (int e = 0; e < nc; e++) {
                        const int c = c0 + e;
                        const uint32_t b_base = smem_u32(sm.W[b][e]);
                        const int nk = chunk_slabs(L, c) >> 1;
                        for (int kk = 0; kk < nk; kk++) {
                            const uint64_t ad = make_smem_desc(a_base + (uint32_t)((c * CHUNK_SLABS + kk * 2) * SLAB), SLAB, 128);
                            const uint64_t bd = make_smem_desc(b_base + (uint32_t)(kk * 2 * (HID / 2) * 16), (HID / 2) * 16, 128);
                            mma_bf16_2cta(tacc, ad, bd, idesc, (uint32_t)((c | kk) > 0));
                        }
                    }
                    mma_commit2(&sm.w_empty[b], 3);     // one commit per group: a tcgen05.commit costs ~120 clk of MMA time
                }
                mma_commit2(&sm.acc_full[s], 3);
                if (L == 3) mma_commit2(&sm.a_free[s], 3);
                tr.ev(50 + 4 * s + L);
            });
        }
    } else {
        // ===================================================== relay (peer CTA): my half of the group landed -> tell the leader
        if (lane == 0) {
            uint32_t g = 0;
            for_each_step(n_my, [&](int s, int L, int it) {
                (void)s; (void)it;
                for (int c0 = 0; c0 < layer_chunks(L); c0 += 2, g++) {
                    mbar_wait(&sm.w_full[g % NGRP], (g / NGRP) & 1);
                    mbar_arrive_remote(&sm.w_peer[g % NGRP], 0);
                }
            });
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // nobody leaves while the peer may still signal it / the pair's MMAs are in flight
    if (warp == ENCW + 2 * EPW + 1) tmem_dealloc2(tmem, 512);
}

// ---------------------------------------------------------------------------------------------- colour network
// mlp_color 280 -> 128 -> 128 -> 128 (LeakyReLU) + rgb head 128 -> 3 (sigmoid, *1.002 - 0.001), SM:355-359.
// CTA pairs again (cta_group::2, M = 256 = 128 samples per CTA, N = 128): each CTA keeps only ITS 64 output features of the three
// weight matrices resident (68 KB instead of 136 KB), which leaves room for TWO sample-tile slots per CTA.  Warps 0-3 / 4-7 own
// slot 0 / 1 (thread = sample = TMEM lane): gather the bf16 features of field_tc_kernel into the A operand [F_s 256 | PE(v) 24 | 0 x 8],
// then the three epilogues; warp 8 of the leader issues the MMAs for whichever slot is ready (the weights are resident, so the order
// is free): one slot's gather and epilogues hide under the other's MMAs.  v1 (one slot, 128 threads, everything serial per tile) ran
// at 12.6 % tensor-pipe activity, 18 k clk per tile.
struct ColorParams {
    const uint8_t* F;              // 128-sample tiles of 32 k-slabs, written by field_tc_kernel
    const int* sample_ids;
    const float* dirs;
    const uint8_t* wpack_c;        // Wc1 | Wc2 | Wc3, each as two N-halves of k-slabs [k/8][64][8]
    const float *bc1, *bc2, *bc3, *wc4, *bc4;
    Cam cam;
    int S, SR, n_tiles;            // capacity when S_dev is set
    const int* S_dev;
    float slope;
    float* rgb;                    // (R*SR,3) by slot
    uint8_t* csave;                // training: operand tiles [C0 | C1 | C2 | C3] per 128-sample tile (tc_layout.cuh), or NULL
};

constexpr int C1H_BYTES = C1_BYTES / 2, C2H_BYTES = C2_BYTES / 2;
struct SmemC {
    uint8_t A[2][A_BYTES];
    uint8_t W1[C1H_BYTES];
    uint8_t W2[C2H_BYTES];
    uint8_t W3[C2H_BYTES];
    float bias[3][HC];
    float w4[3][HC];
    uint64_t bar_w, a_ready[2], acc_full[2], f_full[2];
    uint32_t tmem_base;
};
static_assert(sizeof(SmemC) <= 232448, "color_tc_kernel shared memory exceeds the 227 KB per-CTA limit");

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {   // non-blocking
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}

#ifndef PNERF_COLOR_SPLIT
#define PNERF_COLOR_SPLIT 1
#endif
template <bool SAVE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(288, 1) color_tc_kernel(const ColorParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    SmemC& sm = *reinterpret_cast<SmemC*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = (int)blockIdx.x >> 1, n_pairs = (int)gridDim.x >> 1;
    const int S = dyn_count(p.S_dev, p.S);
    const int n_super = ((S + ROWS - 1) / ROWS + 1) >> 1;
    const int n_my = n_super > pair ? (n_super - pair + n_pairs - 1) / n_pairs : 0;
    if (tid == 0) {
        mbar_init(&sm.bar_w, 1);
        for (int s = 0; s < 2; s++) { mbar_init(&sm.a_ready[s], 8); mbar_init(&sm.acc_full[s], 1); mbar_init(&sm.f_full[s], 1); }
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc2(&sm.tmem_base, 256);
    for (int i = tid; i < HC; i += 288) {
        sm.bias[0][i] = p.bc1[i]; sm.bias[1][i] = p.bc2[i]; sm.bias[2][i] = p.bc3[i];
        sm.w4[0][i] = p.wc4[i]; sm.w4[1][i] = p.wc4[HC + i]; sm.w4[2][i] = p.wc4[2 * HC + i];
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    if (tid == 0) {   // this CTA's halves of the weights: one bulk copy each, once
        mbar_arrive_expect_tx(&sm.bar_w, C1H_BYTES + 2 * C2H_BYTES);
        bulk_g2s(sm.W1, p.wpack_c + rank * C1H_BYTES, C1H_BYTES, &sm.bar_w);
        bulk_g2s(sm.W2, p.wpack_c + C1_BYTES + rank * C2H_BYTES, C2H_BYTES, &sm.bar_w);
        bulk_g2s(sm.W3, p.wpack_c + C1_BYTES + C2_BYTES + rank * C2H_BYTES, C2H_BYTES, &sm.bar_w);
    }
    constexpr int SJ = SLAB / 16;
    if (warp < 8) {
        // ===================================================== slot group: thread = sample row
        const int s = warp >> 2, row = tid & 127;
        const uint32_t tacc_lane = tmem + (uint32_t)(s * HC) + ((uint32_t)((warp & 3) * 32) << 16);
        uint4* Arow = reinterpret_cast<uint4*>(sm.A[s] + row * 16);
        uint32_t ph = 0, fph = 0;
        bool w_ready = false;
        // the F part of the A operand (slabs 0..31 = 64 KB, contiguous in global memory) arrives by bulk copy; the copy for the slot's
        // next tile is issued as soon as the last layer's MMAs have released the buffer
        auto tile_of = [&](int j) { return 2 * (pair + j * n_pairs) + (int)rank; };
        // in two halves: k-slabs 16..31 are free as soon as the tile's FIRST layer is done (layers 2 and 3 read slabs 0..15 only),
        // slabs 0..15 after its last layer; the first half only announces its bytes, the second one arrives
        auto fetch = [&](int j, int half) {
            const int ct = tile_of(j);
            if (ct * ROWS >= S) return;
            constexpr uint32_t HB = (uint32_t)(F_TILE_BYTES / 2);
            if (half == 1) mbar_expect_tx(&sm.f_full[s], HB);
            else mbar_arrive_expect_tx(&sm.f_full[s], HB);
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const uint32_t off = (uint32_t)half * HB + (uint32_t)q * (HB / 2);
                bulk_g2s(sm.A[s] + off, p.F + (int64_t)ct * F_TILE_BYTES + off, HB / 2, &sm.f_full[s]);
            }
        };
        if (row == 0 && s < n_my) { fetch(s, 1); fetch(s, 0); }
        for (int j = s; j < n_my; j += 2) {
            const int ctile = tile_of(j);
            const int si = ctile * ROWS + row;
            uint4* grow = SAVE ? reinterpret_cast<uint4*>(p.csave + (int64_t)ctile * CSAVE_TILE_BYTES + row * 16) : nullptr;
            int slot = -1;
            if (ctile * ROWS < S) { mbar_wait(&sm.f_full[s], fph); fph ^= 1; }
            if (si < S) {
                slot = __ldg(p.sample_ids + si);
                if (SAVE) {
#pragma unroll 8
                    for (int q = 0; q < 32; q++) grow[q * SJ] = Arow[q * SJ];
                }
                const int ray = slot / p.SR;
                const float rd[3] = {__ldg(p.dirs + 3 * (int64_t)ray), __ldg(p.dirs + 3 * (int64_t)ray + 1), __ldg(p.dirs + 3 * (int64_t)ray + 2)};
                float v[3];
                rot_w2c(p.cam, rd, v);
                float t[32];              // ori=True layout minus the raw copy: [sin (d-major, f-minor) 12 | cos 12] (SM:305-306)
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    float q8[8];
                    pe<4>(v[d], q8);
#pragma unroll
                    for (int f = 0; f < 4; f++) { t[d * 4 + f] = q8[2 * f]; t[12 + d * 4 + f] = q8[2 * f + 1]; }
                }
#pragma unroll
                for (int q = 24; q < 32; q++) t[q] = 0.f;
                t[24] = 1.f;      // column 280: multiplied by a zero weight here; the backward's weight-gradient GEMM reads it as the ones column of d bc1
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint4 pv = pack8(t + 8 * q);
                    Arow[(32 + q) * SJ] = pv;
                    if (SAVE) grow[(32 + q) * SJ] = pv;
                }
            } else {
                const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 4
                for (int q = 0; q < 36; q++) {
                    Arow[q * SJ] = z;
                    if (SAVE) grow[q * SJ] = z;
                }
            }
            if (!w_ready) { mbar_wait(&sm.bar_w, 0); w_ready = true; }   // the issuer may read this CTA's weights once we say "ready"
            fence_proxy_async();
            tc_fence_before();                                            // our tcgen05.ld of the previous tile are complete
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(&sm.a_ready[s], 0);
#pragma unroll 1
            for (int L = 0; L < 3; L++) {
                mbar_wait(&sm.acc_full[s], ph); ph ^= 1;
                tc_fence_after();
                if (row == 0 && j + 2 < n_my) {
                    if (PNERF_COLOR_SPLIT) {
                        if (L == 0) fetch(j + 2, 1);
                        if (L == 2) fetch(j + 2, 0);     // the tile's last MMAs are done with the A buffer
                    } else if (L == 2) {
                        fetch(j + 2, 1); fetch(j + 2, 0);
                    }
                }
                float r[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
                for (int c0 = 0; c0 < HC; c0 += 32) {
                    float v[32];
                    tmem_ld32(tacc_lane + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 32; q++) {
                        const float x = v[q] + sm.bias[L][c0 + q];
                        v[q] = fmaxf(x, x * p.slope);
                    }
                    if (SAVE) {   // this layer's activations: C1 / C2 / C3 region of the tile
                        uint4* gact = grow + (L == 0 ? CSAVE_C1 : (L == 1 ? CSAVE_C2 : CSAVE_C3)) * SJ;
#pragma unroll
                        for (int q = 0; q < 4; q++) gact[(c0 / 8 + q) * SJ] = pack8(v + 8 * q);
                    }
                    if (L < 2) {
#pragma unroll
                        for (int q = 0; q < 4; q++) Arow[(c0 / 8 + q) * SJ] = pack8(v + 8 * q);
                    } else {
#pragma unroll
                        for (int q = 0; q < 32; q++) {
                            r[0] = fmaf(v[q], sm.w4[0][c0 + q], r[0]); r[1] = fmaf(v[q], sm.w4[1][c0 + q], r[1]);
                            r[2] = fmaf(v[q], sm.w4[2][c0 + q], r[2]);
                        }
                    }
                }
                if (L < 2) {
                    fence_proxy_async();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(&sm.a_ready[s], 0);
                } else if (slot >= 0) {
#pragma unroll
                    for (int q = 0; q < 3; q++) {
                        const float x = r[q] + __ldg(p.bc4 + q);
                        p.rgb[3 * (int64_t)slot + q] = (1.f / (1.f + __expf(-x))) * (1.f + 2.f * 0.001f) - 0.001f;   // SM:359
                    }
                }
            }
        }
    } else if (rank == 0 && lane == 0) {
        // ===================================================== MMA issuer (leader): the next layer of whichever slot is ready
        const uint32_t idesc = make_idesc_bf16(2 * ROWS, HC);
        const int tot[2] = {3 * ((n_my + 1) >> 1), 3 * (n_my >> 1)};
        int step[2] = {0, 0};
        uint32_t ph[2] = {0, 0};
        uint32_t spins = 0;
        while (step[0] < tot[0] || step[1] < tot[1]) {
            bool any = false;
#pragma unroll
            for (int s = 0; s < 2; s++) {
                if (step[s] >= tot[s] || !mbar_test(&sm.a_ready[s], ph[s])) continue;
                ph[s] ^= 1; any = true;
                tc_fence_after();
                const int L = step[s] % 3;
                const uint32_t a_base = smem_u32(sm.A[s]);
                const uint32_t b_base = smem_u32(L == 0 ? sm.W1 : (L == 1 ? sm.W2 : sm.W3));
                const int nk = L == 0 ? KIN_PAD / 16 : HC / 16;
                for (int ks = 0; ks < nk; ks++) {
                    const uint64_t ad = make_smem_desc(a_base + (uint32_t)(ks * 2 * SLAB), SLAB, 128);
                    const uint64_t bd = make_smem_desc(b_base + (uint32_t)(ks * 2 * (HC / 2) * 16), (HC / 2) * 16, 128);
                    mma_bf16_2cta(tmem + (uint32_t)(s * HC), ad, bd, idesc, (uint32_t)(ks > 0));
                }
                mma_commit2(&sm.acc_full[s], 3);
                step[s]++;
            }
            if (!any) { __nanosleep(40); if (++spins > (1u << 26)) __trap(); }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 8) tmem_dealloc2(tmem, 256);
}

// ---------------------------------------------------------------------------------------------- weight packing
// fp32 nn.Linear weights (out,in) -> bf16 K-slab layout [k/8][out][8], K zero-padded; 34 field chunks then Wc1, Wc2, Wc3.
struct PackJob { const float* w; int out, in, kpad; int64_t dst_off; int split; const float* bias; int bias_col; };
struct PackJobs { PackJob j[7]; };

__global__ void __launch_bounds__(256) pack_weights_kernel(PackJobs jobs, uint8_t* __restrict__ dst) {
    const PackJob jb = jobs.j[blockIdx.y];
    const int total = jb.out * jb.kpad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int k = i % jb.kpad, n = i / jb.kpad;
        const float v = k < jb.in ? jb.w[(int64_t)n * jb.in + k] : ((jb.bias && k == jb.bias_col) ? jb.bias[n] : 0.f);
        __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst + jb.dst_off);
        if (jb.split == 1) {   // field layers: 48-k chunks of [N-half 2][k-slab <=6][128 features][8]: a CTA of a pair bulk-copies one half
            const int ks = k >> 3, h = n >> 7, c = ks / CHUNK_SLABS;
            const int nsl = jb.kpad / 8 - CHUNK_SLABS * c < CHUNK_SLABS ? jb.kpad / 8 - CHUNK_SLABS * c : CHUNK_SLABS;
            d[(int64_t)c * (CHUNK_BYTES / 2) + ((int64_t)h * nsl + (ks - c * CHUNK_SLABS)) * 1024 + (n & 127) * 8 + (k & 7)] = __float2bfloat16(v);
        } else if (jb.split == 2) {   // colour layers: two resident N-halves of k-slabs [k/8][64 features][8]
            d[(((int64_t)(n >> 6) * (jb.kpad >> 3) + (k >> 3)) * 64 + (n & 63)) * 8 + (k & 7)] = __float2bfloat16(v);
        } else {
            d[((int64_t)(k >> 3) * jb.out + n) * 8 + (k & 7)] = __float2bfloat16(v);
        }
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

static unsigned long long* g_trace = nullptr;
extern "C" int pnerf_tc_set_trace(void* buf) { g_trace = (unsigned long long*)buf; return PNERF_OK; }
extern "C" int64_t pnerf_tc_trace_bytes(void) { return (int64_t)32 * TRACE_PER_WARP * 8; }

extern "C" int64_t pnerf_tc_wpack_bytes(void) { return WPACK_BYTES; }

extern "C" int pnerf_tc_pack_weights(const pnerf_mlp* mlp, void* wpack, void* stream) {
    if (!mlp || !wpack) return PNERF_ERR_ARG;
    PackJobs jobs;
    jobs.j[0] = {mlp->w1, 256, 284, 288, 0, 1, mlp->b1, BIAS_COL[0]};
    jobs.j[1] = {mlp->w2, 256, 256, 272, (int64_t)layer_byte0(1), 1, mlp->b2, BIAS_COL[1]};
    jobs.j[2] = {mlp->w3, 256, 263, 272, (int64_t)layer_byte0(2), 1, mlp->b3, BIAS_COL[2]};
    jobs.j[3] = {mlp->w4, 256, 256, 272, (int64_t)layer_byte0(3), 1, mlp->b4, BIAS_COL[3]};
    jobs.j[4] = {mlp->wc1, 128, 280, 288, (int64_t)WPACK_FIELD_BYTES, 2, nullptr, 0};
    jobs.j[5] = {mlp->wc2, 128, 128, 128, (int64_t)WPACK_FIELD_BYTES + C1_BYTES, 2, nullptr, 0};
    jobs.j[6] = {mlp->wc3, 128, 128, 128, (int64_t)WPACK_FIELD_BYTES + C1_BYTES + C2_BYTES, 2, nullptr, 0};
    for (int i = 0; i < 4; i++) if (!jobs.j[i].bias) return PNERF_ERR_ARG;
    for (int i = 0; i < 7; i++) if (!jobs.j[i].w) return PNERF_ERR_ARG;
    pack_weights_kernel<<<dim3(64, 7), 256, 0, (cudaStream_t)stream>>>(jobs, (uint8_t*)wpack);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}

extern "C" int64_t pnerf_field_tc_workspace_bytes(int64_t n_samples) {
    return ((n_samples + ROWS - 1) / ROWS + 1) / 2 * 2 * F_TILE_BYTES + 256;      // whole 128-sample tiles, an even number of them (CTA pairs)
}

namespace pnerf {
// Fused per-neighbour networks (sigma by slot + F (S,256) bf16); with `save` != NULL every MMA operand is kept for the backward
// pass (training); with `color` the tensor-core colour network follows (inference).
int field_tc_launch(const pnerf_points* pts, const pnerf_camera* cam, const pnerf_mlp* mlp, const void* wpack, const pnerf_mode* mode,
                    const float* dirs, const float* sample_loc, const int* sample_pidx, const int* sample_ids, int S, const int* S_dev,
                    int SR, int K, float* sigma, float* rgb, void* F, uint8_t* save, float* save_w, float* save_raw, bool color,
                    uint8_t* csave, cudaStream_t st, int kp_override, int si0, bool field) {
    const int KP = kp_override > 0 ? kp_override : (K <= 8 ? 8 : (K <= 16 ? 16 : 32));
    FieldParams p;
    p.xyz = pts->xyz; p.embed = pts->embed; p.color = pts->color; p.dir = pts->dir; p.conf = pts->conf;
    p.dirs = dirs; p.sample_loc = sample_loc; p.sample_pidx = sample_pidx; p.sample_ids = sample_ids;
    p.wpack = (const uint8_t*)wpack;
    p.wa = mlp->wa; p.ba = mlp->ba;
    p.cam = make_cam(pts, cam);
    p.cam_dev = cam->dev;
    p.S = S; p.S_dev = S_dev; p.si0 = si0; p.SR = SR; p.K = K;
    const int spt = ROWS / KP;
    p.n_tiles = (S + spt - 1) / spt;
    p.slope = mode->lrelu_slope; p.softplus = mode->density_softplus; p.weight_conf = mode->weight_conf;
    p.sigma = sigma; p.F = (uint8_t*)F; p.trace = g_trace;
    p.save = save; p.save_w = save_w; p.save_raw = save_raw;
    const int n_super = (p.n_tiles + 1) / 2;
    const int grid = 2 * (n_super < kSMs / 2 ? n_super : kSMs / 2);   // CTA pairs
    const size_t smem = sizeof(Smem);
    auto launch = [&](auto kern) -> int {
        PNERF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, NT, smem, st>>>(p);
        PNERF_LAUNCH_CHECK();
        return PNERF_OK;
    };
#ifdef PNERF_TC_TIMING      // debug build (tools/sweep_field_tc.sh "-DPNERF_TC_TIMING"): in-stream time of the two kernels of the previous call
    static cudaEvent_t g_tev[3]; static bool tev_init = false; static float tacc[2] = {0.f, 0.f}; static int n_calls = 0;
    if (!tev_init) { for (auto& e : g_tev) cudaEventCreate(&e); tev_init = true; }
    else if (color) {
        cudaEventSynchronize(g_tev[2]);
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, g_tev[0], g_tev[1]); cudaEventElapsedTime(&b, g_tev[1], g_tev[2]);
        tacc[0] += a; tacc[1] += b;
        if (++n_calls % 7 == 0) { fprintf(stderr, "[tc timing] field %.3f ms colour %.3f ms per 7 calls\n", tacc[0], tacc[1]); tacc[0] = tacc[1] = 0.f; }
    }
    cudaEventRecord(g_tev[0], st);
#endif
    int rc = PNERF_OK;
    if (!field) rc = PNERF_OK;
    else if (save) rc = KP == 8 ? launch(field_tc_kernel<8, true>) : (KP == 16 ? launch(field_tc_kernel<16, true>) : (KP == 32 ? launch(field_tc_kernel<32, true>) : PNERF_ERR_ARG));
    else {
        switch (KP) {
            case 2: rc = launch(field_tc_kernel<2, false>); break;
            case 4: rc = launch(field_tc_kernel<4, false>); break;
            case 8: rc = launch(field_tc_kernel<8, false>); break;
            case 16: rc = launch(field_tc_kernel<16, false>); break;
            case 32: rc = launch(field_tc_kernel<32, false>); break;
            default: rc = PNERF_ERR_ARG;
        }
    }
    if (rc || !color) return rc;
#ifdef PNERF_TC_TIMING
    cudaEventRecord(g_tev[1], st);
#endif

    ColorParams c;
    c.F = p.F; c.sample_ids = sample_ids; c.dirs = dirs;
    c.wpack_c = (const uint8_t*)wpack + WPACK_FIELD_BYTES;
    c.bc1 = mlp->bc1; c.bc2 = mlp->bc2; c.bc3 = mlp->bc3; c.wc4 = mlp->wc4; c.bc4 = mlp->bc4;
    c.cam = p.cam; c.S = S; c.S_dev = S_dev; c.SR = SR; c.n_tiles = (S + ROWS - 1) / ROWS;
    c.slope = mode->lrelu_slope; c.rgb = rgb; c.csave = csave;
    const int c_super = (c.n_tiles + 1) / 2;
    const int cgrid = 2 * (c_super < kSMs / 2 ? c_super : kSMs / 2);   // CTA pairs
    const size_t csmem = sizeof(SmemC);
    if (csave) {
        PNERF_CUDA(cudaFuncSetAttribute(color_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
        color_tc_kernel<true><<<cgrid, 288, csmem, st>>>(c);
    } else {
        PNERF_CUDA(cudaFuncSetAttribute(color_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
        color_tc_kernel<false><<<cgrid, 288, csmem, st>>>(c);
    }
    PNERF_LAUNCH_CHECK();
#ifdef PNERF_TC_TIMING
    cudaEventRecord(g_tev[2], st);
#endif
    return PNERF_OK;
}
}  // namespace pnerf

extern "C" int pnerf_field_forward_tc(const pnerf_points* pts, const pnerf_camera* cam, const pnerf_mlp* mlp, const void* wpack,
                                      const pnerf_mode* mode, const float* dirs, const float* sample_loc, const int* sample_pidx,
                                      const int* sample_ids, int S, int SR, int K, float* sigma, float* rgb, void* workspace,
                                      int64_t workspace_bytes, void* stream) {
    if (!pts || !cam || !mlp || !wpack || !mode || S < 0 || K <= 0 || K > 32 || SR <= 0) return PNERF_ERR_ARG;
    if (S == 0) return PNERF_OK;
    if (!workspace || workspace_bytes < pnerf_field_tc_workspace_bytes(S)) return PNERF_ERR_WORKSPACE;
    if (!(mode->lrelu_slope > 0.f && mode->lrelu_slope < 1.f)) return PNERF_ERR_ARG;   // lrelu(x) = max(x, slope x)
    return field_tc_launch(pts, cam, mlp, wpack, mode, dirs, sample_loc, sample_pidx, sample_ids, S, nullptr, SR, K, sigma, rgb, workspace,
                           nullptr, nullptr, nullptr, true, nullptr, (cudaStream_t)stream, 0, 0, true);
}

extern "C" int pnerf_field_forward_tc_part(const pnerf_points* pts, const pnerf_camera* cam, const pnerf_mlp* mlp, const void* wpack,
                                           const pnerf_mode* mode, const float* dirs, const float* sample_loc, const int* sample_pidx,
                                           const int* sample_ids, int S, int rows_per_sample, int first_sample, int SR, int K, float* sigma,
                                           void* workspace, int64_t workspace_bytes, void* stream) {
    if (!pts || !cam || !mlp || !wpack || !mode || S < 0 || K <= 0 || K > 32 || SR <= 0 || first_sample < 0) return PNERF_ERR_ARG;
    const int kp = rows_per_sample;
    if (kp != 2 && kp != 4 && kp != 8 && kp != 16 && kp != 32) return PNERF_ERR_ARG;
    if (S == 0) return PNERF_OK;
    if (!workspace || workspace_bytes < pnerf_field_tc_workspace_bytes((int64_t)first_sample + S)) return PNERF_ERR_WORKSPACE;
    if (!(mode->lrelu_slope > 0.f && mode->lrelu_slope < 1.f)) return PNERF_ERR_ARG;
    return field_tc_launch(pts, cam, mlp, wpack, mode, dirs, sample_loc, sample_pidx, sample_ids, S, nullptr, SR, K, sigma, nullptr, workspace,
                           nullptr, nullptr, nullptr, false, nullptr, (cudaStream_t)stream, kp, first_sample, true);
}

extern "C" int pnerf_color_forward_tc(const pnerf_points* pts, const pnerf_camera* cam, const pnerf_mlp* mlp, const void* wpack,
                                      const pnerf_mode* mode, const float* dirs, const int* sample_ids, int S, int SR, float* rgb,
                                      const void* workspace, int64_t workspace_bytes, void* stream) {
    if (!pts || !cam || !mlp || !wpack || !mode || S < 0 || SR <= 0) return PNERF_ERR_ARG;
    if (S == 0) return PNERF_OK;
    if (!workspace || !rgb || workspace_bytes < pnerf_field_tc_workspace_bytes(S)) return PNERF_ERR_WORKSPACE;
    return field_tc_launch(pts, cam, mlp, wpack, mode, dirs, nullptr, nullptr, sample_ids, S, nullptr, SR, 8, nullptr, rgb, (void*)workspace,
                           nullptr, nullptr, nullptr, true, nullptr, (cudaStream_t)stream, 0, 0, false);
}
