// fp32 SIMT implementation of the field networks, forward and backward (rows P, GA, W, E, M1, A, M2, BWD).
// This is the exact-parity path (fp32 everywhere, like the reference); the tensor-core path lives in
// field_tc.cu.  Replaces SU:190-209 (gather) and SM:270-366 (dists, weights, encodings, three MLPs,
// heads, K-aggregation); original-flow twin PA:486-662,745-830.
//
// Row r = i*K + k is neighbour slot k of the i-th valid sample.  Masked slots (pidx < 0) are kept as
// all-zero rows with zero aggregation weight, so they add nothing to any output or gradient (the
// reference removes them with boolean indexing, SM:310-315).
#include "pnerf_common.cuh"

namespace pnerf {
namespace {

constexpr int C_FEAT = 32;      // point_features_dim (SM:78)
constexpr int F_FEAT = 3;       // num_feat_freqs (SM:74)
constexpr int F_DIST = 5;       // num_dist_freqs (SM:75)
constexpr int F_VIEW = 4;       // num_viewdir_freqs (SM:73)
constexpr int IN1 = 284, LD0 = 288;    // mlp_base input  [feat 32 | PE(feat) 192 | PE(dists6) 60] + pad
constexpr int HID = 256;
constexpr int IN3 = 263, LD2 = 272;    // mlp_head input  [h 256 | color 3 | dir - v 3 | <dir,v> 1] + pad
constexpr int INC = 280, LDC = 288;    // mlp_color input [F_s 256 | PE(v) 24] + pad
constexpr int HC = 128;

struct Ws {   // workspace carve-up (floats)
    float *X0, *H1, *X2, *H3, *G, *w, *wn, *araw;            // per row
    float *C0, *C1, *C2, *C3;                                  // per sample
    float *T0, *T1;                                            // backward ping-pong, per row x 288
};

__host__ int64_t ws_floats(int64_t S, int K) {
    int64_t M = S * K;
    return M * (LD0 + HID + LD2 + HID + HID + 3) + S * (LDC + 3 * HC) + 2 * M * LD0 + 64;
}
__host__ Ws carve(void* base, int64_t S, int K) {
    int64_t M = S * K;
    float* p = (float*)base;
    Ws w;
    w.X0 = p; p += M * LD0;
    w.H1 = p; p += M * HID;
    w.X2 = p; p += M * LD2;
    w.H3 = p; p += M * HID;
    w.G = p; p += M * HID;
    w.w = p; p += M;
    w.wn = p; p += M;
    w.araw = p; p += M;
    w.C0 = p; p += S * LDC;
    w.C1 = p; p += S * HC;
    w.C2 = p; p += S * HC;
    w.C3 = p; p += S * HC;
    w.T0 = p; p += M * LD0;
    w.T1 = p; p += M * LD0;
    return w;
}

struct Cam { float o[3]; float Rc[9]; float Rw[9]; };

// cam = R_c2w^T (p - o) with mul-then-add roundings ((a+b)+c), like torch.sum(a*b) at SU:131,140
__device__ __forceinline__ void to_pers(const Cam& c, float x, float y, float z, float& px, float& py, float& pz) {
    const float sx = __fsub_rn(x, c.o[0]), sy = __fsub_rn(y, c.o[1]), sz = __fsub_rn(z, c.o[2]);
    float cam[3];
#pragma unroll
    for (int j = 0; j < 3; j++)
        cam[j] = __fadd_rn(__fadd_rn(__fmul_rn(sx, c.Rc[j]), __fmul_rn(sy, c.Rc[3 + j])), __fmul_rn(sz, c.Rc[6 + j]));
    px = __fdiv_rn(cam[0], cam[2]);
    py = __fdiv_rn(cam[1], cam[2]);
    pz = cam[2];
}
// u . Rw2c^T  (row vector times Rn = Rw2c^T, SM:303-304,312,330)
__device__ __forceinline__ void rot_w2c(const Cam& c, const float* u, float* out) {
#pragma unroll
    for (int j = 0; j < 3; j++) out[j] = u[0] * c.Rw[3 * j] + u[1] * c.Rw[3 * j + 1] + u[2] * c.Rw[3 * j + 2];
}

// One warp per valid sample: geometry, weights, extras and the 284-wide encoded input of its K rows.
__global__ void __launch_bounds__(256) encode_kernel(Cam cam, const float* __restrict__ xyz, const float* __restrict__ embed,
                                                      const float* __restrict__ color, const float* __restrict__ dir,
                                                      const float* __restrict__ conf, int weight_conf,
                                                      const float* __restrict__ dirs, const float* __restrict__ sample_loc,
                                                      const int* __restrict__ sample_pidx, const int* __restrict__ sample_ids,
                                                      int S, int SR, int K, Ws ws) {
    __shared__ float s_d6[8][32][6];
    __shared__ int s_idx[8][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < S; i += warps) {
        const int slot = sample_ids[i];
        const int ray = slot / SR;
        const float sx = sample_loc[3 * (int64_t)slot], sy = sample_loc[3 * (int64_t)slot + 1], sz = sample_loc[3 * (int64_t)slot + 2];
        const float rd[3] = {dirs[3 * ray], dirs[3 * ray + 1], dirs[3 * ray + 2]};
        float v[3];
        rot_w2c(cam, rd, v);
        float wraw = 0.f, cc = 1.f;
        int p = -1;
        const int64_t row = (int64_t)i * K + lane;
        if (lane < K) {
            p = sample_pidx[(int64_t)slot * K + lane];
            const int idx = max(p, 0);
            const float X = xyz[3 * (int64_t)idx], Y = xyz[3 * (int64_t)idx + 1], Z = xyz[3 * (int64_t)idx + 2];
            float spx, spy, spz, ppx, ppy, ppz;
            to_pers(cam, sx, sy, sz, spx, spy, spz);
            to_pers(cam, X, Y, Z, ppx, ppy, ppz);
            float d[6];
            d[0] = __fsub_rn(X, sx); d[1] = __fsub_rn(Y, sy); d[2] = __fsub_rn(Z, sz);
            d[3] = __fsub_rn(__fmul_rn(ppx, ppz), __fmul_rn(spx, spz));
            d[4] = __fsub_rn(__fmul_rn(ppy, ppz), __fmul_rn(spy, spz));
            d[5] = __fsub_rn(ppz, spz);
            const float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            wraw = p >= 0 ? 1.f / fmaxf(nrm, 1e-6f) : 0.f;                       // SM:471-474
            float d6[3];
            rot_w2c(cam, d, d6);                                                // SM:312
            float* e7 = ws.X2 + row * LD2 + HID;
            if (p >= 0) {
                const float col[3] = {color[3 * (int64_t)idx], color[3 * (int64_t)idx + 1], color[3 * (int64_t)idx + 2]};
                const float dd[3] = {dir[3 * (int64_t)idx], dir[3 * (int64_t)idx + 1], dir[3 * (int64_t)idx + 2]};
                float dr[3];
                rot_w2c(cam, dd, dr);                                           // SM:330
                e7[0] = col[0]; e7[1] = col[1]; e7[2] = col[2];
                e7[3] = dr[0] - v[0]; e7[4] = dr[1] - v[1]; e7[5] = dr[2] - v[2];
                e7[6] = dr[0] * v[0] + dr[1] * v[1] + dr[2] * v[2];            // SM:334
                cc = fminf(fmaxf(conf[idx], 1e-4f), 1.f);                       // PA:740-742
            } else {
#pragma unroll
                for (int j = 0; j < 7; j++) e7[j] = 0.f;
            }
#pragma unroll
            for (int j = 7; j < LD2 - HID; j++) e7[j] = 0.f;
            s_d6[wib][lane][0] = d6[0]; s_d6[wib][lane][1] = d6[1]; s_d6[wib][lane][2] = d6[2];
            s_d6[wib][lane][3] = d[3]; s_d6[wib][lane][4] = d[4]; s_d6[wib][lane][5] = d[5];
            s_idx[wib][lane] = p;
        }
        float wsum = wraw;
#pragma unroll
        for (int o = 16; o; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
        if (lane < K) {
            const float wn = wraw / fmaxf(wsum, 1e-8f);                          // SM:286
            ws.wn[row] = wn;
            ws.w[row] = weight_conf ? wn * cc : wn;                              // PA:826 vs SM:318
        }
        __syncwarp();
        for (int k = 0; k < K; k++) {
            float* x = ws.X0 + ((int64_t)i * K + k) * LD0;
            const int pk = s_idx[wib][k];
            if (pk < 0) {
                for (int c = lane; c < LD0; c += 32) x[c] = 0.f;
                continue;
            }
            const float e = embed[(int64_t)pk * C_FEAT + lane];
            x[lane] = e;
#pragma unroll
            for (int f = 0; f < F_FEAT; f++) {                                   // SU:61-67, ori=False layout
                float s, c;
                sincosf(e * (float)(1 << f), &s, &c);
                x[C_FEAT + (lane * F_FEAT + f) * 2] = s;
                x[C_FEAT + (lane * F_FEAT + f) * 2 + 1] = c;
            }
            if (lane < 6 * F_DIST) {
                const int d = lane / F_DIST, f = lane % F_DIST;
                float s, c;
                sincosf(s_d6[wib][k][d] * (float)(1 << f), &s, &c);
                x[C_FEAT + 2 * F_FEAT * C_FEAT + lane * 2] = s;
                x[C_FEAT + 2 * F_FEAT * C_FEAT + lane * 2 + 1] = c;
            }
            if (lane < LD0 - IN1) x[IN1 + lane] = 0.f;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ SGEMM
// C[m,n] (op)= sum_k A(m,k) B(k,n).  TA: A(m,k) = A[k*lda+m] else A[m*lda+k].  TB: B(k,n) = B[n*ldb+k] else B[k*ldb+n].
// EPI 0: C = acc;  1: C = lrelu(acc + bias[n]);  2: atomicAdd(C, acc) (split-K over blockIdx.z).
template <bool TA, bool TB, int EPI>
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                     float* __restrict__ C, int ldc, int64_t M, int N, int64_t Kd,
                                                     const float* __restrict__ bias, float slope, int64_t k_chunk) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int64_t kb = (int64_t)blockIdx.z * k_chunk, ke = min(Kd, kb + k_chunk);
    const int tx = tid % 16, ty = tid / 16;   // thread computes rows ty*4..+4, cols tx*4..+4
    float acc[4][4] = {};
    for (int64_t k0 = kb; k0 < ke; k0 += BK) {
#pragma unroll
        for (int e = tid; e < BM * BK; e += 256) {
            int mm, kk;
            if (TA) { mm = e % BM; kk = e / BM; } else { kk = e % BK; mm = e / BK; }
            const int64_t m = m0 + mm, k = k0 + kk;
            float v = 0.f;
            if (m < M && k < ke) v = TA ? A[k * lda + m] : A[m * lda + k];
            As[kk][mm] = v;
        }
#pragma unroll
        for (int e = tid; e < BN * BK; e += 256) {
            int nn, kk;
            if (TB) { kk = e % BK; nn = e / BK; } else { nn = e % BN; kk = e / BN; }
            const int n = n0 + nn;
            const int64_t k = k0 + kk;
            float v = 0.f;
            if (n < N && k < ke) v = TB ? B[(int64_t)n * ldb + k] : B[k * ldb + n];
            Bs[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (EPI == 1) { v += bias[n]; v = v > 0.f ? v : v * slope; }
            if (EPI == 2) atomicAdd(C + m * ldc + n, v); else C[m * ldc + n] = v;
        }
    }
}

// Y = lrelu(X W^T + b): X (M,ldx) first Kd cols, W (N,Kd) torch layout
int linear_fwd(const float* X, int ldx, const float* W, const float* b, float* Y, int ldy, int64_t M, int N, int Kd,
               float slope, cudaStream_t st) {
    if (M == 0) return PNERF_OK;
    dim3 grid((unsigned)((M + 63) / 64), (N + 63) / 64, 1);
    sgemm_kernel<false, true, 1><<<grid, 256, 0, st>>>(X, ldx, W, Kd, Y, ldy, M, N, Kd, b, slope, Kd);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}

// dZ = dY * lrelu'(Y) in place on dY (first N cols), and db[n] += sum_m dZ[m,n]
__global__ void __launch_bounds__(256) lrelu_bwd_kernel(float* __restrict__ dY, int lddy, const float* __restrict__ Y, int ldy,
                                                         int64_t M, int N, float slope, float* __restrict__ db) {
    // block handles 64 rows x N cols; thread t handles column t (N <= 256)
    const int n = threadIdx.x;
    if (n >= N) return;
    const int64_t m0 = (int64_t)blockIdx.x * 64;
    float s = 0.f;
    for (int64_t m = m0; m < min(M, m0 + 64); m++) {
        float g = dY[m * lddy + n];
        g = Y[m * ldy + n] > 0.f ? g : g * slope;
        dY[m * lddy + n] = g;
        s += g;
    }
    if (db) atomicAdd(db + n, s);
}

// backward of Y = lrelu(X W^T + b) given dY (overwritten with dZ): dW += dZ^T X, db += colsum dZ, dX = dZ W
int linear_bwd(float* dY, int lddy, const float* Y, int ldy, const float* X, int ldx, const float* W, float* dW, float* db,
               float* dX, int lddx, int64_t M, int N, int Kd, float slope, cudaStream_t st) {
    if (M == 0) return PNERF_OK;
    lrelu_bwd_kernel<<<(unsigned)((M + 63) / 64), 256, 0, st>>>(dY, lddy, Y, ldy, M, N, slope, db);
    PNERF_LAUNCH_CHECK();
    if (dW) {
        const int64_t chunk = M >= 262144 ? 2048 : 256;   // split-K rows per block: keep >= a few hundred blocks at small M
        dim3 grid((N + 63) / 64, (Kd + 63) / 64, (unsigned)((M + chunk - 1) / chunk));
        // C[n,kd] += sum_m dZ[m,n] X[m,kd] : A(m'=n,k=m) = dZ[m*lddy+n] (TA), B(k=m,n'=kd) = X[m*ldx+kd]
        sgemm_kernel<true, false, 2><<<grid, 256, 0, st>>>(dY, lddy, X, ldx, dW, Kd, N, Kd, M, nullptr, 0.f, chunk);
        PNERF_LAUNCH_CHECK();
    }
    if (dX) {
        dim3 grid((unsigned)((M + 63) / 64), (Kd + 63) / 64, 1);
        // dX[m,kd] = sum_n dZ[m,n] W[n,kd]
        sgemm_kernel<false, false, 0><<<grid, 256, 0, st>>>(dY, lddy, W, Kd, dX, lddx, M, Kd, N, nullptr, 0.f, N);
        PNERF_LAUNCH_CHECK();
    }
    return PNERF_OK;
}

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }   // torch Softplus(beta=1, threshold=20)
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// One warp per valid sample: density head on each row, then sigma_s = sum_k w a, F_s = sum_k w g; also the
// view encoding of the colour network input (SM:337-356).
__global__ void __launch_bounds__(256) aggregate_kernel(Cam cam, const float* __restrict__ wa, const float* __restrict__ ba,
                                                         int softplus, const float* __restrict__ dirs,
                                                         const int* __restrict__ sample_ids, int S, int SR, int K, Ws ws,
                                                         float* __restrict__ sigma) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    float wa_r[HID / 32];
#pragma unroll
    for (int j = 0; j < HID / 32; j++) wa_r[j] = wa[lane + 32 * j];
    const float b = ba[0];
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < S; i += warps) {
        float Fs[HID / 32] = {};
        float sg = 0.f;
        for (int k = 0; k < K; k++) {
            const int64_t row = (int64_t)i * K + k;
            const float* g = ws.G + row * HID;
            float gv[HID / 32], dot = 0.f;
#pragma unroll
            for (int j = 0; j < HID / 32; j++) { gv[j] = g[lane + 32 * j]; dot = fmaf(gv[j], wa_r[j], dot); }
#pragma unroll
            for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            const float raw = dot + b;
            if (lane == 0) ws.araw[row] = raw;
            const float a = softplus ? softplus_f(raw - 1.f) : fmaxf(raw, 0.f);
            const float w = ws.w[row];
            sg = fmaf(w, a, sg);
#pragma unroll
            for (int j = 0; j < HID / 32; j++) Fs[j] = fmaf(w, gv[j], Fs[j]);
        }
        const int slot = sample_ids[i];
        if (lane == 0) sigma[slot] = sg;
        float* c0 = ws.C0 + (int64_t)i * LDC;
#pragma unroll
        for (int j = 0; j < HID / 32; j++) c0[lane + 32 * j] = Fs[j];
        const int ray = slot / SR;
        const float rd[3] = {dirs[3 * ray], dirs[3 * ray + 1], dirs[3 * ray + 2]};
        float v[3];
        rot_w2c(cam, rd, v);
        if (lane < 3 * F_VIEW) {                      // ori=True layout minus the raw copy (SM:305-306)
            const int d = lane / F_VIEW, f = lane % F_VIEW;
            float s, c;
            sincosf(v[d] * (float)(1 << f), &s, &c);
            c0[HID + lane] = s;
            c0[HID + 3 * F_VIEW + lane] = c;
        }
        if (lane < LDC - INC) c0[INC + lane] = 0.f;
    }
}

__global__ void __launch_bounds__(256) rgb_head_kernel(const float* __restrict__ C3, const float* __restrict__ wc4,
                                                        const float* __restrict__ bc4, const int* __restrict__ sample_ids,
                                                        int S, float* __restrict__ rgb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    const float* h = C3 + (int64_t)i * HC;
    float r[3] = {bc4[0], bc4[1], bc4[2]};
    for (int c = 0; c < HC; c++) {
        const float x = h[c];
        r[0] = fmaf(x, wc4[c], r[0]); r[1] = fmaf(x, wc4[HC + c], r[1]); r[2] = fmaf(x, wc4[2 * HC + c], r[2]);
    }
    const int slot = sample_ids[i];
#pragma unroll
    for (int j = 0; j < 3; j++) rgb[3 * (int64_t)slot + j] = sigmoid_f(r[j]) * (1.f + 2.f * 0.001f) - 0.001f;   // SM:359
}

// ------------------------------------------------------------------------------------------------ backward
// rgb head backward: dC3 (S,128) and dwc4/dbc4
__global__ void __launch_bounds__(128) rgb_head_bwd_kernel(const float* __restrict__ C3, const float* __restrict__ wc4,
                                                            const float* __restrict__ bc4, const int* __restrict__ sample_ids,
                                                            int S, const float* __restrict__ d_rgb, float* __restrict__ dC3,
                                                            float* __restrict__ dwc4, float* __restrict__ dbc4) {
    // block = 128 threads = one column each; loops over a chunk of samples
    const int c = threadIdx.x;
    const int i0 = blockIdx.x * 64, i1 = min(S, i0 + 64);
    float gw[3] = {0.f, 0.f, 0.f}, gb[3] = {0.f, 0.f, 0.f};
    __shared__ float s_raw[3];
    for (int i = i0; i < i1; i++) {
        const float* h = C3 + (int64_t)i * HC;
        // every thread recomputes nothing: reduce the 3 raw outputs with a block reduction
        float part[3] = {h[c] * wc4[c], h[c] * wc4[HC + c], h[c] * wc4[2 * HC + c]};
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int o = 16; o; o >>= 1) part[j] += __shfl_xor_sync(0xffffffffu, part[j], o);
        __shared__ float s_part[4][3];
        if ((c & 31) == 0) { s_part[c >> 5][0] = part[0]; s_part[c >> 5][1] = part[1]; s_part[c >> 5][2] = part[2]; }
        __syncthreads();
        if (c < 3) s_raw[c] = s_part[0][c] + s_part[1][c] + s_part[2][c] + s_part[3][c] + bc4[c];
        __syncthreads();
        const int slot = sample_ids[i];
        float dx = 0.f;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const float sg = sigmoid_f(s_raw[j]);
            const float dr = d_rgb[3 * (int64_t)slot + j] * (1.f + 2.f * 0.001f) * sg * (1.f - sg);
            dx = fmaf(dr, wc4[j * HC + c], dx);
            gw[j] = fmaf(dr, h[c], gw[j]);
            gb[j] += dr;
        }
        dC3[(int64_t)i * HC + c] = dx;
        __syncthreads();
    }
    if (dwc4) {
#pragma unroll
        for (int j = 0; j < 3; j++) atomicAdd(dwc4 + j * HC + c, gw[j]);
        if (c == 0) { atomicAdd(dbc4, gb[0]); atomicAdd(dbc4 + 1, gb[1]); atomicAdd(dbc4 + 2, gb[2]); }
    }
}

// aggregation backward: dG rows from dF_s (first 256 cols of dC0) and d sigma; d wa, d ba; d w (for conf).
__global__ void __launch_bounds__(256) aggregate_bwd_kernel(const float* __restrict__ wa, int softplus,
                                                             const int* __restrict__ sample_ids, int S, int K, Ws ws,
                                                             const float* __restrict__ dC0, int lddc0,
                                                             const float* __restrict__ d_sigma, float* __restrict__ dG,
                                                             float* __restrict__ dwa, float* __restrict__ dba,
                                                             float* __restrict__ dw_rows) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    float wa_r[HID / 32], gwa[HID / 32] = {};
    float gba = 0.f;
#pragma unroll
    for (int j = 0; j < HID / 32; j++) wa_r[j] = wa[lane + 32 * j];
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < S; i += warps) {
        const float ds = d_sigma[sample_ids[i]];
        float dF[HID / 32];
#pragma unroll
        for (int j = 0; j < HID / 32; j++) dF[j] = dC0[(int64_t)i * lddc0 + lane + 32 * j];
        for (int k = 0; k < K; k++) {
            const int64_t row = (int64_t)i * K + k;
            const float w = ws.w[row], raw = ws.araw[row];
            const float a = softplus ? softplus_f(raw - 1.f) : fmaxf(raw, 0.f);
            const float dact = softplus ? sigmoid_f(raw - 1.f) : (raw > 0.f ? 1.f : 0.f);
            const float draw = ds * w * dact;
            const float* g = ws.G + row * HID;
            float dwk = 0.f;
#pragma unroll
            for (int j = 0; j < HID / 32; j++) {
                const float gv = g[lane + 32 * j];
                dG[row * HID + lane + 32 * j] = fmaf(w, dF[j], draw * wa_r[j]);
                gwa[j] = fmaf(draw, gv, gwa[j]);
                dwk = fmaf(dF[j], gv, dwk);
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) dwk += __shfl_xor_sync(0xffffffffu, dwk, o);
            if (lane == 0) { gba += draw; if (dw_rows) dw_rows[row] = dwk + ds * a; }
        }
    }
    if (dwa) {
#pragma unroll
        for (int j = 0; j < HID / 32; j++) atomicAdd(dwa + lane + 32 * j, gwa[j]);
        if (lane == 0) atomicAdd(dba, gba);
    }
}

// scatter of the row gradients into the point tensors (index_select backward, SU:199-205):
//   embed: raw 32 inputs + through PE (d sin(2^f x) = 2^f cos, d cos = -2^f sin), color, dir (rotated back), conf
__global__ void __launch_bounds__(256) scatter_kernel(Cam cam, const int* __restrict__ sample_pidx,
                                                       const int* __restrict__ sample_ids, int S, int K, Ws ws,
                                                       const float* __restrict__ dX0, const float* __restrict__ dX2,
                                                       const float* __restrict__ dw_rows, const float* __restrict__ dirs, int SR,
                                                       float* __restrict__ g_embed, float* __restrict__ g_color,
                                                       float* __restrict__ g_dir, float* __restrict__ g_conf, int weight_conf) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t M = (int64_t)S * K;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < M; row += warps) {
        const int i = (int)(row / K), k = (int)(row % K);
        const int slot = sample_ids[i];
        const int p = sample_pidx[(int64_t)slot * K + k];
        if (p < 0) continue;
        if (g_embed) {
            const float* x = ws.X0 + row * LD0;
            const float* dx = dX0 + row * LD0;
            float g = dx[lane];
#pragma unroll
            for (int f = 0; f < F_FEAT; f++) {
                const float sc = (float)(1 << f);
                const int c = C_FEAT + (lane * F_FEAT + f) * 2;
                g = fmaf(dx[c], sc * x[c + 1], g);        // d sin = 2^f cos
                g = fmaf(dx[c + 1], -sc * x[c], g);       // d cos = -2^f sin
            }
            atomicAdd(g_embed + (int64_t)p * C_FEAT + lane, g);
        }
        const float* de = dX2 + row * LD2 + HID;
        if (g_color && lane < 3) atomicAdd(g_color + 3 * (int64_t)p + lane, de[lane]);
        if (g_dir && lane < 3) {
            // dr = dir . Rn ; e[3+j] = dr_j - v_j ; e[6] = <dr, v>  ->  d dr_j = de[3+j] + de[6] v_j ; d dir_i = sum_j d dr_j Rn[i][j]
            const int ray = slot / SR;
            const float rd[3] = {dirs[3 * ray], dirs[3 * ray + 1], dirs[3 * ray + 2]};
            float v[3];
            rot_w2c(cam, rd, v);
            float gi = 0.f;
#pragma unroll
            for (int j = 0; j < 3; j++) gi = fmaf(de[3 + j] + de[6] * v[j], cam.Rw[3 * j + lane], gi);   // Rn[i][j] = Rw2c[j][i]
            atomicAdd(g_dir + 3 * (int64_t)p + lane, gi);
        }
        if (g_conf && weight_conf && lane == 0) atomicAdd(g_conf + p, dw_rows[row] * ws.wn[row]);   // straight-through clamp
    }
}

Cam make_cam(const pnerf_points* pts, const pnerf_camera* cam) {
    Cam c;
    for (int i = 0; i < 3; i++) c.o[i] = cam->origin[i];
    for (int i = 0; i < 9; i++) { c.Rc[i] = cam->R_c2w[i]; c.Rw[i] = pts->Rw2c[i]; }
    return c;
}
int grid_warps(int64_t n_warps) {
    int64_t b = (n_warps * 32 + 255) / 256;
    int64_t cap = (int64_t)kSMs * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
}  // namespace
}  // namespace pnerf

using namespace pnerf;

extern "C" int64_t pnerf_field_f32_workspace_bytes(int64_t n_samples, int K) { return ws_floats(n_samples, K) * 4; }

extern "C" int pnerf_field_forward_f32(const pnerf_points* pts, const pnerf_camera* cam, const pnerf_mlp* mlp,
                                       const pnerf_mode* mode, const float* dirs, const float* sample_loc,
                                       const int* sample_pidx, const int* sample_ids, int S, int SR, int K, float* sigma,
                                       float* rgb, void* workspace, int64_t workspace_bytes, void* stream) {
    if (!pts || !cam || !mlp || !mode || S < 0 || K <= 0 || K > 32 || SR <= 0) return PNERF_ERR_ARG;
    if (S == 0) return PNERF_OK;
    if (!workspace || workspace_bytes < pnerf_field_f32_workspace_bytes(S, K)) return PNERF_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const Cam c = make_cam(pts, cam);
    Ws ws = carve(workspace, S, K);
    const int64_t M = (int64_t)S * K;
    const float sl = mode->lrelu_slope;
    encode_kernel<<<grid_warps(S), 256, 0, st>>>(c, pts->xyz, pts->embed, pts->color, pts->dir, pts->conf, mode->weight_conf,
                                                 dirs, sample_loc, sample_pidx, sample_ids, S, SR, K, ws);
    PNERF_LAUNCH_CHECK();
    int rc;
    if ((rc = linear_fwd(ws.X0, LD0, mlp->w1, mlp->b1, ws.H1, HID, M, HID, IN1, sl, st))) return rc;
    if ((rc = linear_fwd(ws.H1, HID, mlp->w2, mlp->b2, ws.X2, LD2, M, HID, HID, sl, st))) return rc;
    if ((rc = linear_fwd(ws.X2, LD2, mlp->w3, mlp->b3, ws.H3, HID, M, HID, IN3, sl, st))) return rc;
    if ((rc = linear_fwd(ws.H3, HID, mlp->w4, mlp->b4, ws.G, HID, M, HID, HID, sl, st))) return rc;
    aggregate_kernel<<<grid_warps(S), 256, 0, st>>>(c, mlp->wa, mlp->ba, mode->density_softplus, dirs, sample_ids, S, SR, K, ws, sigma);
    PNERF_LAUNCH_CHECK();
    if ((rc = linear_fwd(ws.C0, LDC, mlp->wc1, mlp->bc1, ws.C1, HC, S, HC, INC, sl, st))) return rc;
    if ((rc = linear_fwd(ws.C1, HC, mlp->wc2, mlp->bc2, ws.C2, HC, S, HC, HC, sl, st))) return rc;
    if ((rc = linear_fwd(ws.C2, HC, mlp->wc3, mlp->bc3, ws.C3, HC, S, HC, HC, sl, st))) return rc;
    rgb_head_kernel<<<(S + 255) / 256, 256, 0, st>>>(ws.C3, mlp->wc4, mlp->bc4, sample_ids, S, rgb);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}

extern "C" int pnerf_field_backward_f32(const pnerf_points* pts, const pnerf_camera* cam, const pnerf_mlp* mlp,
                                        const pnerf_mode* mode, const float* dirs, const float* sample_loc,
                                        const int* sample_pidx, const int* sample_ids, int S, int SR, int K,
                                        const float* d_sigma, const float* d_rgb, float* g_embed, float* g_color,
                                        float* g_dir, float* g_conf, const pnerf_mlp_grad* gm, void* workspace,
                                        int64_t workspace_bytes, void* stream) {
    if (!pts || !cam || !mlp || !mode || !gm || S < 0 || K <= 0 || K > 32) return PNERF_ERR_ARG;
    if (S == 0) return PNERF_OK;
    if (!workspace || workspace_bytes < pnerf_field_f32_workspace_bytes(S, K)) return PNERF_ERR_WORKSPACE;
    (void)sample_loc;
    cudaStream_t st = (cudaStream_t)stream;
    const Cam c = make_cam(pts, cam);
    Ws ws = carve(workspace, S, K);
    const int64_t M = (int64_t)S * K;
    const float sl = mode->lrelu_slope;
    int rc;
    // colour branch: T0 holds (S,128) / (S,288) gradients
    float* dC3 = ws.T0;
    float* dC2 = ws.T0 + (int64_t)S * HC;
    float* dC1 = ws.T0 + 2 * (int64_t)S * HC;
    float* dC0 = ws.T0 + 3 * (int64_t)S * HC;   // (S, LDC)
    rgb_head_bwd_kernel<<<(S + 63) / 64, 128, 0, st>>>(ws.C3, mlp->wc4, mlp->bc4, sample_ids, S, d_rgb, dC3, gm->wc4, gm->bc4);
    PNERF_LAUNCH_CHECK();
    if ((rc = linear_bwd(dC3, HC, ws.C3, HC, ws.C2, HC, mlp->wc3, gm->wc3, gm->bc3, dC2, HC, S, HC, HC, sl, st))) return rc;
    if ((rc = linear_bwd(dC2, HC, ws.C2, HC, ws.C1, HC, mlp->wc2, gm->wc2, gm->bc2, dC1, HC, S, HC, HC, sl, st))) return rc;
    if ((rc = linear_bwd(dC1, HC, ws.C1, HC, ws.C0, LDC, mlp->wc1, gm->wc1, gm->bc1, dC0, LDC, S, HC, INC, sl, st))) return rc;
    // aggregation: dG into T1 (M,256); per-row d weight into T1 tail
    float* dG = ws.T1;
    float* dw_rows = ws.T1 + M * HID;
    aggregate_bwd_kernel<<<grid_warps(S), 256, 0, st>>>(mlp->wa, mode->density_softplus, sample_ids, S, K, ws, dC0, LDC, d_sigma, dG,
                                                        gm->wa, gm->ba, dw_rows);
    PNERF_LAUNCH_CHECK();
    // mlp_head
    float* dH3 = ws.T0;                  // (M,256)   (colour grads in T0 are dead now)
    if ((rc = linear_bwd(dG, HID, ws.G, HID, ws.H3, HID, mlp->w4, gm->w4, gm->b4, dH3, HID, M, HID, HID, sl, st))) return rc;
    // dw_rows lives behind dG inside T1: copy it out of the way before T1 is reused as dX2
    float* dw_keep = ws.T0 + M * HID;    // T0 has M*288 floats, 32*M spare behind dH3
    PNERF_CUDA(cudaMemcpyAsync(dw_keep, dw_rows, M * 4, cudaMemcpyDeviceToDevice, st));
    float* dX2 = ws.T1;                  // (M,272)
    if ((rc = linear_bwd(dH3, HID, ws.H3, HID, ws.X2, LD2, mlp->w3, gm->w3, gm->b3, dX2, LD2, M, HID, IN3, sl, st))) return rc;
    // mlp_base: dY of layer 2 = first 256 cols of dX2 (stride LD2); Y = X2[:, :256]
    float* dH1 = ws.T0;                  // (M,256) -- dw_keep sits behind it
    if ((rc = linear_bwd(dX2, LD2, ws.X2, LD2, ws.H1, HID, mlp->w2, gm->w2, gm->b2, dH1, HID, M, HID, HID, sl, st))) return rc;
    // dX0 needs (M,288): reuse H1 (dead after this layer's wgrad) is not possible (wgrad reads X0, not H1) -> use H3 + G span
    float* dX0 = ws.H3;                  // H3 (M,256) and G (M,256) are contiguous and both dead now: (M,288) fits
    const bool need_dx0 = g_embed != nullptr;
    if ((rc = linear_bwd(dH1, HID, ws.H1, HID, ws.X0, LD0, mlp->w1, gm->w1, gm->b1, need_dx0 ? dX0 : nullptr, LD0, M, HID, IN1, sl, st))) return rc;
    scatter_kernel<<<grid_warps(M), 256, 0, st>>>(c, sample_pidx, sample_ids, S, K, ws, dX0, dX2, dw_keep, dirs, SR, g_embed, g_color,
                                                  g_dir, g_conf, mode->weight_conf);
    PNERF_LAUNCH_CHECK();
    return PNERF_OK;
}
