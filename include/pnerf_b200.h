/* pnerf_b200.h -- C ABI of libpnerf_b200.so: Point-NeRF's per-ray hot path on B200 (sm_100a).
 *
 * Drop-in boundary for the reference's only native entry point and for the torch code around it
 * (reference = SHUzhekiNg/pointnerf2studio; paths relative to /root/reference/pointnerf/):
 *   CPP = models/neural_points/cuda/query_worldcoords.cpp   CU = .../query_worldcoords.cu
 *   SU  = nerfstudio/studio_utils.py                        SM = nerfstudio/studio_model.py
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _h (host); nothing here allocates
 *     or frees device memory, the caller (torch) owns every buffer, inputs are never written;
 *   - `stream` is a cudaStream_t passed as void*; every launch goes to it, nothing synchronises
 *     (the reference launches on the legacy default stream and blocks the host five times per call,
 *     CU:310,382,426);
 *   - return value: 0 = PNERF_OK, negative = error (never throws; the reference raises nothing and
 *     checks nothing, CPP:51-53);
 *   - all floats are fp32, all indices int32, B = 1 (the reference always runs with B = 1).
 */
#ifndef PNERF_B200_H
#define PNERF_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNERF_OK 0
#define PNERF_ERR_ARG (-1)      /* bad size / null pointer / unsupported K, P, kernel size */
#define PNERF_ERR_CUDA (-2)     /* a CUDA runtime call failed; see pnerf_last_cuda_error() */
#define PNERF_ERR_WORKSPACE (-3)/* workspace too small */
#define PNERF_ERR_ARCH (-4)     /* device is not sm_100 */

int pnerf_version(void);
const char* pnerf_last_cuda_error(void);
/* 0 when the current device can run the tcgen05 kernels (compute capability 10.x) */
int pnerf_device_check(void);

/* ---------------------------------------------------------------- grid frame (row H)
 * Replaces NeuralPoints.get_hyperparameters' two reductions (SU:116): min/max over the N points.
 * out_minmax: 6 floats (min xyz, max xyz).  The fp64 dim arithmetic of SU:125-126 stays on the host. */
int pnerf_bbox(const float* xyz, int64_t n, float* out_minmax, void* stream);

/* ---------------------------------------------------------------- grid build (row G1)
 * Replaces claim_occ / map_coor2occ / fill_occ2pnts (CU:18-162, launched CU:322-365) and the
 * per-call allocations of CU:314-319,337.  Built once per point-cloud version, not per call.
 * Deterministic: a voxel keeps its first P points in ascending index; all occupied voxels kept.
 *   lo_h, sv_h, dim_h : grid origin, scaled voxel size, dims (host, 3 each)  [ranges, scaled_vsize, scaled_vdim]
 *   query_size_h      : dilation box of the occupancy (CU:342 passes query_size)
 * outputs
 *   cell_start [G+1]  : CSR offsets into `recs`, cell id = x*dimy*dimz + y*dimz + z (CU:45)
 *   recs [n] float4   : (x, y, z, bits(point index | (vz & 7) << 28)) of kept points, sorted by
 *                       (cell, index); only the first cell_start[G] entries are valid
 *   occ_bits [ceil(G/32)] : dilated occupancy, bit (id & 31) of word (id >> 5)   [coor_occ]
 * workspace: at least pnerf_grid_workspace_bytes(n, G) bytes. */
int64_t pnerf_grid_workspace_bytes(int64_t n, int64_t cells);
int pnerf_grid_build(const float* xyz, int64_t n, const float* lo_h, const float* sv_h, const int* dim_h,
                     int P, const int* query_size_h, int* cell_start, float* recs, uint32_t* occ_bits,
                     void* workspace, int64_t workspace_bytes, void* stream);

typedef struct {
    float lo[3];
    float sv[3];
    int dim[3];
    const int* cell_start;
    const float* recs;
    const uint32_t* occ_bits;
} pnerf_grid_view;

/* ---------------------------------------------------------------- sample selection (rows G0, G2)
 * Replaces mask_raypos, the torch max/sum/cumsum/masked_select chain and get_shadingloc
 * (CU:165-214, 368-402).  One of three position sources:
 *   raypos != NULL              : explicit coarse positions (R,D,3)  -- the reference's own input
 *   t_vals != NULL, t_stride=0  : pos = origin + dir * t_vals[j]  (one t table for all rays)
 *   t_vals != NULL, t_stride=D  : per-ray t table (R,D)
 * (mul then add, separately rounded, as torch evaluates RM:330).
 * outputs, not compacted over rays:
 *   sample_loc (R,SR,3) zero where empty; sample_cnt (R) = min(#hits, SR); 0 <=> ray not in R'.
 *   fill_missed = 0 leaves the rows of rays without any hit unwritten (the caller compacts the hit rays with
 *   pnerf_hit_rays / pnerf_gather_hit_rays and never reads them: 5 of 6 rays of an 800x800 object view). */
int pnerf_sample_select(const pnerf_grid_view* grid_h, const float* raypos, const float* origin_h,
                        const float* dirs, const float* t_vals, int t_stride, int R, int D, int SR, int fill_missed,
                        float* sample_loc, int* sample_cnt, void* stream);
/* Same, with the coarse t mid-points of near_far_linear_ray_generation (RM:312-329, called with jitter 0.3 at
 * SU:166) generated in registers: segment j = (edge_{j+1} - edge_j) * (1 + jitter * (U - 0.5)), U = Philox4x32-10
 * keyed by `seed` with counter (j/4, ray) -- the reference draws U with torch.rand, so jittered runs agree with it
 * in distribution, not bit for bit; pnerf_coarse_t writes the very t (R,D) and U (R,D, optional) this kernel uses so
 * that a checker can replay them through the t-table source above. */
int pnerf_sample_select_jitter(const pnerf_grid_view* grid_h, const float* origin_h, const float* dirs, float near_t,
                               float far_t, float jitter, uint64_t seed, int R, int D, int SR, int fill_missed,
                               float* sample_loc, int* sample_cnt, void* stream);
/* Same with origin / near / far / seed read from the device-side step constants (pnerf_camera.dev layout) at run time. */
int pnerf_sample_select_jitter_dev(const pnerf_grid_view* grid_h, const float* step_dev, const float* dirs, float jitter, int R, int D,
                                   int SR, int fill_missed, float* sample_loc, int* sample_cnt, void* stream);
/* Hit-ray compaction (R -> R' of the reference's op, CU:381-391): ids of the rays with sample_cnt > 0, ascending, and their
 * rows of sample_loc / sample_cnt / dirs gathered into compact (R',.) arrays; every later stage then runs on R' rays. */
int pnerf_hit_rays(const int* sample_cnt, int R, int* ray_index, int* n_rays, void* workspace, int64_t workspace_bytes,
                   void* stream);
int pnerf_gather_hit_rays(const int* ray_index, int n_rays, int SR, const float* sample_loc, const int* sample_cnt,
                          const float* dirs, float* loc_out, int* cnt_out, float* dirs_out, void* stream);
int pnerf_coarse_t(float near_t, float far_t, float jitter, uint64_t seed, int R, int D, float* t_out, float* u_out,
                   void* stream);

/* ---------------------------------------------------------------- neighbour query (row Q)
 * Replaces query_neigh_along_ray_layered (CU:217-302): layer-truncated, bucket-capped,
 * radius-limited K nearest.  K <= 32; layers = (kernel_size0+1)/2 <= 3.  Tie-break rule: candidates
 * are visited in (layer, ux, uy, uz, point index) order, the K kept are the K smallest by
 * (d2, visit order) and are emitted in that order; d2 = fma(dz,dz,fma(dy,dy,dx*dx)).
 * outputs: sample_pidx (R,SR,K) with -1 padding (every slot is written);
 *          sample_valid (R*SR) uint8: the number of neighbours found for the slot (0 = none; every consumer tests it for != 0);
 *          stats (optional, 2 x uint64): voxel-table entries visited, candidate points examined;
 *          rays_are_neighbours: a scheduling hint with no effect on the result -- non-zero when consecutive rays are neighbouring
 *          pixels of one image (the hit-ray list of a render): a warp then takes the same slot of 32 consecutive rays instead of
 *          32 consecutive slots of one ray (candidate streams of similar length, better balanced lock-step). */
int pnerf_query(const pnerf_grid_view* grid_h, const float* sample_loc, const int* sample_cnt, int R, int SR,
                int K, int kernel_size0, float radius, int* sample_pidx, uint8_t* sample_valid,
                unsigned long long* stats, int rays_are_neighbours, void* stream);

/* Ray compaction of the reference's return value (CU:425-432): ray_mask (R) int8 and, when the
 * caller wants the reference's compact (R'',SR,.) tensors, an index list of the surviving rays.
 *   ray_index [R] : ids of rays with >= 1 neighbour, ascending;  n_rays [1] : R''. */
int pnerf_ray_compact(const uint8_t* sample_valid, int R, int SR, int8_t* ray_mask, int* ray_index,
                      int* n_rays, void* workspace, int64_t workspace_bytes, void* stream);
int pnerf_gather_rays(const int* ray_index, int n_rays, int SR, int K, const int* sample_pidx,
                      const float* sample_loc, int* out_pidx, float* out_loc, void* stream);

/* Compact list of valid samples (ascending slot id r*SR+s): sample_ids [R*SR], n_samples [1]. */
int pnerf_sample_compact(const uint8_t* sample_valid, int64_t n_slots, int* sample_ids, int* n_samples,
                         void* workspace, int64_t workspace_bytes, void* stream);
int64_t pnerf_scan_workspace_bytes(int64_t n);
/* The same list bucketed by neighbour count, for the tensor-core field kernels: class c (class_kp_h strictly descending powers of
 * two, class_kp_h[0] >= K, e.g. {8, 4, 2}) holds the slots whose number n of valid neighbours satisfies
 * class_kp_h[c+1] < n <= class_kp_h[c] (the last class: 0 < n); sample_ids = the classes back to back, n_per_class [n_classes]
 * (device) their sizes, each class ascending.  A class-c sample then occupies class_kp_h[c] MMA rows instead of K.
 * workspace: pnerf_scan_workspace_bytes(n_slots) bytes is enough. */
int pnerf_sample_compact_classes(const uint8_t* sample_count /* pnerf_query's sample_valid */, int64_t n_slots, int K, int n_classes, const int* class_kp_h,
                                 int* sample_ids, int* n_per_class, void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- field networks (rows P, GA, W, E, M1, A, M2)
 * Replaces NeuralPoints.forward's gather (SU:190-209) and PointNerf.get_outputs' field math
 * (SM:270-366): perspective coords, dists6, inverse-distance weights, positional encodings,
 * mlp_base, mlp_head, density head, K-aggregation, mlp_color, rgb head. */
typedef struct {
    const float* xyz;     /* (N,3)  points_xyz      */
    const float* embed;   /* (N,32) points_embeding */
    const float* color;   /* (N,3)  points_color    */
    const float* dir;     /* (N,3)  points_dir      */
    const float* conf;    /* (N,1)  points_conf     */
    float Rw2c[9];        /* points_Rw2c, row major (host copy) */
    int64_t n;
} pnerf_points;

typedef struct {
    float origin[3];      /* ray_bundle.origins[0]             (SU:152) */
    float R_c2w[9];       /* metadata["camrotc2w"], row major  (SU:148-151) */
    /* Optional DEVICE copy of the per-step constants, PNERF_STEP_WORDS 32-bit words:
     *   [0..2] origin, [3..11] R_c2w (row major), [12] near, [13] far, [14] jitter seed low word, [15] high word (as bits).
     * When non-NULL the kernels read the camera (and pnerf_sample_select_jitter the near / far / seed) from it at RUN time and
     * ignore the host values above: a CUDA graph captured once is replayed for every camera / step by rewriting these 64 bytes. */
    const float* dev;
} pnerf_camera;
#define PNERF_STEP_WORDS 16

typedef struct {          /* fp32 weights, torch nn.Linear layout (out,in) row major */
    const float *w1, *b1; /* mlp_base.layers.0  256x284  [aggregator.block1.0]       */
    const float *w2, *b2; /* mlp_base.layers.1  256x256  [aggregator.block1.2]       */
    const float *w3, *b3; /* mlp_head.layers.0  256x263  [aggregator.block3.0]       */
    const float *w4, *b4; /* mlp_head.layers.1  256x256  [aggregator.block3.2]       */
    const float *wa, *ba; /* field_output_density.net 1x256 [aggregator.alpha_branch.0] */
    const float *wc1, *bc1; /* mlp_color.layers.0 128x280 [aggregator.color_branch.0] */
    const float *wc2, *bc2; /* mlp_color.layers.1 128x128 [aggregator.color_branch.2] */
    const float *wc3, *bc3; /* mlp_color.layers.2 128x128 [aggregator.color_branch.4] */
    const float *wc4, *bc4; /* field_output_color.net 3x128 [aggregator.color_branch.6] */
} pnerf_mlp;

typedef struct {
    float lrelu_slope;    /* 0.1 plugin (SM:197), 0.01 original flow (PA:286 default slope)      */
    int density_softplus; /* 0: ReLU (SM:221);  1: Softplus(raw-1) (PA:260-265)                   */
    int weight_conf;      /* 0: plugin (SM:318); 1: weight * clamp(conf,1e-4,1) (PA:822-826)      */
    int bg_mode;          /* 0: C += bg*(1-sum w) (SM:387-390); 1: C += bg*T_end (RM:529-532)     */
    int eval_clamp;       /* 1: nan_to_num + clamp [0,1] (nerfstudio RGBRenderer in eval)         */
    float bg[3];
    float vsize_z;        /* config.vsize[2], unscaled (SM:369-374)                               */
} pnerf_mode;

/* fp32 SIMT implementation (exact-parity path).  Row = (valid sample, neighbour slot), M = S*K rows
 * incl. masked slots (zero weight).  Saves every activation for the backward pass.
 * Workspace layout is private; query its size with pnerf_field_f32_workspace_bytes(S, K). */
int64_t pnerf_field_f32_workspace_bytes(int64_t n_samples, int K);
int pnerf_field_forward_f32(const pnerf_points* pts_h, const pnerf_camera* cam_h, const pnerf_mlp* mlp_h,
                            const pnerf_mode* mode_h, const float* dirs, const float* sample_loc,
                            const int* sample_pidx, const int* sample_ids, int n_samples, int SR, int K,
                            float* sigma /* (R*SR) by slot, 0 elsewhere: caller zero-fills */,
                            float* rgb /* (R*SR,3) by slot */, void* workspace, int64_t workspace_bytes,
                            void* stream);
typedef struct {
    float *w1, *b1, *w2, *b2, *w3, *b3, *w4, *b4, *wa, *ba, *wc1, *bc1, *wc2, *bc2, *wc3, *bc3, *wc4, *bc4;
} pnerf_mlp_grad;
/* Backward of the above: consumes d sigma / d rgb (by slot), accumulates (+=) into the point
 * gradients (N,32),(N,3),(N,3),(N,1) (may be NULL) and the MLP gradients. */
int pnerf_field_backward_f32(const pnerf_points* pts_h, const pnerf_camera* cam_h, const pnerf_mlp* mlp_h,
                             const pnerf_mode* mode_h, const float* dirs, const float* sample_loc,
                             const int* sample_pidx, const int* sample_ids, int n_samples, int SR, int K,
                             const float* d_sigma, const float* d_rgb, float* g_embed, float* g_color,
                             float* g_dir, float* g_conf, const pnerf_mlp_grad* g_mlp_h, void* workspace,
                             int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- field networks, tensor-core path (same rows)
 * bf16 operands / fp32 accumulation on tcgen05 + TMEM: gather, encodings, mlp_base, mlp_head, density head and
 * K-aggregation fused in one persistent kernel (no encoded input or activation ever reaches HBM), mlp_color +
 * rgb head in a second one.  `wpack` is the bf16 K-slab copy of the seven weight matrices made by
 * pnerf_tc_pack_weights (pnerf_tc_wpack_bytes() bytes; re-pack after every optimiser step).
 * Saves nothing (inference); sigma / rgb as in pnerf_field_forward_f32.  Training: pnerf_field_forward_tc_train below.
 * workspace: the aggregated features between the two kernels, 512 B per sample in whole 128-sample tiles
 * (pnerf_field_tc_workspace_bytes(n_samples) bytes, 256-byte aligned); a caller bounds it by calling with pieces of the sample list
 * (sample_ids + offset, n_samples of the piece) -- samples are independent. */
int64_t pnerf_tc_wpack_bytes(void);
int pnerf_tc_pack_weights(const pnerf_mlp* mlp_h, void* wpack, void* stream);
int64_t pnerf_field_tc_workspace_bytes(int64_t n_samples);
/* The two kernels of pnerf_field_forward_tc on their own, for sample lists bucketed by neighbour count
 * (pnerf_sample_compact_classes): the per-neighbour networks of `n_samples` samples with `rows_per_sample` (2, 4, 8, 16 or 32,
 * >= every listed sample's neighbour count) MMA rows each, writing the aggregated features of sample i to position
 * first_sample + i of `workspace` (512 B per sample, 128-sample tiles); then ONE colour-network launch over the whole list. */
int pnerf_field_forward_tc_part(const pnerf_points* pts_h, const pnerf_camera* cam_h, const pnerf_mlp* mlp_h, const void* wpack,
                                const pnerf_mode* mode_h, const float* dirs, const float* sample_loc, const int* sample_pidx,
                                const int* sample_ids, int n_samples, int rows_per_sample, int first_sample, int SR, int K,
                                float* sigma, void* workspace, int64_t workspace_bytes, void* stream);
int pnerf_color_forward_tc(const pnerf_points* pts_h, const pnerf_camera* cam_h, const pnerf_mlp* mlp_h, const void* wpack,
                           const pnerf_mode* mode_h, const float* dirs, const int* sample_ids, int n_samples, int SR, float* rgb,
                           const void* workspace, int64_t workspace_bytes, void* stream);
int pnerf_field_forward_tc(const pnerf_points* pts_h, const pnerf_camera* cam_h, const pnerf_mlp* mlp_h, const void* wpack,
                           const pnerf_mode* mode_h, const float* dirs, const float* sample_loc, const int* sample_pidx,
                           const int* sample_ids, int n_samples, int SR, int K, float* sigma, float* rgb, void* workspace,
                           int64_t workspace_bytes, void* stream);

/* Training on the tensor-core path.  Forward = pnerf_field_forward_tc for the per-neighbour networks with every MMA operand kept
 * in `workspace` (2.7 KB per neighbour row + 1.3 KB per sample, tile layout of csrc/tc_layout.cuh).  Backward (what torch
 * autograd computes for the reference through SU:190-209, SM:270-359): consumes d sigma / d rgb (by slot) and the SAME workspace,
 * accumulates (+=) into the point gradients (may be NULL) and the MLP gradients: bf16 tcgen05 GEMMs for dgrad and wgrad of
 * mlp_base / mlp_head / mlp_color, fp32 accumulation.  `workspace` must be 256-byte aligned. */
int64_t pnerf_field_tc_train_workspace_bytes(int64_t n_samples, int K);
/* n_samples_dev (optional, device): the number of valid samples when the HOST does not know it (pnerf_sample_compact wrote it and
 * nobody read it back).  `n_samples` is then a capacity (<= R*SR always works): workspace and grids are sized from it and every
 * kernel clamps its tile loop to min(*n_samples_dev, n_samples) -- a training step never synchronises the host.
 * points_done_event (optional, cudaEvent_t): recorded on `stream` right after the point gradients (g_embed .. g_conf) are complete
 * and before the weight-gradient GEMMs, so that a data-parallel caller can start reducing them under the rest of the backward. */
int pnerf_field_forward_tc_train(const pnerf_points* pts_h, const pnerf_camera* cam_h, const pnerf_mlp* mlp_h, const void* wpack,
                                 const pnerf_mode* mode_h, const float* dirs, const float* sample_loc, const int* sample_pidx,
                                 const int* sample_ids, int n_samples, const int* n_samples_dev, int SR, int K, float* sigma,
                                 float* rgb, void* workspace, int64_t workspace_bytes, void* stream);
int pnerf_field_backward_tc(const pnerf_points* pts_h, const pnerf_camera* cam_h, const pnerf_mlp* mlp_h, const pnerf_mode* mode_h,
                            const float* dirs, const float* sample_loc, const int* sample_pidx, const int* sample_ids, int n_samples,
                            const int* n_samples_dev, int SR, int K, const float* d_sigma, const float* d_rgb,
                            const float* rgb /* forward output, by slot */, float* g_embed, float* g_color, float* g_dir, float* g_conf,
                            const pnerf_mlp_grad* g_mlp_h, void* workspace, int64_t workspace_bytes, void* points_done_event,
                            void* stream);

/* ---------------------------------------------------------------- one-call training pass (rows G0 .. C and their gradients)
 * What PointNerf.get_outputs (SM:263-399) and torch autograd do per training step, as ONE host call per direction: sample
 * selection (in-kernel jitter, or the (D,) / (R,D) table t_vals) -> neighbour query -> sample compaction -> tensor-core field
 * networks (operands kept) -> compositing -> ray mask.  The host never learns R'' or S (device-side counts), so nothing
 * synchronises.  All buffers are the caller's (torch allocates):
 *   sample_loc (R,SR,3) f32, sample_cnt (R) i32, sample_pidx (R,SR,K) i32, sample_valid (R,SR) u8, sample_ids (R*SR) i32,
 *   n_samples (1) i32, sigma (R,SR) f32, rgb (R,SR,3) f32 [both zero-filled here], out_rgb (R,3) f32, ray_mask (R) i8,
 *   ray_index (R) i32, n_rays (1) i32; workspace: pnerf_field_tc_train_workspace_bytes(R*SR, K) bytes, 256-byte aligned, kept
 *   until the backward call; scratch: pnerf_render_train_scratch_bytes(R, SR) bytes, may be reused right after each call.
 * n_samples_cap: capacity of the sample workspace (pnerf_field_tc_train_workspace_bytes(n_samples_cap, K)); R * SR never overflows.
 * phases: bit 0 = selection + query + sample compaction, bit 1 = field networks + compositing + ray mask; a caller that cannot
 *   afford the worst-case workspace runs phase 1, reads n_samples back, and runs phase 2 with n_samples_cap = that count.
 * backward: d_out (R,3) -> += into the point gradients (may be NULL) and the MLP gradients (as pnerf_field_backward_tc). */
typedef struct {
    float* sample_loc; int* sample_cnt; int* sample_pidx; uint8_t* sample_valid; int* sample_ids; int* n_samples;
    float* sigma; float* rgb; float* out_rgb; int8_t* ray_mask; int* ray_index; int* n_rays;
    void* workspace; int64_t workspace_bytes; void* scratch; int64_t scratch_bytes;
} pnerf_render_buffers;
int64_t pnerf_render_train_scratch_bytes(int R, int SR);
int pnerf_render_train_forward(const pnerf_grid_view* grid_h, const pnerf_points* pts_h, const pnerf_camera* cam_h, const pnerf_mlp* mlp_h,
                               const void* wpack, const pnerf_mode* mode_h, const float* dirs, const float* t_vals, int t_stride,
                               float near_t, float far_t, float jitter, uint64_t seed, int R, int D, int SR, int K, int kernel_size0,
                               float radius, int n_samples_cap, int phases, const pnerf_render_buffers* buf_h, void* stream);
int pnerf_render_train_backward(const pnerf_points* pts_h, const pnerf_camera* cam_h, const pnerf_mlp* mlp_h, const pnerf_mode* mode_h,
                                const float* dirs, const float* d_out, int R, int SR, int K, int n_samples_cap,
                                const pnerf_render_buffers* buf_h,
                                float* g_embed, float* g_color, float* g_dir, float* g_conf, const pnerf_mlp_grad* g_mlp_h,
                                void* points_done_event, void* stream);

/* Profiling hook: when `buf` (device, pnerf_tc_trace_bytes() bytes, zero-filled) is set, CTA 0 of the next field_tc launches
 * appends clock64-stamped pipeline events per role warp (tools/tc_trace.py decodes them).  NULL switches it off. */
int pnerf_tc_set_trace(void* buf);
int64_t pnerf_tc_trace_bytes(void);

/* ---------------------------------------------------------------- step length + compositing (rows D, C, F)
 * Replaces SM:368-390 (+ nerfstudio RGBRenderer) and fill_invalid SM:491-504; original-flow twin
 * NPV:271-279 + ray_march RM:495-541.  One warp per ray, all R rays (missed rays -> bg).
 *   out_rgb (R,3);  out_weights (R,SR) optional blend weights;  out_T_end (R) optional. */
int pnerf_composite_forward(const pnerf_camera* cam_h, const pnerf_mode* mode_h, const float* sample_loc,
                            const uint8_t* sample_valid, const float* sigma, const float* rgb, int R, int SR,
                            float* out_rgb, float* out_weights, float* out_T_end, void* stream);
/* d_out (R,3) -> d_sigma (R*SR), d_rgb (R*SR,3), every slot written (0 where invalid). */
int pnerf_composite_backward(const pnerf_camera* cam_h, const pnerf_mode* mode_h, const float* sample_loc,
                             const uint8_t* sample_valid, const float* sigma, const float* rgb,
                             const float* d_out, int R, int SR, float* d_sigma, float* d_rgb, void* stream);

/* Hole probing of the original flow (models/neural_points_volumetric_model.py:331-362): per ray the sample of
 * largest opacity (first on ties), its world position, the distance to its nearest gathered neighbour (invalid slots gather point 0)
 * and the (weight * confidence)-averaged colour / dir / conf / embedding of its K neighbours.  Outputs by ray: (R), (R,3), (R),
 * (R,3), (R,3), (R), (R,32).  sigma as produced by the field kernels (by slot). */
int pnerf_probe(const pnerf_points* pts_h, const pnerf_camera* cam_h, const pnerf_mode* mode_h, const float* sample_loc,
                const uint8_t* sample_valid, const float* sigma, const int* sample_pidx, int R, int SR, int K, float* max_opacity,
                float* max_loc, float* far_dist, float* avg_color, float* avg_dir, float* avg_conf, float* avg_embed, void* stream);

/* Which probed rays become new neural points (the tensor code of probe_hole, run/train_studio.py:414-423 with
 * bloat_inds :447-455): over the H x W maps a probe pass produced (pnerf_probe outputs scattered to pixels),
 *   keep[p] = ray_mask[p] > 0  and  opacity[p] > opacity_thresh  and
 *             ( some pixel q of p's 3x3 window, clipped to the image, has edge_mask[q] and ray_mask[q] < 1 and |gt[q] - bg| > 0.002
 *               or (far_thresh > 0 and far_dist[p] > far_thresh and |gt[p] - color[p]| < 0.1) ).
 * edge_mask may be NULL (every pixel belongs to the frame); color / far_dist may be NULL when far_thresh <= 0.  bg_h: 3 host floats. */
int pnerf_probe_filter(const int8_t* ray_mask, const float* gt, const float* color, const float* far_dist, const float* opacity,
                       const uint8_t* edge_mask, const float* bg_h, int H, int W, float far_thresh, float opacity_thresh,
                       uint8_t* keep, void* stream);

/* Neural-point initialisation, voxel down-sample (construct_vox_points_closest, models/mvs/mvs_utils.py:537-561,
 * called from run/gen_pnts.py): voxel of a point = floor((xyz - space_min) / vox_size) (fp32); for every occupied voxel, in
 * (x, y, z) lexicographic order (torch.unique's), its centroid (mean of its points), its integer coordinates and the index of the
 * point closest to the centroid (lowest index on ties, scatter_min's first-minimum rule).  dim_h: voxels per axis; points outside
 * [0, dim) are counted in n_outside and ignored.  Outputs hold at most max_out voxels (n_out[0] = the true count).
 * workspace: pnerf_vox_closest_workspace_bytes(dim_h) bytes (a dense counting grid: 40 B per voxel of the frame). */
int64_t pnerf_vox_closest_workspace_bytes(const int* dim_h);
int pnerf_vox_closest(const float* xyz, int64_t n, const float* space_min_h, const float* vox_size_h, const int* dim_h, int max_out,
                      float* centroid, int* grid_idx, int* min_idx, int* n_out, int* n_outside, void* workspace,
                      int64_t workspace_bytes, void* stream);

/* Masked-ray MSE of get_loss_dict (studio_model.py:415-426): loss_out[0] = sum_{ray_mask > 0} |pred - image|^2 / (3 * #masked) + 1e-6.
 * acc: 3 floats of workspace, ZERO on entry, kept for the backward call (acc[1] = #masked rays).
 * backward: g_pred (R,3) = d_loss[0] * 2 (pred - image) / (3 * #masked) on masked rays, 0 elsewhere; d_loss is a device scalar. */
int pnerf_masked_mse_forward(const float* pred, const float* image, const int8_t* ray_mask, int R, float* acc, float* loss_out,
                             void* stream);
int pnerf_masked_mse_backward(const float* pred, const float* image, const int8_t* ray_mask, int R, const float* acc,
                              const float* d_loss, float* g_pred, void* stream);

/* Confidence ("zero-one") loss term of SM:288-292,427-429 over ALL R''*SR*K slots (invalid slots
 * read point 0, SU:194): adds the value to loss_out[0] and its gradient to g_conf (N). */
int pnerf_conf_loss(const float* conf, const int* sample_pidx, const int8_t* ray_mask, int R, int SR, int K,
                    float eps, float weight, const int* n_rays /* device R'' */, float* loss_out,
                    float* g_conf, float grad_scale, void* stream);

/* ---------------------------------------------------------------- sharded image read-back (render partitioning)
 * When ONE image is split over the ranks by interleaved rows, every rank copies the rows it rendered straight into one
 * image in shared host memory (page-locked in each process with pnerf_host_register): row i of `src` (device, contiguous
 * rows of row_bytes) lands at dst_h + i * dst_pitch.  One strided copy per rank, no collective and no staging buffer
 * (the reference has no sharded render; its eval loop renders 2304-ray chunks on one GPU, studio_config.py:25). */
int pnerf_host_register(void* host_ptr, int64_t bytes);
int pnerf_host_unregister(void* host_ptr);
int pnerf_copy_rows_to_host(void* dst_h, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t row_bytes,
                            int64_t n_rows, void* stream);

/* ---------------------------------------------------------------- tensor-core plumbing self-test
 * D[128xN] = A[128xK] * W[NxK]^T (bf16 in, fp32 out) through tcgen05.mma / TMEM / bulk async copy.
 * A: bf16 row major; Wp: bf16 in the K-slab layout [K/8][N][8] (see csrc/umma.cuh). */
int pnerf_umma_selftest(const void* A, const void* Wp, float* D, int N, int K, void* stream);

/* ---------------------------------------------------------------- optimiser step
 * torch.optim.Adam (betas, eps, bias correction; no weight decay, no amsgrad -- what nerfstudio's AdamOptimizerConfig builds for
 * the plugin's two groups, studio_config.py:33-48) for up to PNERF_ADAM_MAX_SEGS tensors in ONE launch: p, m, v updated in place
 * from g * grad_scale.  `step` is the tensor's own 1-based step count (torch keeps it per parameter: a parameter without a gradient
 * is skipped and does not advance); lr is per tensor (the host applies the schedule). */
#define PNERF_ADAM_MAX_SEGS 32
typedef struct { float* p; const float* g; float* m; float* v; int64_t n; int64_t step; float lr; } pnerf_adam_seg;
int pnerf_adam_step(const pnerf_adam_seg* segs_h, int n_segs, float beta1, float beta2, float eps, float grad_scale, void* stream);

/* Data-parallel optimiser step (training partitioning): gradient averaging over the ranks + Adam + parameter
 * broadcast in ONE kernel over peer-mapped memory, replacing DDP's all-reduce (studio_pipeline.py:48-53) followed by a
 * replicated torch.optim.Adam over both groups (studio_config.py:33-48).  Every rank keeps ALL its trainable parameters in one
 * flat fp32 buffer and all gradients in another (same layout on every rank: neural-point tensors first, then the MLP tensors),
 * both mapped into every peer (CUDA IPC / symmetric memory: p[w], g[w] = rank w's buffers as seen from this process).  This
 * rank owns elements [lo, hi) (multiples of 4): it sums g[*][lo:hi), scales by grad_scale (1 / world), applies Adam with its
 * slice of the moments m, v (hi - lo elements each: optimiser state is sharded) and writes the new values to p[*][lo:hi).
 * Elements below `boundary` use lr[0] (neural points), the others lr[1] (fields).  world = 1: a plain fused Adam over one flat
 * buffer.  The caller orders the launch between two cross-rank barriers (all gradients complete / all parameters delivered). */
#define PNERF_DP_MAX_RANKS 8
typedef struct {
    float* p[PNERF_DP_MAX_RANKS];
    const float* g[PNERF_DP_MAX_RANKS];
    float* m; float* v;
    int64_t lo, hi, boundary, step;
    float lr[2];
    int world, rank;
    /* optional DEVICE array of 3 floats {lr[0] / bias_correction1, lr[1] / bias_correction1, 1 / sqrt(bias_correction2)} read at
     * run time instead of the values derived from lr / step above (CUDA-graph replay with a moving schedule) */
    const float* hyper_dev;
    /* optional NVLS multicast addresses of the SAME gradient / parameter buffers (one virtual address that reaches every rank's
     * copy through the NVSwitch): the gradient sum then is one multimem.ld_reduce per 16 bytes (reduced inside the switch) and the
     * parameter broadcast one multimem.st -- a GPU's links carry 2/world of the flat size instead of 2 (world-1)/world.  Both or none. */
    const float* mc_g; float* mc_p;
} pnerf_dp_adam;
int pnerf_dp_adam_step(const pnerf_dp_adam* h, float beta1, float beta2, float eps, float grad_scale, void* stream);

/* Micro-benchmarks of the resources the tensor-core kernels lean on (one CTA per SM, all SMs): which = 0 tcgen05.mma
 * rate (param = N), 1 L2 -> shared bulk-copy ring (param = chunk bytes, src >= 557056 bytes), 2 tcgen05.ld rate
 * (param = warps).  out[148] = cycles for `iters` operations per CTA. */
int pnerf_tc_microbench(int which, int iters, int param, const void* src, unsigned long long* out, float* sink, void* stream);

#ifdef __cplusplus
}
#endif
#endif
