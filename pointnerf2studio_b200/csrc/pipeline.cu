// One-call launchers of a whole training forward / backward pass (rows G0 .. C of SURVEY.md 8a and their gradients): the reference's
// PointNerf.get_outputs (SM:263-399) issues ~150 torch launches and blocks the host five times per call; here the host makes ONE call
// per direction, never learns a data-dependent size (R', S travel as device-side counts) and so never synchronises.  Every kernel is
// launched through the per-stage entry points of this same library; nothing here computes.
#include "pnerf_common.cuh"

using namespace pnerf;

extern "C" int64_t pnerf_render_train_scratch_bytes(int R, int SR) {
    if (R < 0 || SR <= 0) return 0;
    const int64_t slots = (int64_t)R * SR;
    // d_sigma (slots) + d_rgb (slots x 3) for the backward pass + the scan workspace of the two compactions
    return align_up(slots * 16, 256) + align_up(pnerf_scan_workspace_bytes(slots > R ? slots : R), 256) + 256;
}

extern "C" int pnerf_render_train_forward(const pnerf_grid_view* grid, const pnerf_points* pts, const pnerf_camera* cam, const pnerf_mlp* mlp,
                                          const void* wpack, const pnerf_mode* mode, const float* dirs, const float* t_vals, int t_stride,
                                          float near_t, float far_t, float jitter, uint64_t seed, int R, int D, int SR, int K,
                                          int kernel_size0, float radius, int n_samples_cap, int phases, const pnerf_render_buffers* b,
                                          void* stream) {
    if (!grid || !pts || !cam || !mlp || !wpack || !mode || !b || R < 0 || D <= 0 || SR <= 0 || K <= 0) return PNERF_ERR_ARG;
    if (R == 0) return PNERF_OK;
    if (!(phases & 3) || n_samples_cap < 0) return PNERF_ERR_ARG;
    if (!b->sample_loc || !b->sample_cnt || !b->sample_pidx || !b->sample_valid || !b->sample_ids || !b->n_samples || !b->scratch)
        return PNERF_ERR_ARG;
    if ((phases & 2) && (!b->sigma || !b->rgb || !b->out_rgb || !b->ray_mask || !b->ray_index || !b->n_rays || (!b->workspace && n_samples_cap > 0)))
        return PNERF_ERR_ARG;
    if (b->scratch_bytes < pnerf_render_train_scratch_bytes(R, SR)) return PNERF_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t slots = (int64_t)R * SR;
    if (slots > 0x7fffffff / 4) return PNERF_ERR_ARG;
    uint8_t* scan_ws = (uint8_t*)b->scratch + align_up(slots * 16, 256);
    const int64_t scan_bytes = b->scratch_bytes - align_up(slots * 16, 256);
    int rc;
    if (phases & 1) {
        if (t_vals)
            rc = pnerf_sample_select(grid, nullptr, cam->origin, dirs, t_vals, t_stride, R, D, SR, 1, b->sample_loc, b->sample_cnt, stream);
        else if (cam->dev)      // camera / near / far / seed from the device-side step constants (CUDA-graph replay)
            rc = pnerf_sample_select_jitter_dev(grid, cam->dev, dirs, jitter, R, D, SR, 1, b->sample_loc, b->sample_cnt, stream);
        else
            rc = pnerf_sample_select_jitter(grid, cam->origin, dirs, near_t, far_t, jitter, seed, R, D, SR, 1, b->sample_loc, b->sample_cnt, stream);
        if (rc) return rc;
        if ((rc = pnerf_query(grid, b->sample_loc, b->sample_cnt, R, SR, K, kernel_size0, radius, b->sample_pidx, b->sample_valid, nullptr, 0, stream))) return rc;
        if ((rc = pnerf_sample_compact(b->sample_valid, slots, b->sample_ids, b->n_samples, scan_ws, scan_bytes, stream))) return rc;
    }
    if (!(phases & 2)) return PNERF_OK;
    PNERF_CUDA(cudaMemsetAsync(b->sigma, 0, (size_t)slots * 4, st));
    PNERF_CUDA(cudaMemsetAsync(b->rgb, 0, (size_t)slots * 12, st));
    // n_samples_cap: R * SR always works (sync-free); a caller that read the count back passes the count itself
    if ((rc = pnerf_field_forward_tc_train(pts, cam, mlp, wpack, mode, dirs, b->sample_loc, b->sample_pidx, b->sample_ids, n_samples_cap, b->n_samples,
                                           SR, K, b->sigma, b->rgb, b->workspace, b->workspace_bytes, stream))) return rc;
    if ((rc = pnerf_composite_forward(cam, mode, b->sample_loc, b->sample_valid, b->sigma, b->rgb, R, SR, b->out_rgb, nullptr, nullptr, stream))) return rc;
    return pnerf_ray_compact(b->sample_valid, R, SR, b->ray_mask, b->ray_index, b->n_rays, scan_ws, scan_bytes, stream);
}

extern "C" int pnerf_render_train_backward(const pnerf_points* pts, const pnerf_camera* cam, const pnerf_mlp* mlp, const pnerf_mode* mode,
                                           const float* dirs, const float* d_out, int R, int SR, int K, int n_samples_cap,
                                           const pnerf_render_buffers* b,
                                           float* g_embed, float* g_color, float* g_dir, float* g_conf, const pnerf_mlp_grad* g_mlp,
                                           void* points_done_event, void* stream) {
    if (!pts || !cam || !mlp || !mode || !b || !g_mlp || !d_out || R < 0 || SR <= 0 || K <= 0) return PNERF_ERR_ARG;
    if (R == 0) return PNERF_OK;
    if (b->scratch_bytes < pnerf_render_train_scratch_bytes(R, SR)) return PNERF_ERR_WORKSPACE;
    const int64_t slots = (int64_t)R * SR;
    float* d_sigma = (float*)b->scratch;
    float* d_rgb = d_sigma + slots;
    int rc;
    if ((rc = pnerf_composite_backward(cam, mode, b->sample_loc, b->sample_valid, b->sigma, b->rgb, d_out, R, SR, d_sigma, d_rgb, stream))) return rc;
    return pnerf_field_backward_tc(pts, cam, mlp, mode, dirs, b->sample_loc, b->sample_pidx, b->sample_ids, n_samples_cap, b->n_samples, SR, K, d_sigma,
                                   d_rgb, b->rgb, g_embed, g_color, g_dir, g_conf, g_mlp, b->workspace, b->workspace_bytes, points_done_event,
                                   stream);
}
