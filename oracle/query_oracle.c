/* Oracle (plain C) for the voxel-grid neighbour querier.  TEST INFRASTRUCTURE ONLY.
 *
 * Serial restatement of the six kernels of the reference's only native component,
 * /root/reference/pointnerf/models/neural_points/cuda/query_worldcoords.cu ("CU"):
 *   claim_occ CU:18-78, map_coor2occ CU:80-115, fill_occ2pnts CU:117-162,
 *   mask_raypos CU:165-189, get_shadingloc CU:192-214 (+ the torch cumsum at CU:390-391),
 *   query_neigh_along_ray_layered CU:217-302.
 * The reference runs one GPU thread per point / position / sample and resolves collisions
 * with atomics, so voxel ids and bucket order are race order.  Here the loops run in index
 * order, which pins the deterministic rule of DESIGN.md: a voxel keeps its first P points in
 * ascending index, every occupied voxel is kept (no max_o eviction, CU:64-73) and the voxel
 * with id 0 keeps its points (CU:147 drops them).  The replace-the-farthest selection is the
 * reference's own (CU:274-293) fed in the visit order (layer, x, y, z, slot); the K survivors are
 * then emitted sorted by (d2, point index).  d2 is the fmul/ffma/ffma chain nvcc emits.
 *
 * Independent of oracle/grid_query.py (brute-force formulation); tests check they agree.
 * Build: make -C oracle   ->  oracle/libquery_oracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    float lo[3], sv[3];
    int dim[3];
    int P;
    int64_t G;
    int V;              /* occupied voxels */
    int *coor_2_occ;    /* G  : voxel id or -1            */
    int *occ_2_pnts;    /* V*P: point ids, -1 padded      */
    int *occ_numpnts;   /* V  : points seen (may exceed P) */
    uint8_t *coor_occ;  /* G  : dilated occupancy 0/1     */
} qo_grid;

static int vox(const float *p, const qo_grid *g, int c[3]) {
    for (int a = 0; a < 3; a++) {
        float f = floorf((p[a] - g->lo[a]) / g->sv[a]);
        if (!(f >= 0.0f && f < (float)g->dim[a])) return 0;
        c[a] = (int)f;
    }
    return 1;
}
static int64_t lin(const qo_grid *g, int x, int y, int z) {
    return (int64_t)x * g->dim[1] * g->dim[2] + (int64_t)y * g->dim[2] + z;
}

qo_grid *qo_build(const float *xyz, int64_t N, const float lo[3], const float sv[3], const int dim[3],
                  int P, const int query_size[3]) {
    qo_grid *g = (qo_grid *)calloc(1, sizeof(qo_grid));
    memcpy(g->lo, lo, 12); memcpy(g->sv, sv, 12); memcpy(g->dim, dim, 12);
    g->P = P;
    g->G = (int64_t)dim[0] * dim[1] * dim[2];
    g->coor_2_occ = (int *)malloc(sizeof(int) * g->G);
    for (int64_t i = 0; i < g->G; i++) g->coor_2_occ[i] = -1;
    g->coor_occ = (uint8_t *)calloc(g->G, 1);
    /* claim_occ: first point (lowest index) into an empty voxel claims the next id */
    int *occ_2_coor = (int *)malloc(sizeof(int) * 3 * (N > 0 ? N : 1));
    int c[3];
    for (int64_t i = 0; i < N; i++) {
        if (!vox(xyz + 3 * i, g, c)) continue;
        int64_t id = lin(g, c[0], c[1], c[2]);
        if (g->coor_2_occ[id] == -1) {
            g->coor_2_occ[id] = g->V;
            memcpy(occ_2_coor + 3 * g->V, c, 12);
            g->V++;
        }
    }
    /* map_coor2occ: dilate by query_size */
    for (int v = 0; v < g->V; v++) {
        const int *u = occ_2_coor + 3 * v;
        for (int x = (u[0] - query_size[0] / 2 > 0 ? u[0] - query_size[0] / 2 : 0);
             x < (u[0] + (query_size[0] + 1) / 2 < dim[0] ? u[0] + (query_size[0] + 1) / 2 : dim[0]); x++)
            for (int y = (u[1] - query_size[1] / 2 > 0 ? u[1] - query_size[1] / 2 : 0);
                 y < (u[1] + (query_size[1] + 1) / 2 < dim[1] ? u[1] + (query_size[1] + 1) / 2 : dim[1]); y++)
                for (int z = (u[2] - query_size[2] / 2 > 0 ? u[2] - query_size[2] / 2 : 0);
                     z < (u[2] + (query_size[2] + 1) / 2 < dim[2] ? u[2] + (query_size[2] + 1) / 2 : dim[2]); z++)
                    g->coor_occ[lin(g, x, y, z)] = 1;
    }
    free(occ_2_coor);
    /* fill_occ2pnts */
    g->occ_2_pnts = (int *)malloc(sizeof(int) * (size_t)(g->V > 0 ? g->V : 1) * P);
    for (int64_t i = 0; i < (int64_t)g->V * P; i++) g->occ_2_pnts[i] = -1;
    g->occ_numpnts = (int *)calloc(g->V > 0 ? g->V : 1, sizeof(int));
    for (int64_t i = 0; i < N; i++) {
        if (!vox(xyz + 3 * i, g, c)) continue;
        int v = g->coor_2_occ[lin(g, c[0], c[1], c[2])];
        int slot = g->occ_numpnts[v]++;
        if (slot < P) g->occ_2_pnts[(int64_t)v * P + slot] = (int)i;
    }
    return g;
}

void qo_free(qo_grid *g) {
    if (!g) return;
    free(g->coor_2_occ); free(g->occ_2_pnts); free(g->occ_numpnts); free(g->coor_occ); free(g);
}
int qo_num_voxels(const qo_grid *g) { return g->V; }

/* mask_raypos + cumsum + get_shadingloc.  sample_loc (R*SR*3) and sample_mask (R*SR) must be zeroed. */
void qo_select(const qo_grid *g, const float *raypos, int R, int D, int SR,
               float *sample_loc, int *sample_mask, uint8_t *ray_hit) {
    int c[3];
    for (int r = 0; r < R; r++) {
        int n = 0;
        ray_hit[r] = 0;
        for (int j = 0; j < D; j++) {
            const float *p = raypos + ((int64_t)r * D + j) * 3;
            if (!vox(p, g, c) || !g->coor_occ[lin(g, c[0], c[1], c[2])]) continue;
            ray_hit[r] = 1;
            if (n < SR) {
                memcpy(sample_loc + ((int64_t)r * SR + n) * 3, p, 12);
                sample_mask[(int64_t)r * SR + n] = 1;
            }
            n++;
        }
    }
}

static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

/* query_neigh_along_ray_layered.  sample_pidx (S*K) must be filled with -1.
 * stats (optional, 2 per sample): voxel-table entries visited, candidate points examined. */
void qo_query(const qo_grid *g, const float *xyz, const float *sample_loc, const int *sample_mask,
              int64_t S, int K, int kernel0, float radius, int *sample_pidx, int *stats) {
    const float r2 = radius * radius;
    float *d2buf = (float *)malloc(sizeof(float) * K);
    for (int64_t s = 0; s < S; s++) {
        if (sample_mask[s] <= 0) continue;
        const float *q = sample_loc + 3 * s;
        int f[3];
        if (!vox(q, g, f)) continue;   /* cannot happen for a selected sample */
        int *out = sample_pidx + s * K;
        int kid = 0, far_ind = 0, nvis = 0, ncand = 0;
        float far2 = 0.0f;
        for (int layer = 0; layer < (kernel0 + 1) / 2; layer++) {
            for (int x = imax(-f[0], -layer); x < imin(g->dim[0] - f[0], layer + 1); x++)
                for (int y = imax(-f[1], -layer); y < imin(g->dim[1] - f[1], layer + 1); y++)
                    for (int z = imax(-f[2], -layer); z < imin(g->dim[2] - f[2], layer + 1); z++) {
                        if (imax(abs(z), imax(abs(x), abs(y))) != layer) continue;
                        nvis++;
                        int v = g->coor_2_occ[lin(g, f[0] + x, f[1] + y, f[2] + z)];
                        if (v < 0) continue;
                        int n = imin(g->P, g->occ_numpnts[v]);
                        for (int k = 0; k < n; k++) {
                            int p = g->occ_2_pnts[(int64_t)v * g->P + k];
                            ncand++;
                            float dx = xyz[3 * (int64_t)p] - q[0], dy = xyz[3 * (int64_t)p + 1] - q[1],
                                  dz = xyz[3 * (int64_t)p + 2] - q[2];
                            float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                            if (r2 == 0.0f || d2 <= r2) {
                                if (kid++ < K) {
                                    out[kid - 1] = p; d2buf[kid - 1] = d2;
                                    if (d2 > far2) { far2 = d2; far_ind = kid - 1; }
                                } else if (d2 < far2) {
                                    out[far_ind] = p; d2buf[far_ind] = d2; far2 = d2;
                                    for (int i = 0; i < K; i++)
                                        if (d2buf[i] > far2) { far2 = d2buf[i]; far_ind = i; }
                                }
                            }
                        }
                    }
            if (kid >= K) break;
        }
        int n = imin(kid, K);
        for (int i = 1; i < n; i++) {   /* emit sorted by (d2, index) */
            float d = d2buf[i]; int p = out[i], j = i - 1;
            while (j >= 0 && (d2buf[j] > d || (d2buf[j] == d && out[j] > p))) { d2buf[j + 1] = d2buf[j]; out[j + 1] = out[j]; j--; }
            d2buf[j + 1] = d; out[j + 1] = p;
        }
        if (stats) { stats[2 * s] = nvis; stats[2 * s + 1] = ncand; }
    }
    free(d2buf);
}
