"""oracle/cloud_ops.py against the golden vectors produced by executing the reference's own source
(tests/golden/make_golden_cloud_ops.py): probe_hole's candidate filter and construct_vox_points_closest."""
import os

import numpy as np

from oracle import cloud_ops as oc

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cloud_ops_golden.npz"))


def assert_same_argmin(xyz, centroid, got, want, max_tied_frac=2e-3):
    """min_idx must agree except in mathematically tied voxels: a voxel with two points has both exactly equidistant from their
    mean, and which one wins then hangs on the last ulp of the norm (torch.norm's summation order in the reference)."""
    bad = np.nonzero(got != want)[0]
    assert len(bad) <= max_tied_frac * len(want), len(bad)
    for b in bad:
        ra, rb = np.linalg.norm(xyz[got[b]] - centroid[b]), np.linalg.norm(xyz[want[b]] - centroid[b])
        assert abs(ra - rb) <= 1e-6 * max(ra, rb), (b, ra, rb)


def test_probe_filter_matches_the_reference_lines():
    for tag in ("nofar", "far"):
        k = lambda n: G[f"probe_{tag}_{n}"]
        got = oc.probe_filter(k("ray_mask"), k("gt"), k("color"), k("far_dist"), k("opacity"), k("edge"), [1.0, 1.0, 1.0],
                              float(k("far_thresh")), 0.7)
        np.testing.assert_array_equal(got, k("keep"))
        assert got.sum() > 20


def test_vox_closest_matches_the_reference_function():
    for tag in ("a", "b"):
        xyz, res = G[f"vox_{tag}_xyz"], int(G[f"vox_{tag}_res"])
        cen, grid, amin, _, _ = oc.construct_vox_points_closest(xyz, res)
        np.testing.assert_array_equal(grid, G[f"vox_{tag}_grid"])
        np.testing.assert_allclose(cen, G[f"vox_{tag}_centroid"], rtol=0, atol=2e-7)
        assert_same_argmin(xyz, cen, amin, G[f"vox_{tag}_min_idx"])


def test_probe_outputs_match_the_reference_lines():
    """oracle.field.probe (the checker of pnerf_probe, tests/test_gpu_parity.py::test_probe_prune_grow) against the outputs of the
    reference's own lines (models/neural_points_volumetric_model.py:334-355) executed on the same tensors."""
    import torch
    from oracle import field as of
    t = lambda k: torch.from_numpy(G[f"probeout_in_{k}"])[0]
    g = {"mask": torch.ones(t("weight").shape, dtype=torch.bool), "loc_w": t("sample_loc_w"), "xyz": t("sampled_xyz"),
         "color": t("sampled_color"), "dir": t("sampled_dir"), "conf": t("sampled_conf"), "embed": t("sampled_embedding")}
    extras = {"weight_used": t("weight"), "conf_coefficient": t("conf_coefficient")}
    got = of.probe(None, g, extras, t("coarse_point_opacity"))
    for k in ("ray_max_shading_opacity", "ray_max_sample_loc_w", "ray_max_far_dist", "shading_avg_color", "shading_avg_dir",
              "shading_avg_conf", "shading_avg_embedding"):
        want = G[f"probeout_{k}"][0]
        np.testing.assert_allclose(got[k].numpy().reshape(want.shape), want, rtol=1e-6, atol=1e-7, err_msg=k)
