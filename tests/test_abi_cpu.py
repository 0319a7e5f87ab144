"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol include/pnerf_b200.h
declares (no compute calls), the host-side mirror keeps the reference's config surface, and the oracle is not
reachable from the product package."""
import ast
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    h = open(os.path.join(ROOT, "include", "pnerf_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(pnerf_[a-z0-9_]+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    from pointnerf2studio_b200 import _lib, build
    build.build()
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), s
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)
    assert lib.pnerf_version() >= 100
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for s in syms:
        assert re.search(rf"\bT {s}\b", nm), f"{s} is not an exported C symbol"


def test_library_is_sm100a_only_and_torch_free():
    from pointnerf2studio_b200 import _lib, build
    build.build()
    ldd = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "c10" not in ldd
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_config_surface_matches_reference():
    from pointnerf2studio_b200 import PointNerfConfig
    c = PointNerfConfig()
    ref_defaults = dict(eval_num_rays_per_chunk=4096, feat_grad=True, conf_grad=True, dir_grad=True, color_grad=True,
                        num_pos_freqs=10, num_viewdir_freqs=4, num_feat_freqs=3, num_dist_freqs=5, agg_dist_pers=20,
                        point_features_dim=32, point_color_mode=True, point_dir_mode=True, num_samples=80,
                        use_biased_sampler=False, field_dim=64, num_mlp_base_layers=2, num_mlp_head_layers=2,
                        num_color_layers=3, num_alpha_layers=1, hidden_size=256, hidden_size_color=128, apply_pnt_mask=True,
                        act_super=False, axis_weight=[1., 1., 1.], kernel_size=[3, 3, 3], vscale=[2, 2, 2],
                        vsize=[0.004] * 3, query_size=[3, 3, 3], ranges=[-1.2] * 3 + [1.2] * 3, z_depth_dim=400, SR=80, K=8,
                        max_o=1000000, P=12, NN=2, gpu_maxthr=1024, zero_epsilon=1e-3, zero_one_loss_weights=0.0001)
    for k, v in ref_defaults.items():
        assert getattr(c, k) == v, k
    with pytest.raises(NotImplementedError):
        PointNerfConfig(hidden_size=128)


def test_plugin_constants():
    from pointnerf2studio_b200 import nerfstudio_plugin as p
    assert p.METHOD_NAME == "pointnerf-original"
    assert p.TRAINER_VALUES["eval_num_rays_per_chunk"] == 2304 and p.TRAINER_VALUES["train_num_rays_per_batch"] == 4096
    assert abs(p.lr_lambda(1000000) - 0.1) < 1e-12 and p.lr_lambda(0) == 1.0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pointnerf2studio_b200")
    for fn in os.listdir(pkg):
        if not fn.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(pkg, fn)).read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n == "oracle" or n.startswith("oracle.") for n in names), fn


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from pointnerf2studio_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.PnerfError):
        _lib.load()


def test_checkpoint_layout_roundtrip(tmp_path):
    import torch
    from pointnerf2studio_b200 import checkpoint
    from pointnerf2studio_b200.synth import make_cloud
    cloud = make_cloud(200, seed=3, radii=(0.03,))
    sd = cloud.state_dict()
    torch.save(sd, tmp_path / "0_net_ray_marching.pth")
    torch.save({"total_steps": 0}, tmp_path / "0_states.pth")
    torch.save(sd, tmp_path / "120_net_ray_marching.pth")
    torch.save({"total_steps": 120}, tmp_path / "120_states.pth")
    assert checkpoint.latest_epoch(str(tmp_path)) == "120"
    back = checkpoint.load_point_cloud_checkpoint(tmp_path)
    assert back["neural_points.points_embeding"].shape == (1, len(cloud.xyz), 32)
    assert back["neural_points.xyz"].shape == (len(cloud.xyz), 3)
    with pytest.raises(RuntimeError):
        checkpoint.load_point_cloud_checkpoint(tmp_path / "missing")


def test_argument_errors_are_return_codes_not_crashes():
    """Every entry validates its arguments before touching the device and reports through the return code (the reference's op
    validates nothing, CPP:51-53): callable on a box without a GPU."""
    import ctypes as C
    from pointnerf2studio_b200 import _lib
    lib = _lib.load()
    ERR_ARG, OK = -1, 0
    assert lib.pnerf_bbox(None, 10, None, None) == ERR_ARG
    assert lib.pnerf_query(None, None, None, 4, 8, 8, 3, C.c_float(0.016), None, None, None, 0, None) == ERR_ARG
    g = _lib.GridView()
    assert lib.pnerf_query(C.byref(g), None, None, 4, 8, 40, 3, C.c_float(0.016), None, None, None, 0, None) == ERR_ARG      # K > 32
    assert lib.pnerf_query(C.byref(g), None, None, 4, 8, 8, 9, C.c_float(0.016), None, None, None, 0, None) == ERR_ARG       # 9^3 kernel
    assert lib.pnerf_query(C.byref(g), None, None, 0, 8, 8, 3, C.c_float(0.016), None, None, None, 0, None) == OK            # no rays
    assert lib.pnerf_sample_select(C.byref(g), None, None, None, None, 0, 4, 400, 8, 1, None, None, None) == ERR_ARG
    assert lib.pnerf_sample_select_jitter(C.byref(g), None, None, C.c_float(6.0), C.c_float(2.0), C.c_float(0.3), 1, 4, 400, 8, 1,
                                          None, None, None) == ERR_ARG                                                     # far <= near
    assert lib.pnerf_composite_forward(None, None, None, None, None, None, 4, 8, None, None, None, None) == ERR_ARG
    cam, mode = _lib.Camera(), _lib.Mode()
    assert lib.pnerf_composite_forward(C.byref(cam), C.byref(mode), None, None, None, None, 4, 200, None, None, None, None) == ERR_ARG  # SR > 128
    assert lib.pnerf_composite_forward(C.byref(cam), C.byref(mode), None, None, None, None, 0, 8, None, None, None, None) == OK
    pts, mlp = _lib.Points(), _lib.Mlp()
    mode.lrelu_slope = 0.1
    assert lib.pnerf_field_forward_tc(C.byref(pts), C.byref(cam), C.byref(mlp), None, C.byref(mode), None, None, None, None, 5, 8, 8,
                                      None, None, None, 0, None) == ERR_ARG                                               # no packed weights
    assert lib.pnerf_adam_step(None, 1, C.c_float(0.9), C.c_float(0.999), C.c_float(1e-8), C.c_float(1.0), None) == ERR_ARG
    assert lib.pnerf_adam_step(None, 0, C.c_float(0.9), C.c_float(0.999), C.c_float(1e-8), C.c_float(1.0), None) == ERR_ARG
    assert lib.pnerf_hit_rays(None, 0, None, None, None, 0, None) == ERR_ARG
    assert lib.pnerf_tc_wpack_bytes() == 565248 + 73728 + 2 * 32768
    assert lib.pnerf_field_tc_train_workspace_bytes(0, 8) > 0 and lib.pnerf_scan_workspace_bytes(1000) > 0


def test_ray_bundle_for_camera_views_and_hint():
    """RayBundle.for_camera: per-camera fields are zero-copy (R,.) views of one small tensor and the host values ride along."""
    import numpy as np
    import torch
    from pointnerf2studio_b200 import RayBundle
    d = torch.rand(7, 3)
    rot = np.arange(9, dtype=np.float32).reshape(3, 3)
    rb = RayBundle.for_camera(d, [1.0, 2.0, 3.0], rot, 2.0, 6.0)
    assert len(rb) == 7 and rb.origins.shape == (7, 3) and rb.nears.shape == (7, 1) and rb.fars.shape == (7, 1)
    assert rb.origins.stride(0) == 0 and rb.nears.stride(0) == 0                 # no R copies
    assert rb.origins[4].tolist() == [1.0, 2.0, 3.0] and float(rb.nears[6]) == 2.0 and float(rb.fars[0]) == 6.0
    assert torch.equal(rb.metadata["camrotc2w"], torch.from_numpy(rot))
    hint = rb.metadata["camera_host"]
    assert hint["near"] == 2.0 and hint["far"] == 6.0 and np.array_equal(hint["camrotc2w"], rot)
    sl = RayBundle(rb.origins[2:5], rb.directions[2:5], rb.nears[2:5], rb.fars[2:5], rb.metadata)   # chunking keeps working
    assert len(sl) == 3 and sl.origins.shape == (3, 3)
