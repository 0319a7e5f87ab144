"""Import the reference's own Python modules for the hot path (CPU), with stub modules for the
third-party imports they do not need on this path.  TEST INFRASTRUCTURE ONLY; works only where
/root/reference exists (the authoring container), and is used only by tests/golden/make_golden.py.

Stubs: nerfstudio.utils.printing (NW:8), nerfstudio.field_components.encodings
(PA:8), utils.spherical (PA:6; needs scipy.special.sph_harm which this scipy no longer has).
"""
import argparse
import importlib
import sys
import types

REF_ROOT = "/root/reference/pointnerf"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_reference():
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if "nerfstudio" not in sys.modules:
        ns = _stub("nerfstudio"); ns.__path__ = []
        u = _stub("nerfstudio.utils"); u.__path__ = []
        _stub("nerfstudio.utils.printing", print_tcnn_speed_warning=lambda *a, **k: None)
        fc = _stub("nerfstudio.field_components"); fc.__path__ = []
        _stub("nerfstudio.field_components.encodings", NeRFEncoding=object)
    import utils as ref_utils  # the reference's own `utils` package
    assert any(str(q).startswith(REF_ROOT) for q in ref_utils.__path__), list(ref_utils.__path__)
    _stub("utils.spherical", SphericalHarm_table=object, SphericalHarm=object)
    mods = types.SimpleNamespace()
    mods.point_aggregators = importlib.import_module("models.aggregators.point_aggregators")
    mods.diff_ray_marching = importlib.import_module("models.rendering.diff_ray_marching")
    mods.diff_render_func = importlib.import_module("models.rendering.diff_render_func")
    mods.networks = importlib.import_module("models.helpers.networks")
    return mods


def chair_opt():
    """argparse.Namespace with the values of dev_scripts/w_n360/chair_points.sh:35-88 that
    PointAggregator.__init__/viewmlp read; agg_axis_weight=None avoids the hard-coded
    device="cuda" at PA:246 (the value 1,1,1 selects the same branch of `linear`, PA:422)."""
    return argparse.Namespace(
        act_type="LeakyReLU", point_hyper_dim=256, point_features_dim=32, agg_distance_kernel="linear",
        agg_dist_pers=20, agg_axis_weight=None, num_pos_freqs=10, num_viewdir_freqs=4, view_ori=0,
        dist_xyz_freq=5, agg_feat_xyz_mode="None", agg_alpha_xyz_mode="None", agg_color_xyz_mode="None",
        sh_degree=4, weight_feat_dim=8, weight_xyz_freq=2, num_feat_freqs=3, agg_intrp_order=2,
        shading_feature_mlp_layer0=1, shading_feature_mlp_layer1=2, shading_feature_mlp_layer2=0,
        shading_feature_mlp_layer3=2, shading_feature_num=256, point_color_mode="1", point_dir_mode="1",
        shading_alpha_mlp_layer=1, shading_color_mlp_layer=4, shading_color_channel_num=3,
        apply_pnt_mask=1, dist_xyz_deno=0, agg_weight_norm=1, act_super=1, sparse_loss_weight=0,
        zero_one_loss_items="conf_coefficient", prob=0, which_agg_model="viewmlp",
    )
