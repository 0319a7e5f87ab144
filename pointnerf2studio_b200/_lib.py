"""ctypes binding of libpnerf_b200.so (the C ABI declared in include/pnerf_b200.h).

There is no CPU fallback and no alternative backend: if the library is missing or a call fails the
error is raised.  `load()` only dlopen()s (safe on a box without a GPU); compute entry points need
a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpnerf_b200.so")

_lib = None

c_float_p = C.c_void_p
c_void_p = C.c_void_p


class GridView(C.Structure):
    _fields_ = [("lo", C.c_float * 3), ("sv", C.c_float * 3), ("dim", C.c_int * 3),
                ("cell_start", C.c_void_p), ("recs", C.c_void_p), ("occ_bits", C.c_void_p)]


class Points(C.Structure):
    _fields_ = [("xyz", C.c_void_p), ("embed", C.c_void_p), ("color", C.c_void_p), ("dir", C.c_void_p),
                ("conf", C.c_void_p), ("Rw2c", C.c_float * 9), ("n", C.c_int64)]


class Camera(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("R_c2w", C.c_float * 9), ("dev", C.c_void_p)]


MLP_FIELDS = ["w1", "b1", "w2", "b2", "w3", "b3", "w4", "b4", "wa", "ba",
              "wc1", "bc1", "wc2", "bc2", "wc3", "bc3", "wc4", "bc4"]


class Mlp(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in MLP_FIELDS]


class MlpGrad(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in MLP_FIELDS]


class Mode(C.Structure):
    _fields_ = [("lrelu_slope", C.c_float), ("density_softplus", C.c_int), ("weight_conf", C.c_int),
                ("bg_mode", C.c_int), ("eval_clamp", C.c_int), ("bg", C.c_float * 3), ("vsize_z", C.c_float)]


class AdamSeg(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_int64), ("step", C.c_int64),
                ("lr", C.c_float)]


class DpAdam(C.Structure):
    _fields_ = [("p", C.c_void_p * 8), ("g", C.c_void_p * 8), ("m", C.c_void_p), ("v", C.c_void_p), ("lo", C.c_int64), ("hi", C.c_int64),
                ("boundary", C.c_int64), ("step", C.c_int64), ("lr", C.c_float * 2), ("world", C.c_int), ("rank", C.c_int), ("hyper_dev", C.c_void_p), ("mc_g", C.c_void_p), ("mc_p", C.c_void_p)]


class RenderBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("sample_loc", "sample_cnt", "sample_pidx", "sample_valid", "sample_ids", "n_samples", "sigma",
                                          "rgb", "out_rgb", "ray_mask", "ray_index", "n_rays")] + \
               [("workspace", C.c_void_p), ("workspace_bytes", C.c_int64), ("scratch", C.c_void_p), ("scratch_bytes", C.c_int64)]


class PnerfError(RuntimeError):
    pass


# name -> (restype, argtypes); every symbol include/pnerf_b200.h declares
SIGNATURES = {
    "pnerf_version": (C.c_int, []),
    "pnerf_last_cuda_error": (C.c_char_p, []),
    "pnerf_device_check": (C.c_int, []),
    "pnerf_host_register": (C.c_int, [C.c_void_p, C.c_int64]),
    "pnerf_host_unregister": (C.c_int, [C.c_void_p]),
    "pnerf_copy_rows_to_host": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]),
    "pnerf_bbox": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pnerf_grid_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int64]),
    "pnerf_grid_build": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "pnerf_sample_select": (C.c_int, [C.POINTER(GridView), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pnerf_sample_select_jitter": (C.c_int, [C.POINTER(GridView), C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float,
                                             C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pnerf_sample_select_jitter_dev": (C.c_int, [C.POINTER(GridView), C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                                 C.c_void_p, C.c_void_p, C.c_void_p]),
    "pnerf_hit_rays": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "pnerf_gather_hit_rays": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
    "pnerf_coarse_t": (C.c_int, [C.c_float, C.c_float, C.c_float, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "pnerf_query": (C.c_int, [C.POINTER(GridView), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                              C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "pnerf_ray_compact": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int64, C.c_void_p]),
    "pnerf_gather_rays": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "pnerf_sample_compact": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                       C.c_void_p]),
    "pnerf_sample_compact_classes": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_int64, C.c_void_p]),
    "pnerf_field_forward_tc_part": (C.c_int, [C.POINTER(Points), C.POINTER(Camera), C.POINTER(Mlp), C.c_void_p, C.POINTER(Mode),
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "pnerf_color_forward_tc": (C.c_int, [C.POINTER(Points), C.POINTER(Camera), C.POINTER(Mlp), C.c_void_p, C.POINTER(Mode), C.c_void_p,
                                         C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "pnerf_scan_workspace_bytes": (C.c_int64, [C.c_int64]),
    "pnerf_field_f32_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int]),
    "pnerf_field_forward_f32": (C.c_int, [C.POINTER(Points), C.POINTER(Camera), C.POINTER(Mlp), C.POINTER(Mode),
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "pnerf_field_backward_f32": (C.c_int, [C.POINTER(Points), C.POINTER(Camera), C.POINTER(Mlp), C.POINTER(Mode),
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.POINTER(MlpGrad), C.c_void_p, C.c_int64, C.c_void_p]),
    "pnerf_field_tc_train_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int]),
    "pnerf_field_forward_tc_train": (C.c_int, [C.POINTER(Points), C.POINTER(Camera), C.POINTER(Mlp), C.c_void_p, C.POINTER(Mode),
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "pnerf_field_backward_tc": (C.c_int, [C.POINTER(Points), C.POINTER(Camera), C.POINTER(Mlp), C.POINTER(Mode),
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.POINTER(MlpGrad), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pnerf_render_train_scratch_bytes": (C.c_int64, [C.c_int, C.c_int]),
    "pnerf_render_train_forward": (C.c_int, [C.POINTER(GridView), C.POINTER(Points), C.POINTER(Camera), C.POINTER(Mlp), C.c_void_p,
                                             C.POINTER(Mode), C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_uint64,
                                             C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                             C.POINTER(RenderBuffers), C.c_void_p]),
    "pnerf_render_train_backward": (C.c_int, [C.POINTER(Points), C.POINTER(Camera), C.POINTER(Mlp), C.POINTER(Mode), C.c_void_p, C.c_void_p,
                                              C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(RenderBuffers), C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.POINTER(MlpGrad), C.c_void_p, C.c_void_p]),
    "pnerf_tc_set_trace": (C.c_int, [C.c_void_p]),
    "pnerf_tc_trace_bytes": (C.c_int64, []),
    "pnerf_tc_wpack_bytes": (C.c_int64, []),
    "pnerf_tc_pack_weights": (C.c_int, [C.POINTER(Mlp), C.c_void_p, C.c_void_p]),
    "pnerf_field_tc_workspace_bytes": (C.c_int64, [C.c_int64]),
    "pnerf_field_forward_tc": (C.c_int, [C.POINTER(Points), C.POINTER(Camera), C.POINTER(Mlp), C.c_void_p, C.POINTER(Mode),
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "pnerf_composite_forward": (C.c_int, [C.POINTER(Camera), C.POINTER(Mode), C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pnerf_composite_backward": (C.c_int, [C.POINTER(Camera), C.POINTER(Mode), C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pnerf_umma_selftest": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "pnerf_adam_step": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "pnerf_dp_adam_step": (C.c_int, [C.POINTER(DpAdam), C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "pnerf_tc_microbench": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pnerf_probe": (C.c_int, [C.POINTER(Points), C.POINTER(Camera), C.POINTER(Mode), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p]),
    "pnerf_probe_filter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                     C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "pnerf_vox_closest_workspace_bytes": (C.c_int64, [C.c_void_p]),
    "pnerf_vox_closest": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "pnerf_masked_mse_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pnerf_masked_mse_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p]),
    "pnerf_conf_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]),
}


def load():
    """dlopen the library and bind every declared symbol; raises if the library or a symbol is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PnerfError(f"{LIB_PATH} not found: build it with `python -m pointnerf2studio_b200.build` "
                         "(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().pnerf_last_cuda_error().decode() if rc == -2 else ""
        names = {-1: "PNERF_ERR_ARG", -2: "PNERF_ERR_CUDA", -3: "PNERF_ERR_WORKSPACE", -4: "PNERF_ERR_ARCH"}
        raise PnerfError(f"{what} failed: {names.get(rc, rc)} {msg}")
