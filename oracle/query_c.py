"""ctypes front-end of oracle/query_oracle.c.  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "libquery_oracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "libquery_oracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.qo_build.restype = ctypes.c_void_p
        L.qo_build.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                               ctypes.c_int, ctypes.c_void_p]
        L.qo_free.argtypes = [ctypes.c_void_p]
        L.qo_num_voxels.argtypes = [ctypes.c_void_p]
        L.qo_select.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.qo_query.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                               ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def woord_query_grid_point_index(raypos, xyz, kernel_size, query_size, SR, K, frame, P, radius, want_stats=False):
    """Same contract as oracle.grid_query.woord_query_grid_point_index, computed by the C restatement.
    Returns uncompacted (R,SR,K) pidx, (R,SR,3) loc, (R,SR) mask, (R,) ray_hit [, stats (R,SR,2)]."""
    L = lib()
    raypos = np.ascontiguousarray(raypos, dtype=np.float32)
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    R, D, _ = raypos.shape
    lo, sv = np.ascontiguousarray(frame.lo, np.float32), np.ascontiguousarray(frame.sv, np.float32)
    dim, qs = np.ascontiguousarray(frame.dim, np.int32), np.ascontiguousarray(query_size, np.int32)
    g = L.qo_build(_p(xyz), len(xyz), _p(lo), _p(sv), _p(dim), int(P), _p(qs))
    try:
        loc = np.zeros((R, SR, 3), np.float32)
        mask = np.zeros((R, SR), np.int32)
        hit = np.zeros(R, np.uint8)
        L.qo_select(g, _p(raypos), R, D, SR, _p(loc), _p(mask), _p(hit))
        pidx = np.full((R, SR, K), -1, np.int32)
        stats = np.zeros((R, SR, 2), np.int32)
        L.qo_query(g, _p(xyz), _p(loc), _p(mask), R * SR, K, int(kernel_size[0]), ctypes.c_float(float(radius)),
                   _p(pidx), _p(stats))
    finally:
        L.qo_free(g)
    if want_stats:
        return pidx, loc, mask.astype(bool), hit.astype(bool), stats
    return pidx, loc, mask.astype(bool), hit.astype(bool)
