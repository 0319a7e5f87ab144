"""pointnerf2studio_b200: Point-NeRF's per-ray hot path (pointnerf-original plugin) on B200 / sm_100a.

Host code is Python/PyTorch (device memory, streams, autograd edges, torch.distributed); all compute
is hand-written CUDA in libpnerf_b200.so behind the C ABI of include/pnerf_b200.h.
"""
from .model import PointNerf, PointNerfConfig, NeuralPoints, PointNeRFEncoding, RayBundle  # noqa: F401
from .query_worldcoords import woord_query_grid_point_index  # noqa: F401

__version__ = "0.1.0"
