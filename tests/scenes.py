"""Shared seeded scenes for the tests (must stay in sync with tests/golden/make_golden.py)."""
import numpy as np

from pointnerf2studio_b200.synth import make_camera, make_cloud

GOLDEN_CFG = dict(vsize=[0.004] * 3, vscale=[2, 2, 2], kernel_size=[3, 3, 3], query_size=[3, 3, 3],
                  ranges=[-1.2, -1.2, -1.2, 1.2, 1.2, 1.2], D=400, SR=12, K=8, P=12, near=2.0, far=6.0)


def golden_scene():
    cloud = make_cloud(1800, seed=4242, radii=(0.03, 0.042), P=GOLDEN_CFG["P"])
    cam = make_camera()
    c = cam.H // 2
    pix = np.array([(c - 15 + i) * cam.W + (c - 6 + j) for i in range(24) for j in range(24)])
    return cloud, cam, pix
