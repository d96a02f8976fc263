"""The seeded inputs the golden fixtures were generated from (shared by
tests/golden/make_golden.py and the tests that consume the fixtures)."""
import numpy as np

from gsdf_slam_b200 import scenes

FULL_SCENES = {
    # small enough that the stored reference outputs stay ~1 MB per scene
    "tiny_sh3": lambda: scenes.tiny(seed=101, P=1500, W=152, H=104, degree=3),
    "tiny_sh0_empty": lambda: scenes.frustum_scene(1500, 144, 96, 80.0, 80.0, seed=102, degree=0,
                                                   sh_coeffs=0, zmin=0.4, zmax=5.0, log_scale=0.05,
                                                   log_sigma=0.6, bg=(0.2, 0.1, 0.0), name="tiny_sh0"),
}

DIGEST_SCENES = {
    "cfg1_30k": lambda: scenes.cfg1(P=30_000),
    "room_60k": lambda: scenes.room_scene(60_000, 1200, 680, 600.0, 600.0, seed=2, name="room60k"),
    "outdoor_50k": lambda: scenes.outdoor_scene(50_000, 1241, 376, 718.856, 718.856, seed=4),
}


def knn_points():
    rng = np.random.default_rng(77)
    return (rng.standard_normal((5000, 3)) * np.asarray([4.0, 1.0, 3.0])).astype(np.float32)


def adam_case():
    rng = np.random.default_rng(78)
    p = rng.standard_normal(20_000).astype(np.float32)
    grads = []
    for s in range(4):
        g = (rng.standard_normal(20_000) * 10.0 ** (-s)).astype(np.float32)
        g[::5] = 0.0
        grads.append(g)
    return p, grads, 0.0025
