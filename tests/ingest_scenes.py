"""Seeded frame-ingest inputs: a KITTI-like LiDAR scan, camera matrices and image.
Shared by tests/golden/make_golden_colorize.py (which ran the reference on them) and the tests."""
import numpy as np


def kitti_like(n, seed, W=1241, H=376):
    rng = np.random.default_rng(seed)
    # LiDAR frame: x forward, y left, z up; points all around the vehicle, some very close / behind
    r = rng.uniform(0.3, 60.0, n).astype(np.float32)
    az = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    z = rng.uniform(-2.0, 3.0, n).astype(np.float32)
    pts = np.stack([r * np.cos(az), r * np.sin(az), z], 1).astype(np.float32)
    # velodyne -> camera (camera: x right, y down, z forward) with a small offset, row-major 4x4
    Tr = np.array([[0, -1, 0, 0.02], [0, 0, -1, -0.06], [1, 0, 0, -0.27], [0, 0, 0, 1]], np.float32)
    Tr[:3, :3] += rng.normal(0, 0.004, (3, 3)).astype(np.float32)
    fx = 718.856
    P = np.array([[fx, 0, W * 0.49, 44.9], [0, fx, H * 0.47, 0.2], [0, 0, 1, 0.003]], np.float32)
    th = 0.3
    pose = np.array([[np.cos(th), -np.sin(th), 0, 12.5], [np.sin(th), np.cos(th), 0, -3.25], [0, 0, 1, 0.8],
                     [0, 0, 0, 1]], np.float32)
    image = rng.random((H, W, 3), dtype=np.float32)
    return {"points": pts, "Tr": Tr.reshape(-1), "proj": P.reshape(-1), "pose": pose.reshape(-1), "image": image}


CASES = {"colorize_kitti_20k": dict(n=20_000, seed=5), "colorize_small_image": dict(n=5_000, seed=6, W=97, H=61)}
