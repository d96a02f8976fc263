"""Seeded synthetic scenes and cameras for tests, smoke and bench (numpy only).

Follows SURVEY.md §8(d): identical bytes are fed to the reference build, the CPU
oracle and the product.  Camera conventions are the reference's
(/root/reference/src/gaussian_keyframe.cpp:138-232): ``viewmatrix`` is the
TRANSPOSED world->camera matrix, ``projmatrix`` = (P @ W2C)^T with a symmetric
frustum built from the FoV only, ``campos`` the camera centre.
All Gaussian inputs are already ACTIVATED (opacity in (0,1), scale > 0, unit
quaternion), as the rasterizer API expects (src/gaussian_renderer.cpp:74-85).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np


@dataclass
class Camera:
    W: int
    H: int
    tanfovx: float
    tanfovy: float
    viewmatrix: np.ndarray  # [4,4] f32, transposed W2C (element [c][r])
    projmatrix: np.ndarray  # [4,4] f32, transposed full projection
    campos: np.ndarray      # [3] f32


def projection_matrix(znear: float, zfar: float, tanx: float, tany: float) -> np.ndarray:
    """gaussian_keyframe.cpp:204-232 (getProjectionMatrix), fp32 like the reference."""
    top = np.float32(tany) * np.float32(znear)
    bottom = -top
    right = np.float32(tanx) * np.float32(znear)
    left = -right
    P = np.zeros((4, 4), dtype=np.float32)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def make_camera(W: int, H: int, fx: float, fy: float, znear: float = 0.01, zfar: float = 100.0,
                R: Optional[np.ndarray] = None, t: Optional[np.ndarray] = None) -> Camera:
    """R, t: world->camera rotation / translation (x_cam = R x_world + t)."""
    tanx = W / (2.0 * fx)
    tany = H / (2.0 * fy)
    w2c = np.eye(4, dtype=np.float32)
    if R is not None:
        w2c[:3, :3] = R.astype(np.float32)
    if t is not None:
        w2c[:3, 3] = t.astype(np.float32)
    P = projection_matrix(znear, zfar, tanx, tany)
    view_t = np.ascontiguousarray(w2c.T)
    full_t = np.ascontiguousarray((view_t @ P.T).astype(np.float32))
    c2w = np.linalg.inv(w2c.astype(np.float64))
    campos = c2w[:3, 3].astype(np.float32)
    return Camera(W, H, float(tanx), float(tany), view_t, full_t, campos)


def _unit_quats(rng, n):
    q = rng.standard_normal((n, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True).astype(np.float32)
    return q.astype(np.float32)


def _sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x))).astype(np.float32)


@dataclass
class Scene:
    name: str
    cam: Camera
    degree: int
    means3D: np.ndarray
    scales: np.ndarray
    rotations: np.ndarray
    opacity: np.ndarray   # [P,1]
    dc: np.ndarray        # [P,1,3]
    sh: np.ndarray        # [P,M,3]  (M may be 0)
    bg: np.ndarray        # [3]
    dL_dout: np.ndarray   # [3,H,W]
    extra_cams: list = field(default_factory=list)

    @property
    def P(self) -> int:
        return int(self.means3D.shape[0])

    def arrays(self) -> Dict[str, np.ndarray]:
        return dict(means3D=self.means3D, scales=self.scales, rotations=self.rotations,
                    opacity=self.opacity, dc=self.dc, sh=self.sh, bg=self.bg, dL_dout=self.dL_dout,
                    viewmatrix=self.cam.viewmatrix, projmatrix=self.cam.projmatrix,
                    campos=self.cam.campos)


def _grad_image(rng, H, W, sky_rows: int = 0):
    g = (rng.standard_normal((3, H, W)) / np.sqrt(float(H * W))).astype(np.float32)
    if sky_rows > 0:
        g[:, :sky_rows, :] = 0.0
    return g


def frustum_scene(P: int, W: int, H: int, fx: float, fy: float, seed: int, degree: int = 0,
                  sh_coeffs: int = 0, zmin: float = 0.5, zmax: float = 6.0, log_scale: float = 0.03,
                  log_sigma: float = 0.4, op_sigma: float = 1.5, bg=(0.0, 0.0, 0.0),
                  name: str = "frustum", behind_frac: float = 0.05) -> Scene:
    """cfg1-style: points uniform in the view-frustum slab (plus a few behind the camera)."""
    rng = np.random.default_rng(seed)
    cam = make_camera(W, H, fx, fy)
    z = rng.uniform(zmin, zmax, P).astype(np.float32)
    x = (rng.uniform(-1.15, 1.15, P) * cam.tanfovx).astype(np.float32) * z
    y = (rng.uniform(-1.15, 1.15, P) * cam.tanfovy).astype(np.float32) * z
    nb = int(P * behind_frac)
    if nb > 0:
        z[:nb] = rng.uniform(-2.0, 0.25, nb).astype(np.float32)  # culled by z <= 0.2 (mostly)
    means = np.stack([x, y, z], 1).astype(np.float32)
    scales = np.exp(rng.normal(np.log(log_scale), log_sigma, (P, 3))).astype(np.float32)
    rots = _unit_quats(rng, P)
    opacity = _sigmoid(rng.normal(0.0, op_sigma, (P, 1)))
    dc = rng.uniform(-1.5, 1.5, (P, 1, 3)).astype(np.float32)
    sh = (rng.normal(0.0, 0.05, (P, sh_coeffs, 3))).astype(np.float32)
    perm = rng.permutation(P)  # culled points interleaved with visible ones
    return Scene(name, cam, degree, means[perm], scales[perm], rots[perm], opacity[perm], dc[perm],
                 sh[perm], np.asarray(bg, np.float32), _grad_image(rng, H, W))


def room_scene(P: int, W: int, H: int, fx: float, fy: float, seed: int, box=(8.0, 3.0, 6.0),
               log_scale: float = 0.02, log_sigma: float = 0.5, degree: int = 3,
               bg=(0.0, 0.0, 0.0), name: str = "room") -> Scene:
    """cfg2-style: 70 % of the points on the faces of a box around the camera, 30 % inside."""
    rng = np.random.default_rng(seed)
    cam = make_camera(W, H, fx, fy)
    half = np.asarray(box, np.float32) / 2
    pts = rng.uniform(-1.0, 1.0, (P, 3)).astype(np.float32) * half
    n_face = int(0.7 * P)
    face = rng.integers(0, 6, n_face)
    axis, sign = face // 2, (face % 2) * 2 - 1
    pts[np.arange(n_face), axis] = sign.astype(np.float32) * half[axis]
    perm = rng.permutation(P)
    pts = pts[perm]
    scales = np.exp(rng.normal(np.log(log_scale), log_sigma, (P, 3))).astype(np.float32)
    rots = _unit_quats(rng, P)
    opacity = _sigmoid(rng.normal(0.0, 2.0, (P, 1)))
    dc = rng.uniform(-1.5, 1.5, (P, 1, 3)).astype(np.float32)
    sh = rng.normal(0.0, 0.05, (P, 15, 3)).astype(np.float32)
    return Scene(name, cam, degree, pts, scales, rots, opacity, dc, sh, np.asarray(bg, np.float32),
                 _grad_image(rng, H, W))


def outdoor_scene(P: int, W: int, H: int, fx: float, fy: float, seed: int,
                  name: str = "outdoor") -> Scene:
    """cfg4-style (KITTI-like): log-uniform depths in [0.3, 400] m, a ground band of large flat
    splats near the camera (many-tile Gaussians), top 35 % of the rows masked in the gradient."""
    rng = np.random.default_rng(seed)
    cam = make_camera(W, H, fx, fy, znear=0.01, zfar=40.0)
    z = np.exp(rng.uniform(np.log(0.3), np.log(400.0), P)).astype(np.float32)
    x = (rng.uniform(-1.1, 1.1, P) * cam.tanfovx).astype(np.float32) * z
    y = (rng.uniform(-1.1, 1.1, P) * cam.tanfovy).astype(np.float32) * z
    scales = (np.exp(rng.normal(np.log(0.01), 0.6, (P, 3))) * np.sqrt(z)[:, None]).astype(np.float32)
    n_ground = P // 20
    gz = np.exp(rng.uniform(np.log(1.0), np.log(60.0), n_ground)).astype(np.float32)
    z[:n_ground] = gz
    y[:n_ground] = 1.6 + rng.normal(0, 0.02, n_ground).astype(np.float32)
    x[:n_ground] = (rng.uniform(-1.0, 1.0, n_ground) * cam.tanfovx).astype(np.float32) * gz
    scales[:n_ground, 0] = rng.uniform(0.1, 1.0, n_ground)
    scales[:n_ground, 1] = rng.uniform(0.002, 0.01, n_ground)
    scales[:n_ground, 2] = rng.uniform(0.1, 1.0, n_ground)
    rots = _unit_quats(rng, P)
    rots[:n_ground] = np.asarray([1, 0, 0, 0], np.float32) + \
        rng.normal(0, 0.02, (n_ground, 4)).astype(np.float32)
    rots /= np.linalg.norm(rots, axis=1, keepdims=True)
    perm = rng.permutation(P)
    means = np.stack([x, y, z], 1).astype(np.float32)[perm]
    opacity = _sigmoid(rng.normal(0.0, 2.0, (P, 1)))
    dc = rng.uniform(-1.5, 1.5, (P, 1, 3)).astype(np.float32)
    sh = rng.normal(0.0, 0.05, (P, 15, 3)).astype(np.float32)
    return Scene(name, cam, 3, means, scales[perm].astype(np.float32), rots[perm].astype(np.float32),
                 opacity, dc, sh, np.zeros(3, np.float32), _grad_image(rng, H, W, int(0.35 * H)))


def ring_cameras(n: int, W: int, H: int, fx: float, fy: float, radius: float = 1.0):
    """cfg5: n cameras on a circle in the x-z plane, looking outward."""
    cams = []
    for k in range(n):
        th = 2.0 * np.pi * k / n
        fwd = np.asarray([np.sin(th), 0.0, np.cos(th)])
        up = np.asarray([0.0, -1.0, 0.0])  # camera y points down
        right = np.cross(up, fwd)
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        R = np.stack([right, down, fwd], 0)  # rows: camera axes in world coords
        c = radius * fwd
        t = -R @ c
        cams.append(make_camera(W, H, fx, fy, R=R, t=t))
    return cams


# ---- the BASELINE.json configurations ---------------------------------------
def cfg1(P: int = 100_000, sh_coeffs: int = 0) -> Scene:
    return frustum_scene(P, 640, 480, 517.306, 516.469, seed=1, degree=0, sh_coeffs=sh_coeffs,
                         name="cfg1_tum_640x480_sh0")


def cfg2(P: int = 500_000) -> Scene:
    return room_scene(P, 1200, 680, 600.0, 600.0, seed=2, name="cfg2_replica_1200x680_sh3")


def cfg3_sequence(P0: int = 100_000, P1: int = 400_000, iters: int = 200, seed: int = 3):
    """cfg3: TUM-style 640x480 mapping loop whose Gaussian count follows
    P0 -> P1 with clone / split / prune events every 10 iterations (random 5 % split x2, 3 %
    prune), and whose every 4th iteration renders at a pyramid level (80x60, 160x120, 320x240), so
    P, R, T, N all change from call to call.  Yields (iteration, Scene)."""
    rng = np.random.default_rng(seed)
    base = frustum_scene(P1 + P1 // 4, 640, 480, 517.306, 516.469, seed=seed, degree=3, sh_coeffs=15,
                         name="cfg3_pool")
    alive = np.zeros(base.P, dtype=bool)
    alive[rng.permutation(base.P)[:P0]] = True
    grow = (P1 / P0) ** (1.0 / max(1, iters // 10))
    levels = [(80, 60), (160, 120), (320, 240)]
    for it in range(iters):
        if it > 0 and it % 10 == 0:
            idx = np.flatnonzero(alive)
            alive[rng.choice(idx, size=max(1, int(0.03 * idx.size)), replace=False)] = False        # prune
            want = min(int(alive.sum() * grow), P1)
            dead = np.flatnonzero(~alive)
            add = min(dead.size, max(0, want - int(alive.sum())) + int(0.05 * idx.size))            # split / clone
            alive[rng.choice(dead, size=add, replace=False)] = True
        W, H = (640, 480) if it % 4 else levels[(it // 4) % 3]
        cam = make_camera(W, H, 517.306 * W / 640.0, 516.469 * H / 480.0)
        sel = np.flatnonzero(alive)
        yield it, Scene(f"cfg3_it{it}", cam, 3, base.means3D[sel], base.scales[sel], base.rotations[sel],
                        base.opacity[sel], base.dc[sel], base.sh[sel], base.bg, _grad_image(rng, H, W))


def cfg4(P: int = 2_000_000) -> Scene:
    return outdoor_scene(P, 1241, 376, 718.856, 718.856, seed=4, name="cfg4_kitti_1241x376_sh3")


def cfg5(P: int = 3_000_000, n_cams: int = 8) -> Scene:
    s = room_scene(P, 1200, 680, 600.0, 600.0, seed=5, box=(16.0, 3.0, 12.0),
                   name="cfg5_replica_3M_8kf")
    s.extra_cams = ring_cameras(n_cams, 1200, 680, 600.0, 600.0, 1.0)
    return s


def tiny(seed: int = 7, P: int = 3000, W: int = 200, H: int = 136, degree: int = 3,
         bg=(0.1, 0.2, 0.3)) -> Scene:
    """Small scene for unit tests (image size deliberately not a multiple of 16)."""
    fx = W * 0.55
    s = frustum_scene(P, W, H, fx, fx, seed=seed, degree=degree, sh_coeffs=15, zmin=0.4, zmax=5.0,
                      log_scale=0.05, log_sigma=0.6, bg=bg, name=f"tiny_{seed}")
    return s
