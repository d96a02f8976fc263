"""Host-side mirror of the reference's ``GaussianModel`` map surgery.

Same method names and argument meaning as /root/reference/src/gaussian_model.cpp:

* ``prunePoints(mask)``                      — :431-490
* ``densificationPostfix(new_xyz, ...)``     — :492-569
* ``densifyAndPrune(max_grad, min_opacity, extent, max_screen_size)`` — :644-665
  (densifyAndClone :614-642 and densifyAndSplit :571-612 are fused into it)
* ``addDensificationStats(viewspace_grad, radii)`` — :667-681 plus the max_radii2D update of
  src/gaussian_mapper.cpp:443-447
* ``resetOpacity()``                         — :399-408
* ``increasePcd(points, colors, iteration)`` — :201-317
* ``createFromPcd(points, colors, spatial_lr_scale)`` — :130-199
* ``savePly(path)`` / ``loadPly(path)``       — :861-953 / :688-859 (ply.py)
* ``step(grads)``                            — optimizer_->step() over the six groups (:319-352)

What differs from the reference is where the rows live.  The reference holds 22 separately
allocated tensors and rebuilds every one of them (index / cat / zeros_like, then emptyCache) for
each of these calls.  Here the map is a ``GaussianStore``: one set of row arrays with spare
capacity plus a twin for out-of-place compaction; the public tensors (``xyz_``, ``features_dc_`` …)
are zero-copy views of the first ``P`` rows.  PyTorch is only the memory container: every operation
is one call into libgsr.so (csrc/densify.cu), there is no eager fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from .rasterizer import _stream, distCUDA2

GROUP_NAMES = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")  # param_groups order


class GaussianStore:
    """Row arrays (six parameter groups, their Adam moments, four statistics) with capacity."""

    def __init__(self, n_rest: int, capacity: int, device, with_moments: bool = True):
        self.n_rest = int(n_rest)
        self.capacity = int(capacity)
        self.rows = 0
        self.device = torch.device(device)
        self.widths = (3, 3, 3 * self.n_rest, 1, 3, 4)
        f32 = dict(dtype=torch.float32, device=self.device)
        alloc = lambda w: torch.empty((self.capacity * w,), **f32)  # noqa: E731
        self.param: List[torch.Tensor] = [alloc(w) for w in self.widths]
        self.exp_avg: List[Optional[torch.Tensor]] = [alloc(w) if with_moments else None for w in self.widths]
        self.exp_avg_sq: List[Optional[torch.Tensor]] = [alloc(w) if with_moments else None for w in self.widths]
        self.exist_since_iter = torch.empty((self.capacity,), dtype=torch.int32, device=self.device)
        self.xyz_gradient_accum = alloc(1)
        self.denom = alloc(1)
        self.max_radii2D = alloc(1)

    def c_struct(self) -> _lib.GaussianStoreC:
        s = _lib.GaussianStoreC()
        s.rows, s.capacity, s.n_rest = self.rows, self.capacity, self.n_rest
        for g in range(6):
            s.param[g] = self.param[g].data_ptr() if self.widths[g] else None
            s.exp_avg[g] = self.exp_avg[g].data_ptr() if (self.exp_avg[g] is not None and self.widths[g]) else None
            s.exp_avg_sq[g] = (self.exp_avg_sq[g].data_ptr()
                               if (self.exp_avg_sq[g] is not None and self.widths[g]) else None)
        s.exist_since_iter = self.exist_since_iter.data_ptr()
        s.xyz_gradient_accum = self.xyz_gradient_accum.data_ptr()
        s.denom = self.denom.data_ptr()
        s.max_radii2D = self.max_radii2D.data_ptr()
        return s

    def rows_view(self, t: torch.Tensor, width: int, shape_tail: Sequence[int]) -> torch.Tensor:
        return t[: self.rows * width].view(self.rows, *shape_tail)


class GaussianModel:
    """The map: parameters, Adam state and densification statistics of all Gaussians."""

    VOXEL_SIZE = 0.05  # increasePcd, src/gaussian_model.cpp:214

    def __init__(self, sh_degree: int = 3, device="cuda", capacity: int = 0, percent_dense: float = 0.01,
                 growth: float = 1.25):
        self.max_sh_degree_ = int(sh_degree)
        self.n_rest = (self.max_sh_degree_ + 1) ** 2 - 1
        self.device = torch.device(device)
        self.percent_dense_ = float(percent_dense)
        self.growth = float(growth)
        self._cur = GaussianStore(self.n_rest, max(int(capacity), 1024), self.device)
        self._alt: Optional[GaussianStore] = None
        self._ws = torch.empty(0, dtype=torch.uint8, device=self.device)
        self.adam_step_ = 0          # torch::optim::AdamParamState::step, shared by the six groups
        self.lr_ = {n: 0.0 for n in GROUP_NAMES}
        self.betas_ = (0.9, 0.999)
        self.eps_ = 1e-15            # src/gaussian_model.cpp:349
        self.split_seed_ = 0

    # ---- zero-copy views with the reference's shapes ----
    @property
    def P(self) -> int:
        return self._cur.rows

    @property
    def xyz_(self):
        return self._cur.rows_view(self._cur.param[0], 3, (3,))

    @property
    def features_dc_(self):
        return self._cur.rows_view(self._cur.param[1], 3, (1, 3))

    @property
    def features_rest_(self):
        return self._cur.rows_view(self._cur.param[2], 3 * self.n_rest, (self.n_rest, 3))

    @property
    def opacity_(self):
        return self._cur.rows_view(self._cur.param[3], 1, (1,))

    @property
    def scaling_(self):
        return self._cur.rows_view(self._cur.param[4], 3, (3,))

    @property
    def rotation_(self):
        return self._cur.rows_view(self._cur.param[5], 4, (4,))

    @property
    def exist_since_iter_(self):
        return self._cur.exist_since_iter[: self.P]

    @property
    def xyz_gradient_accum_(self):
        return self._cur.xyz_gradient_accum[: self.P].view(self.P, 1)

    @property
    def denom_(self):
        return self._cur.denom[: self.P].view(self.P, 1)

    @property
    def max_radii2D_(self):
        return self._cur.max_radii2D[: self.P]

    def params(self) -> List[torch.Tensor]:
        return [self.xyz_, self.features_dc_, self.features_rest_, self.opacity_, self.scaling_, self.rotation_]

    def exp_avg(self, g: int) -> torch.Tensor:
        w = self._cur.widths[g]
        return self._cur.exp_avg[g][: self.P * w]

    def exp_avg_sq(self, g: int) -> torch.Tensor:
        w = self._cur.widths[g]
        return self._cur.exp_avg_sq[g][: self.P * w]

    # ---- plumbing ----
    def _workspace(self, nbytes: int) -> C.c_void_p:
        if self._ws.numel() < nbytes:
            self._ws = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=self.device)
        return C.c_void_p(self._ws.data_ptr())

    def _twin(self, min_capacity: int) -> GaussianStore:
        if self._alt is None or self._alt.capacity < min_capacity:
            self._alt = None  # release before allocating the bigger one
            self._alt = GaussianStore(self.n_rest, int(min_capacity), self.device)
        return self._alt

    def _swap(self, new_rows: int) -> None:
        self._cur, self._alt = self._alt, self._cur
        self._cur.rows = int(new_rows)

    def _reserve(self, rows: int) -> None:
        """Make the current store hold at least `rows` rows (copying once into a larger twin)."""
        if rows <= self._cur.capacity:
            return
        old = self._cur
        big = GaussianStore(self.n_rest, int(rows * self.growth) + 4096, self.device)
        big.rows = old.rows
        n = old.rows
        for g, w in enumerate(old.widths):
            big.param[g][: n * w].copy_(old.param[g][: n * w])
            big.exp_avg[g][: n * w].copy_(old.exp_avg[g][: n * w])
            big.exp_avg_sq[g][: n * w].copy_(old.exp_avg_sq[g][: n * w])
        big.exist_since_iter[:n].copy_(old.exist_since_iter[:n])
        big.xyz_gradient_accum[:n].copy_(old.xyz_gradient_accum[:n])
        big.denom[:n].copy_(old.denom[:n])
        big.max_radii2D[:n].copy_(old.max_radii2D[:n])
        self._cur = big
        self._alt = None

    # ---- construction ----
    @torch.no_grad()
    def load(self, xyz, features_dc, features_rest, opacity, scaling, rotation, exist_since_iter=None,
             exp_avg: Optional[Sequence[torch.Tensor]] = None, exp_avg_sq: Optional[Sequence[torch.Tensor]] = None,
             xyz_gradient_accum=None, denom=None, max_radii2D=None) -> None:
        """Fill the store from explicit tensors (createFromPcd's result, a checkpoint, a test state)."""
        n = int(xyz.shape[0])
        self._reserve(n)
        st = self._cur
        st.rows = n
        src = [xyz, features_dc, features_rest, opacity, scaling, rotation]
        for g, w in enumerate(st.widths):
            if w == 0:
                continue
            st.param[g][: n * w].copy_(src[g].reshape(-1))
            if exp_avg is not None:
                st.exp_avg[g][: n * w].copy_(exp_avg[g].reshape(-1))
            else:
                st.exp_avg[g][: n * w].zero_()
            if exp_avg_sq is not None:
                st.exp_avg_sq[g][: n * w].copy_(exp_avg_sq[g].reshape(-1))
            else:
                st.exp_avg_sq[g][: n * w].zero_()
        if exist_since_iter is not None:
            st.exist_since_iter[:n].copy_(exist_since_iter)
        else:
            st.exist_since_iter[:n].zero_()
        for dst, val in ((st.xyz_gradient_accum, xyz_gradient_accum), (st.denom, denom),
                         (st.max_radii2D, max_radii2D)):
            if val is not None:
                dst[:n].copy_(val.reshape(-1))
            else:
                dst[:n].zero_()

    # ---- reference methods ----
    @torch.no_grad()
    def prunePoints(self, mask: torch.Tensor) -> None:
        """Drop the rows where `mask` is true (src/gaussian_model.cpp:431-490)."""
        L = _lib.lib()
        n = self.P
        if mask.numel() != n:
            raise RuntimeError(f"prunePoints: mask has {mask.numel()} elements, the model {n} rows")
        if n == 0:
            return
        m8 = mask.to(device=self.device).reshape(-1).to(torch.uint8).contiguous()
        dst = self._twin(n)
        dst.rows = 0
        count = (C.c_int64 * 4)()
        with torch.cuda.device(self.device):
            ws = self._workspace(L.gsr_store_workspace_bytes(n))
            s, d = self._cur.c_struct(), dst.c_struct()
            _lib.check(L.gsr_prune_points(C.byref(s), C.byref(d), C.c_void_p(m8.data_ptr()), ws, count,
                                          _stream(self.device)), "gsr_prune_points")
        self._swap(count[0])

    @torch.no_grad()
    def densifyAndPrune(self, max_grad: float, min_opacity: float, extent: float, max_screen_size: int,
                        noise: Optional[torch.Tensor] = None, n_split: int = 2) -> Dict[str, int]:
        """Clone + split + prune in one pass (src/gaussian_model.cpp:644-665).  `noise` ([n_split*S, 3]
        or longer, standard normal) pins the split samples; by default they are drawn on the device."""
        L = _lib.lib()
        n = self.P
        if n == 0:
            return {"rows": 0, "kept": 0, "cloned": 0, "split": 0}
        a = _lib.DensifyArgs(max_grad=float(max_grad), min_opacity=float(min_opacity), extent=float(extent),
                             percent_dense=float(self.percent_dense_), max_screen_size=int(max_screen_size),
                             n_split=int(n_split), noise=None, seed=self.split_seed_)
        self.split_seed_ += 1
        if noise is not None:
            noise = noise.to(device=self.device, dtype=torch.float32).contiguous()
            a.noise = noise.data_ptr()
            a.noise_rows = int(noise.numel() // 3)
        counts = (C.c_int64 * 4)()
        want = int(n * self.growth) + 4096
        with torch.cuda.device(self.device):
            ws = self._workspace(L.gsr_store_workspace_bytes(n))
            for _ in range(2):
                dst = self._twin(want)
                dst.rows = 0
                s, d = self._cur.c_struct(), dst.c_struct()
                rc = L.gsr_densify_and_prune(C.byref(s), C.byref(d), C.byref(a), ws, counts, _stream(self.device))
                if rc != -2:  # GSR_ERR_ALLOC: the twin is too small, counts[0] rows are needed
                    break
                want = int(counts[0] * self.growth) + 4096
            _lib.check(rc, "gsr_densify_and_prune")
        self._swap(counts[0])
        return {"rows": int(counts[0]), "kept": int(counts[1]), "cloned": int(counts[2]), "split": int(counts[3])}

    @torch.no_grad()
    def densificationPostfix(self, new_xyz, new_features_dc, new_features_rest, new_opacities, new_scaling,
                             new_rotation, new_exist_since_iter) -> None:
        """Append rows with zero Adam moments; resets the statistics (src/gaussian_model.cpp:492-569)."""
        L = _lib.lib()
        n_new = int(new_xyz.shape[0])
        self._reserve(self.P + n_new)
        new = [t.to(device=self.device, dtype=torch.float32).contiguous()
               for t in (new_xyz, new_features_dc, new_features_rest, new_opacities, new_scaling, new_rotation)]
        ex = new_exist_since_iter.to(device=self.device, dtype=torch.int32).contiguous()
        ptrs = (C.c_void_p * 6)(*[t.data_ptr() if t.numel() else None for t in new])
        with torch.cuda.device(self.device):
            s = self._cur.c_struct()
            _lib.check(L.gsr_densification_postfix(C.byref(s), ptrs, C.c_void_p(ex.data_ptr()) if n_new else None,
                                                   n_new, _stream(self.device)), "gsr_densification_postfix")
        self._cur.rows += n_new

    @torch.no_grad()
    def addDensificationStats(self, viewspace_point_grad: torch.Tensor, radii: torch.Tensor) -> None:
        """max_radii2D / xyz_gradient_accum / denom update for the rows with radii > 0
        (src/gaussian_mapper.cpp:443-447, src/gaussian_model.cpp:667-681).  `viewspace_point_grad` is
        dL_dmeans2D [P,3] as RasterizeGaussiansBackwardCUDA returns it."""
        L = _lib.lib()
        n = self.P
        if viewspace_point_grad.shape != (n, 3) or radii.numel() != n:
            raise RuntimeError("addDensificationStats: expected dL_dmeans2D [P,3] and radii [P]")
        g = viewspace_point_grad.contiguous()
        r = radii.to(torch.int32).contiguous()
        st = self._cur
        with torch.cuda.device(self.device):
            _lib.check(L.gsr_add_densification_stats(n, C.c_void_p(r.data_ptr()), C.c_void_p(g.data_ptr()),
                                                     C.c_void_p(st.max_radii2D.data_ptr()),
                                                     C.c_void_p(st.xyz_gradient_accum.data_ptr()),
                                                     C.c_void_p(st.denom.data_ptr()), _stream(self.device)),
                       "gsr_add_densification_stats")

    @torch.no_grad()
    def resetOpacity(self, cap: float = 1.0) -> None:
        """src/gaussian_model.cpp:399-408.  The reference's cap tensor is ones_like(opacity * 0.01), i.e.
        1 — its reset only round-trips the values and zeroes the Adam moments; cap=0.01 gives the
        original 3DGS behaviour."""
        L = _lib.lib()
        st = self._cur
        with torch.cuda.device(self.device):
            _lib.check(L.gsr_reset_opacity(self.P, C.c_void_p(st.param[3].data_ptr()),
                                           C.c_void_p(st.exp_avg[3].data_ptr()),
                                           C.c_void_p(st.exp_avg_sq[3].data_ptr()), float(cap),
                                           _stream(self.device)), "gsr_reset_opacity")

    @torch.no_grad()
    def increasePcd(self, new_point_cloud: torch.Tensor, new_colors: torch.Tensor, iteration: int) -> int:
        """Add the points that do not fall into an occupied 5 cm voxel hash (src/gaussian_model.cpp:207-317).
        Returns the number of Gaussians added."""
        L = _lib.lib()
        m = int(new_point_cloud.shape[0])
        if m == 0:
            return 0
        pts = new_point_cloud.to(device=self.device, dtype=torch.float32).contiguous()
        rgb = new_colors.to(device=self.device, dtype=torch.float32).contiguous()
        keep = torch.empty((m,), dtype=torch.uint8, device=self.device)
        idx = torch.empty((m,), dtype=torch.int32, device=self.device)
        count = (C.c_int64 * 4)()
        stream = _stream(self.device)
        with torch.cuda.device(self.device):
            nbytes = max(L.gsr_voxel_dedup_workspace_bytes(self.P), L.gsr_store_workspace_bytes(m))
            ws = self._workspace(nbytes)
            _lib.check(L.gsr_voxel_dedup(C.c_void_p(self._cur.param[0].data_ptr()), self.P,
                                         C.c_void_p(pts.data_ptr()), m, float(self.VOXEL_SIZE),
                                         C.c_void_p(keep.data_ptr()), ws, stream), "gsr_voxel_dedup")
            _lib.check(L.gsr_select_rows(C.c_void_p(keep.data_ptr()), m, 0, C.c_void_p(idx.data_ptr()), None, ws,
                                         count, stream), "gsr_select_rows")
            n = int(count[0])
            self._reserve(self.P + n)
            s = self._cur.c_struct()
            _lib.check(L.gsr_increase_pcd_rows(C.byref(s), C.c_void_p(pts.data_ptr()), C.c_void_p(rgb.data_ptr()),
                                               C.c_void_p(idx.data_ptr()), n, 0, None, int(iteration), stream),
                       "gsr_increase_pcd_rows")
            P = self.P
            new_xyz = self._cur.param[0][3 * P: 3 * (P + n)].view(n, 3)
            dist2 = distCUDA2(new_xyz) if n else torch.empty(0, dtype=torch.float32, device=self.device)
            _lib.check(L.gsr_increase_pcd_rows(C.byref(s), C.c_void_p(pts.data_ptr()), C.c_void_p(rgb.data_ptr()),
                                               C.c_void_p(idx.data_ptr()), n, 1,
                                               C.c_void_p(dist2.data_ptr()) if n else None, int(iteration), stream),
                       "gsr_increase_pcd_rows")
        self._cur.rows += n
        return n

    @torch.no_grad()
    def createFromPcd(self, point_cloud: torch.Tensor, color: torch.Tensor, spatial_lr_scale: float = 1.0) -> None:
        """Initialise the map from a coloured point cloud (src/gaussian_model.cpp:130-199).  The reference
        computes a voxel-unique index set there but discards the result of its index_select calls, so every
        point becomes a Gaussian; features_dc = RGB2SH(colour), scaling from distCUDA2, identity rotation,
        opacity inverse_sigmoid(0.1), exist_since_iter 0."""
        L = _lib.lib()
        self.spatial_lr_scale_ = float(spatial_lr_scale)
        n = int(point_cloud.shape[0])
        pts = point_cloud.to(device=self.device, dtype=torch.float32).contiguous()
        rgb = color.to(device=self.device, dtype=torch.float32).contiguous()
        self._cur.rows = 0
        self._reserve(n)
        idx = torch.arange(n, dtype=torch.int32, device=self.device)
        stream = _stream(self.device)
        with torch.cuda.device(self.device):
            s = self._cur.c_struct()
            args = (C.byref(s), C.c_void_p(pts.data_ptr()), C.c_void_p(rgb.data_ptr()), C.c_void_p(idx.data_ptr()), n)
            _lib.check(L.gsr_increase_pcd_rows(*args, 0, None, 0, stream), "gsr_increase_pcd_rows")
            dist2 = distCUDA2(self._cur.param[0][: 3 * n].view(n, 3)) if n else pts.new_empty(0)
            _lib.check(L.gsr_increase_pcd_rows(*args, 1, C.c_void_p(dist2.data_ptr()) if n else None, 0, stream),
                       "gsr_increase_pcd_rows")
        self._cur.rows = n
        self.adam_step_ = 0

    def savePly(self, result_path) -> None:
        """src/gaussian_model.cpp:861-953 (see ply.py)."""
        from . import ply
        ply.savePly(self, result_path)

    def loadPly(self, ply_path) -> None:
        """src/gaussian_model.cpp:688-859 (see ply.py)."""
        from . import ply
        ply.loadPly(self, ply_path)

    @torch.no_grad()
    def step(self, grads: Sequence[Optional[torch.Tensor]], row_mask: Optional[torch.Tensor] = None) -> None:
        """optimizer_->step() (src/gaussian_mapper.cpp:465): one fused Adam launch over the six groups,
        moments in the store.  With `row_mask` (radii) the gradient is only read for rendered rows."""
        L = _lib.lib()
        self.adam_step_ += 1
        items = []
        for g, name in enumerate(GROUP_NAMES):
            w = self._cur.widths[g]
            if w == 0 or grads[g] is None:
                continue
            gr = grads[g].contiguous()
            items.append(_lib.AdamGroup(param=self._cur.param[g].data_ptr(), grad=gr.data_ptr(),
                                        exp_avg=self._cur.exp_avg[g].data_ptr(),
                                        exp_avg_sq=self._cur.exp_avg_sq[g].data_ptr(), numel=self.P * w,
                                        lr=float(self.lr_[name]), step=self.adam_step_))
        if not items:
            return
        arr = (_lib.AdamGroup * len(items))(*items)
        with torch.cuda.device(self.device):
            if row_mask is None:
                _lib.check(L.gsr_adam_step(arr, len(items), self.betas_[0], self.betas_[1], self.eps_, 1.0,
                                           _stream(self.device)), "gsr_adam_step")
            else:
                rm = row_mask.to(torch.int32).contiguous()
                _lib.check(L.gsr_adam_step_rows(arr, len(items), self.betas_[0], self.betas_[1], self.eps_, 1.0,
                                                self.P, C.c_void_p(rm.data_ptr()), 2, _stream(self.device)),
                           "gsr_adam_step_rows")
