"""The map's on-disk interchange format: binary little-endian PLY with the property
order every 3D-Gaussian-splatting tool reads — x y z nx ny nz f_dc_* f_rest_* opacity scale_* rot_*.

``savePly`` / ``loadPly`` mirror GaussianModel::savePly (/root/reference/src/gaussian_model.cpp:861-953)
and GaussianModel::loadPly (:688-859).  The reference copies seven tensors to the host (two of them
through transposing copies) and has tinyply interleave them value by value; here ``gsr_ply_pack_rows``
builds the interleaved vertex rows on the device and the host writes the header plus one buffer
(``gsr_ply_unpack_rows`` is the inverse).  Files are byte-identical to the reference's
(tests/test_ply.py checks against files written by the reference's own tinyply)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .rasterizer import _stream


def _property_names(n_rest: int):
    names = ["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)]
    names += [f"f_rest_{i}" for i in range(3 * n_rest)]
    names += ["opacity"] + [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)]
    return names


def _header(num_vertices: int, n_rest: int) -> bytes:
    # tinyply write_header (submodules/utils/ply_utils/tinyply.h:664-703): no comments, one element
    lines = ["ply", "format binary_little_endian 1.0", f"element vertex {num_vertices}"]
    lines += [f"property float {n}" for n in _property_names(n_rest)]
    lines.append("end_header")
    return ("\n".join(lines) + "\n").encode("ascii")


@torch.no_grad()
def savePly(model, result_path) -> None:
    """Write `model` (a gaussian_model.GaussianModel) to `result_path`."""
    L = _lib.lib()
    P, M = model.P, model.n_rest
    rows = torch.empty((P, 17 + 3 * M), dtype=torch.float32, device=model.device)
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t.numel() else None  # noqa: E731
    with torch.cuda.device(model.device):
        _lib.check(L.gsr_ply_pack_rows(P, M, ptr(model.xyz_), ptr(model.features_dc_), ptr(model.features_rest_),
                                       ptr(model.opacity_), ptr(model.scaling_), ptr(model.rotation_), ptr(rows),
                                       _stream(model.device)), "gsr_ply_pack_rows")
    host = rows.cpu().numpy()
    try:
        f = open(result_path, "wb")
    except OSError as e:
        raise RuntimeError("failed to open " + str(result_path)) from e   # :874-875
    with f:
        f.write(_header(P, M))
        f.write(host.astype("<f4", copy=False).tobytes())


def _parse_header(buf: bytes):
    end = buf.find(b"end_header\n")
    if not buf.startswith(b"ply") or end < 0:
        raise RuntimeError("not a PLY file")
    lines = buf[:end].decode("ascii", "replace").split("\n")
    fmt, count, props, in_vertex = None, 0, [], False
    for ln in lines:
        tok = ln.split()
        if not tok:
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            in_vertex = tok[1] == "vertex"
            if in_vertex:
                count = int(tok[2])
            elif props:
                break  # only the first (vertex) element is read
        elif tok[0] == "property" and in_vertex:
            if tok[1] == "list":
                raise RuntimeError("list properties in the vertex element are not supported")
            props.append((tok[2], tok[1]))
    return fmt, count, props, end + len(b"end_header\n")


_PLY_TYPES = {"float": "f4", "float32": "f4", "double": "f8", "float64": "f8", "uchar": "u1", "uint8": "u1",
              "char": "i1", "int8": "i1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
              "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4"}


@torch.no_grad()
def loadPly(model, ply_path) -> None:
    """Fill `model` from a PLY file: the properties loadPly requests (x y z, f_dc_0..2, f_rest_0..3M-1 for the
    model's max SH degree, opacity, scale_0..2, rot_0..3) are looked up BY NAME, like tinyply does, so files
    with extra properties or another property order load too.  Adam state and statistics start at zero."""
    if not os.path.exists(ply_path):
        raise RuntimeError("Fail to open ply file at " + str(ply_path))   # :691-692
    buf = open(ply_path, "rb").read()
    fmt, count, props, off = _parse_header(buf)
    if fmt not in ("binary_little_endian", "binary_big_endian"):
        raise RuntimeError("only binary PLY files are supported")
    end = "<" if fmt == "binary_little_endian" else ">"
    dt = np.dtype([(n, end + _PLY_TYPES[t]) for n, t in props])
    verts = np.frombuffer(buf, dtype=dt, count=count, offset=off)
    M = model.n_rest
    want = [n for n in _property_names(M) if n not in ("nx", "ny", "nz")]
    missing = [n for n in want if n not in dt.names]
    if missing:
        raise RuntimeError("PLY file lacks properties: " + ", ".join(missing[:4]))
    rows = np.zeros((count, 17 + 3 * M), dtype=np.float32)
    for j, n in enumerate(_property_names(M)):
        if n in dt.names:
            rows[:, j] = verts[n]
    L = _lib.lib()
    dev = model.device
    drows = torch.from_numpy(rows).to(dev)
    z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)  # noqa: E731
    model.load(z(count, 3), z(count, 1, 3), z(count, M, 3), z(count, 1), z(count, 3), z(count, 4))
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t.numel() else None  # noqa: E731
    with torch.cuda.device(dev):
        _lib.check(L.gsr_ply_unpack_rows(count, M, ptr(drows), ptr(model.xyz_), ptr(model.features_dc_),
                                         ptr(model.features_rest_), ptr(model.opacity_), ptr(model.scaling_),
                                         ptr(model.rotation_), _stream(dev)), "gsr_ply_unpack_rows")
        torch.cuda.current_stream(dev).synchronize()   # drows is released when this returns
