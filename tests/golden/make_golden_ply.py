"""Generates tests/golden/gaussians_deg{1,3}.ply with the reference's own PLY writer (its vendored tinyply,
driven with savePly's property sequence by oracle/_ref/ply_ref — see oracle/ref_ply.cpp), from the seeded
model of tests/test_ply.py::model_arrays.  Runs in the build container (CPU only):
    bash oracle/build_ref_ply.sh && python tests/golden/make_golden_ply.py"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def model_arrays(P, degree, seed):
    rng = np.random.default_rng(seed)
    M = (degree + 1) ** 2 - 1
    f = lambda *s: rng.standard_normal(s).astype(np.float32)  # noqa: E731
    return {"xyz": f(P, 3), "f_dc": f(P, 1, 3), "f_rest": f(P, M, 3), "opacity": f(P, 1), "scaling": f(P, 3),
            "rotation": f(P, 4)}


def write_raw(path, a):
    P, M = a["xyz"].shape[0], a["f_rest"].shape[1]
    with open(path, "wb") as fh:
        fh.write(np.array([P, M], np.int32).tobytes())
        for k in ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation"):
            fh.write(np.ascontiguousarray(a[k], np.float32).tobytes())


CASES = {"gaussians_deg1": dict(P=7, degree=1, seed=1), "gaussians_deg3": dict(P=33, degree=3, seed=2)}

if __name__ == "__main__":
    exe = os.path.join(ROOT, "oracle", "_ref", "ply_ref")
    for name, kw in CASES.items():
        raw = f"/tmp/{name}.bin"
        write_raw(raw, model_arrays(**kw))
        out = os.path.join(ROOT, "tests", "golden", name + ".ply")
        subprocess.check_call([exe, "write", raw, out])
        print(out, os.path.getsize(out))
