"""PLY interchange format (row f-3): the files this repo writes must be byte-identical to what the
reference's savePly writes.  Golden files tests/golden/gaussians_deg*.ply come from the reference's own
vendored tinyply driven with savePly's property sequence (oracle/ref_ply.cpp, tests/golden/make_golden_ply.py).
CPU tests pin the header writer and a numpy restatement of the row layout against them (and, where
oracle/_ref/ply_ref exists, the reference READER against our writer); the GPU tests run the real kernels."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden_ply import CASES, model_arrays, write_raw  # noqa: E402

PLY_REF = os.path.join(ROOT, "oracle", "_ref", "ply_ref")


def _rows_numpy(a):
    """savePly's layout restated with numpy (src/gaussian_model.cpp:864-871)."""
    P = a["xyz"].shape[0]
    f_dc = a["f_dc"].transpose(0, 2, 1).reshape(P, -1)
    f_rest = a["f_rest"].transpose(0, 2, 1).reshape(P, -1)
    return np.concatenate([a["xyz"], np.zeros_like(a["xyz"]), f_dc, f_rest, a["opacity"], a["scaling"], a["rotation"]],
                          axis=1).astype("<f4")


@pytest.mark.parametrize("name", sorted(CASES))
def test_header_and_layout_match_reference_file(name):
    from gsdf_slam_b200 import ply
    a = model_arrays(**CASES[name])
    gold = open(os.path.join(ROOT, "tests", "golden", name + ".ply"), "rb").read()
    mine = ply._header(a["xyz"].shape[0], a["f_rest"].shape[1]) + _rows_numpy(a).tobytes()
    assert mine == gold
    fmt, count, props, off = ply._parse_header(gold)
    assert fmt == "binary_little_endian" and count == a["xyz"].shape[0] and len(props) == 17 + 3 * a["f_rest"].shape[1]
    assert gold[off:] == _rows_numpy(a).tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_save_ply_is_byte_identical_to_the_reference(name, tmp_path):
    import torch
    from gsdf_slam_b200 import ply
    from gsdf_slam_b200.gaussian_model import GaussianModel
    kw = CASES[name]
    a = model_arrays(**kw)
    m = GaussianModel(sh_degree=kw["degree"], device="cuda")
    t = {k: torch.from_numpy(v).cuda() for k, v in a.items()}
    m.load(t["xyz"], t["f_dc"], t["f_rest"], t["opacity"], t["scaling"], t["rotation"])
    out = tmp_path / "out.ply"
    ply.savePly(m, out)
    assert open(out, "rb").read() == open(os.path.join(ROOT, "tests", "golden", name + ".ply"), "rb").read()
    # and back: loadPly restores the model bit for bit
    m2 = GaussianModel(sh_degree=kw["degree"], device="cuda")
    ply.loadPly(m2, os.path.join(ROOT, "tests", "golden", name + ".ply"))
    for got, key in zip(m2.params(), ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")):
        assert torch.equal(got, t[key]), key
    assert float(m2.exp_avg(0).abs().sum()) == 0.0
    if os.path.exists(PLY_REF):
        # the reference's reader (loadPly's property requests) on OUR file
        raw = tmp_path / "back.bin"
        subprocess.check_call([PLY_REF, "read", str(out), str(raw), str(kw["degree"])])
        ref_raw = tmp_path / "ref.bin"
        write_raw(ref_raw, a)
        assert open(raw, "rb").read() == open(ref_raw, "rb").read()


@pytest.mark.gpu
def test_ply_large_round_trip_and_foreign_property_order(tmp_path):
    import torch
    from gsdf_slam_b200 import ply
    from gsdf_slam_b200.gaussian_model import GaussianModel
    a = model_arrays(P=200_003, degree=3, seed=9)
    m = GaussianModel(sh_degree=3, device="cuda")
    t = {k: torch.from_numpy(v).cuda() for k, v in a.items()}
    m.load(t["xyz"], t["f_dc"], t["f_rest"], t["opacity"], t["scaling"], t["rotation"])
    out = tmp_path / "big.ply"
    ply.savePly(m, out)
    assert os.path.getsize(out) == len(ply._header(200_003, 15)) + 200_003 * 62 * 4
    m2 = GaussianModel(sh_degree=3, device="cuda")
    ply.loadPly(m2, out)
    for got, ref in zip(m2.params(), m.params()):
        assert torch.equal(got, ref)
    # a file from another tool: properties shuffled, one extra column, doubles for xyz
    names = ply._property_names(15)
    order = list(reversed(names)) + ["extra"]
    rows = _rows_numpy(a)[:1000]
    dt = np.dtype([(n, "<f8" if n in ("x", "y", "z") else "<f4") for n in order])
    v = np.zeros(1000, dt)
    for j, n in enumerate(names):
        v[n] = rows[:, j]
    hdr = "ply\nformat binary_little_endian 1.0\ncomment shuffled\nelement vertex 1000\n" + "".join(
        f"property {'double' if n in ('x', 'y', 'z') else 'float'} {n}\n" for n in order) + "end_header\n"
    p = tmp_path / "foreign.ply"
    open(p, "wb").write(hdr.encode() + v.tobytes())
    m3 = GaussianModel(sh_degree=3, device="cuda")
    ply.loadPly(m3, p)
    assert m3.P == 1000 and torch.equal(m3.features_rest_, t["f_rest"][:1000]) and torch.equal(m3.xyz_, t["xyz"][:1000])
    with pytest.raises(RuntimeError):
        ply.loadPly(m3, tmp_path / "missing.ply")
