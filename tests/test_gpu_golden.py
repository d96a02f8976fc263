"""GPU parity against the COMMITTED golden fixtures (tests/golden/, generated from the
reference itself by tests/golden/make_golden.py) — needs no reference library at run time."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from golden_scenes import DIGEST_SCENES, FULL_SCENES, adam_case, knn_points
from util import product_exports, rel_l2, run_product_backward, run_product_forward, to_dev

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sha(t):
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()


@pytest.mark.parametrize("name", sorted(FULL_SCENES))
def test_product_matches_stored_reference_outputs(name):
    scene = FULL_SCENES[name]()
    gold = np.load(os.path.join(GOLD, f"ref_{name}.npz"))
    t = to_dev(scene)
    fwd = run_product_forward(scene, t)
    ex = product_exports(scene, t, fwd)
    dev = fwd[2].device
    G = lambda k: torch.from_numpy(gold[k]).to(dev)  # noqa: E731
    assert fwd[0] == int(gold["R_valid"])
    assert torch.equal(fwd[3], G("radii"))
    assert torch.equal(ex["keys"], G("keys")) and torch.equal(ex["ids"], G("ids"))
    assert torch.equal(ex["ranges"], G("ranges"))
    assert torch.equal(ex["n_contrib"], G("n_contrib"))
    assert (fwd[2] - G("color")).abs().max().item() <= 1e-5
    assert (ex["final_T"] - G("final_T")).abs().max().item() <= 1e-5
    g = run_product_backward(scene, t, fwd)
    for k, v in g.items():
        r = G("g_" + k)
        assert rel_l2(v.reshape(r.shape), r) <= 1e-4, k


@pytest.mark.parametrize("name", sorted(DIGEST_SCENES))
def test_bit_exact_digests_of_larger_scenes(name):
    d = json.load(open(os.path.join(GOLD, "digests.json")))[name]
    scene = DIGEST_SCENES[name]()
    t = to_dev(scene)
    fwd = run_product_forward(scene, t)
    ex = product_exports(scene, t, fwd)
    assert fwd[0] == d["R_valid"]
    assert _sha(ex["keys"]) == d["keys"], "sorted 64-bit keys"
    assert _sha(ex["ids"]) == d["ids"], "sorted Gaussian ids"
    assert _sha(ex["ranges"]) == d["ranges"], "tile ranges"
    assert _sha(ex["n_contrib"]) == d["n_contrib"], "per-pixel contributor counts"
    assert _sha(fwd[3]) == d["radii"], "radii"


def test_knn_and_adam_against_golden():
    import gsdf_slam_b200 as gs
    pts = torch.from_numpy(knn_points()).cuda()
    assert torch.equal(gs.distCUDA2(pts).cpu(), torch.from_numpy(np.load(os.path.join(GOLD, "knn.npz"))["dist"]))
    gold = np.load(os.path.join(GOLD, "adam.npz"))
    p, grads, lr = adam_case()
    pt = torch.from_numpy(p).cuda()
    opt = gs.FusedAdam([{"params": [pt], "lr": lr}], eps=1e-15)
    for g in grads:
        opt.step(grads={pt: torch.from_numpy(g).cuda()})
    st = opt.state[pt]
    assert torch.equal(st["exp_avg"].cpu(), torch.from_numpy(gold["m"]))
    assert torch.equal(st["exp_avg_sq"].cpu(), torch.from_numpy(gold["v"]))
    ulp = (pt.cpu().view(torch.int32).long() - torch.from_numpy(gold["p"]).view(torch.int32).long()).abs().max()
    assert int(ulp) <= 2
