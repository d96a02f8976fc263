"""Generates the golden fixtures of tests/golden/ from the REFERENCE ITSELF.

Runs on a B200 box (``gpurun -- python tests/golden/make_golden.py``): feeds seeded scenes to
the unmodified reference CUDA rasterizer / simple-knn compiled by oracle/build_ref.sh
(oracle/_ref/libgsdf_ref.so) and stores its outputs.  Output goes to gpurun_out/golden/ (the
only directory gpurun brings back); copy the files into tests/golden/ and commit them.

* ``ref_<scene>.npz``  — full reference outputs of small scenes (forward state, image, all
  gradient tensors); used on CPU to pin the CPU oracle and on GPU as a parity target that does
  not need the reference library at run time.
* ``digests.json``     — SHA-256 of the bit-exact outputs (valid sorted keys / ids, ranges,
  n_contrib, radii) of larger scenes.
* ``knn.npz``, ``adam.npz`` — reference simple-knn output; LibTorch-Adam op sequence output.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from golden_scenes import DIGEST_SCENES, FULL_SCENES, adam_case, knn_points  # noqa: E402
from util import run_reference, to_dev, valid_prefix  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "golden")


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, mk in FULL_SCENES.items():
        scene = mk()
        t = to_dev(scene)
        ref = run_reference(scene, t)
        nv = valid_prefix(ref["keys"])
        arrs = dict(
            R_valid=np.int64(nv), radii=ref["radii"], depths=ref["depths"], means2D=ref["means2D"],
            conic_opacity=ref["conic_opacity"], rgb=ref["rgb"], clamped=ref["clamped"],
            keys=ref["keys"][:nv], ids=ref["ids"][:nv], ranges=ref["ranges"],
            n_contrib=ref["n_contrib"], final_T=ref["final_T"], color=ref["color"])
        for k, v in ref["grads"].items():
            arrs["g_" + k] = v
        np.savez_compressed(os.path.join(OUT, f"ref_{name}.npz"),
                            **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else v)
                               for k, v in arrs.items()})
        print("wrote", name, "R_valid", nv)

    digests = {}
    for name, mk in DIGEST_SCENES.items():
        scene = mk()
        t = to_dev(scene)
        ref = run_reference(scene, t, backward=False)
        nv = valid_prefix(ref["keys"])
        digests[name] = dict(R_valid=nv, keys=sha(ref["keys"][:nv]), ids=sha(ref["ids"][:nv]),
                             ranges=sha(ref["ranges"]), n_contrib=sha(ref["n_contrib"]),
                             radii=sha(ref["radii"]), color=sha(ref["color"]),
                             color_absmax=float(ref["color"].abs().max()))
        print("digest", name, nv)
    with open(os.path.join(OUT, "digests.json"), "w") as f:
        json.dump(digests, f, indent=1, sort_keys=True)

    from oracle import ref_binding as rb
    pts = torch.from_numpy(knn_points()).cuda()
    np.savez_compressed(os.path.join(OUT, "knn.npz"), dist=rb.ref_knn(pts).cpu().numpy())

    p, g_list, lr = adam_case()
    pt = torch.from_numpy(p).cuda()
    m = torch.zeros_like(pt)
    v = torch.zeros_like(pt)
    beta1, beta2, eps = 0.9, 0.999, 1e-15
    for step, g in enumerate(g_list, start=1):  # torch::optim::Adam::step, op by op
        gt = torch.from_numpy(g).cuda()
        bc1 = 1 - beta1 ** step
        bc2 = 1 - beta2 ** step
        m.mul_(beta1).add_(gt, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(gt, gt, value=1 - beta2)
        denom = (v.sqrt() / np.sqrt(bc2)).add_(eps)
        pt.addcdiv_(m, denom, value=-(lr / bc1))
    np.savez_compressed(os.path.join(OUT, "adam.npz"), p=pt.cpu().numpy(), m=m.cpu().numpy(),
                        v=v.cpu().numpy())
    print("done")


if __name__ == "__main__":
    main()
