"""Generates tests/golden/colorize_*.npz by running the UNMODIFIED reference colorize_launcher
(oracle/_ref/gsdf_ref_colorize.so, built by oracle/build_ref_colorize.py) on a B200:
    gpurun -- 'python tests/golden/make_golden_colorize.py gpurun_out/'
and copying the files into tests/golden/.  Inputs: tests/ingest_scenes.py."""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ingest_scenes  # noqa: E402


def load_ref():
    path = os.path.join(ROOT, "oracle", "_ref", "gsdf_ref_colorize.so")
    spec = importlib.util.spec_from_file_location("gsdf_ref_colorize", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main(out_dir):
    ref = load_ref()
    for name, kw in ingest_scenes.CASES.items():
        s = ingest_scenes.kitti_like(**kw)
        t = {k: torch.from_numpy(v).cuda() for k, v in s.items()}
        pts, cols = ref.colorize(t["points"].clone(), torch.zeros_like(t["points"]), t["proj"], t["Tr"], t["pose"],
                                 t["image"])
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), points=pts.cpu().numpy(), colors=cols.cpu().numpy())
        print(name, tuple(pts.shape))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden"))
