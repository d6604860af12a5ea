"""Golden vectors of the eval-time paste, produced by the REFERENCE's own ``RandomPatchTransform.simulation_random_patch``
(VLAAttacker/white_patch/appply_random_transform.py:43-78, exec'd from /root/reference with its one-character indent fix,
exactly as make_golden.py does).  Run in the authoring container:  python tests/golden/make_golden_sim.py
Writes tests/golden/reference_golden_sim.npz (inputs + outputs; nothing from /root/reference is copied)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference_frontend  # noqa: E402


def main():
    fe = load_reference_frontend()
    t = fe.RandomPatchTransform("cpu", False)
    out = {}
    rng = np.random.default_rng(2024)
    cases = [("s64_plain", 64, 16, False, 0, 0.0, 0.0, (5, 9)), ("s64_geo", 64, 16, True, 12, 0.1, -0.05, (20, 30)),
             ("s224_plain", 224, 50, False, 0, 0.0, 0.0, (160, 80)), ("s224_geo", 224, 50, True, -17, 0.15, 0.1, (60, 100)),
             ("s224_geo_default", 224, 70, True, 1, 0.1, 0.1, (0, 0))]
    for tag, S, p, geo, angle, shx, shy, pos in cases:
        img = rng.integers(0, 256, size=(S, S, 3), dtype=np.uint8)
        torch.manual_seed(7)
        patch = torch.rand(3, p, p)
        patch[0, 0, 0] = 1.0   # ToPILImage edge: 1.0 -> 255
        patch[1, 0, 0] = 0.0
        y = t.simulation_random_patch(img, patch, geometry=geo, colorjitter=False, angle=angle, shx=shx, shy=shy, position=pos)
        out[f"{tag}_img"] = img
        out[f"{tag}_patch"] = patch.numpy()
        out[f"{tag}_args"] = np.array([int(geo), angle, shx, shy, pos[0], pos[1]], dtype=np.float64)
        out[f"{tag}_out"] = y
    np.savez_compressed(os.path.join(HERE, "reference_golden_sim.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
