"""Golden outputs of the reference's stand-alone layer classes that `USFlow` does not assemble itself: `Rotation`,
`CompositeRotation` (transforms.py:476-616) and `BlockLUTransform` (transforms.py:1488-1622).  Generated from the REAL
reference (build container only; test infrastructure, never imported by the product path).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_layers.py        ->  tests/golden/layers.npz
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "ref_shim"), "/root/reference", os.path.dirname(HERE)]
warnings.filterwarnings("ignore")

from src.usflows import transforms as RT  # noqa: E402  (the real reference)

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "layers.npz")


def main():
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(77)
    out = {}
    # Rotation: forward, matrix (backward of the reference is NOT the inverse, see usflows_b200.transforms.Rotation)
    rot = RT.Rotation(5, (1, 3), 0.7)
    x = torch.randn(6, 5, generator=g)
    out["rot:x"], out["rot:y"], out["rot:matrix"] = x.numpy(), rot.forward(x).numpy(), rot.as_matrix().numpy()
    out["rot:ref_backward_of_y"] = rot.backward(rot.forward(x)).numpy()          # documents the reference's behaviour
    comp = RT.CompositeRotation([RT.Rotation(5, (0, 1), 0.3), RT.Rotation(5, (1, 4), -1.1), RT.Rotation(5, (2, 0), 2.0)])
    out["comp:y"], out["comp:as_matrix"] = comp.forward(x).numpy(), comp.as_matrix().numpy()
    # BlockLUTransform: flat and [C, H, W]
    for tag, in_dims, n in (("blu_flat", [6], 9), ("blu_img", [4, 3, 5], 7)):
        torch.manual_seed(5)
        t = RT.BlockLUTransform(in_dims).to("cpu")      # LUTransform.to sets the `device` attribute log_prior reads
        xs = torch.randn(n, *in_dims, generator=g)
        with torch.no_grad():
            out[f"{tag}:x"] = xs.numpy()
            out[f"{tag}:y"] = t.forward(xs).numpy()
            out[f"{tag}:z"] = t.backward(xs).numpy()
            out[f"{tag}:ladj"] = np.float64(float(t.log_abs_det_jacobian(xs, xs)))
            out[f"{tag}:log_prior"] = np.float64(float(t.log_prior()))
        for k, v in t.state_dict().items():
            out[f"{tag}:param:{k}"] = v.numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, "| reference Rotation round-trip error:",
          float(np.abs(out["rot:ref_backward_of_y"] - out["rot:x"]).max()))


if __name__ == "__main__":
    main()
