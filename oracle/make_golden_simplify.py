"""Golden outputs of the reference's `Flow.simplify()` (flows.py:600-606), generated from the REAL reference (build
container only; test infrastructure, never imported by the product path).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_simplify.py

For a subset of the cases of `make_golden.py` (same specs, seeds, parameters and inputs, read back from the committed
fixtures) the reference flow is simplified -- LU / Householder / sequential affine layers become
`PlaneBijectiveLinearTransform` (transforms.py:618-695), 1x1-convolution blocks become `Bijective1x1Conv2d`
(transforms.py:1031-1176) -- and the simplified flow's `log_prob`, `backward`, `_forward`, layer class names and
state-dict keys are written to `tests/golden/simplify.npz`.
"""
import importlib.util
import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "ref_shim"), "/root/reference", os.path.dirname(HERE)]
warnings.filterwarnings("ignore")

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
MG = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(MG)

CASES = ["c1_d2_laplace", "d6_hh_normal", "d5_noconj", "d32_h64", "d100_h50_hh", "d64_convnet", "d32_radial_inf",
         "img_c4_4x4", "img_mnist_16x7x7", "img_c6_5x3_plain_channel",
         # a soft-training flow over ConditionalDenseNN: the simplified flow is a plain `Flow`, which substitutes no zero context
         "soft_d40_conddense"]


def main():
    torch.set_num_threads(1)
    out = {}
    for name in CASES:
        z = np.load(os.path.join(MG.OUT, name + ".npz"))
        spec = json.loads(bytes(z["spec"]).decode())
        params = {k[len("param:"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param:")}
        torch.manual_seed(int(z["seed"]))
        ref = MG.build_reference(spec, torch.float32)
        res = ref.load_state_dict(params, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        x, z0 = torch.from_numpy(z["x"]), torch.from_numpy(z["z0"])
        with torch.no_grad():
            assert float((ref.log_prob(x) - torch.from_numpy(z["lp32"])).abs().max()) == 0.0     # the fixture's own flow
            simple = ref.simplify()
            lp, lat, y = simple.log_prob(x), simple.backward(x), simple._forward(z0)
        kinds = [type(l).__name__ + ("/" + type(l.transform).__name__ if hasattr(l, "transform") else "")
                 + ("/" + type(l.block_transform).__name__ if hasattr(l, "block_transform") else "")
                 for l in simple.layers]
        meta = dict(layers=kinds, state_keys=list(simple.state_dict().keys()))
        out[name + ":lp"], out[name + ":z"], out[name + ":y"] = lp.numpy(), lat.numpy(), y.numpy()
        out[name + ":meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        print(name, "simplified vs original log_prob: max |delta| =", float((lp - torch.from_numpy(z["lp32"])).abs().max()),
              kinds[:3])
    np.savez_compressed(os.path.join(MG.OUT, "simplify.npz"), **out)
    print("wrote", os.path.join(MG.OUT, "simplify.npz"))


if __name__ == "__main__":
    main()
