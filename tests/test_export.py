"""ONNX-exportable reference semantics (usflows_b200/export.py): a pure-PyTorch, traceable reading of a flow that must
agree with the outputs of the REAL reference (golden fixtures) -- and must not leak into the product path."""
import pytest
import torch

from helpers import EXT_CASES, IMG_CASES, SMALL_CASES, SOFT_CASES, build_flow, load_case, rel_err


def _flow(name):
    spec, params, arr = load_case(name)
    return build_flow(spec, params, device="cpu"), arr


@pytest.mark.parametrize("name", SMALL_CASES)
def test_reference_module_reproduces_the_reference(name):
    flow, arr = _flow(name)
    lp = flow.reference_module("log_prob")(arr["x"])
    z = flow.reference_module("backward")(arr["x"])
    y = flow.reference_module("forward")(arr["z0"])
    bound = lambda a32, a64, floor: 3.0 * rel_err(a32, a64) + floor      # noqa: E731  (as tests/test_oracle.py)
    assert rel_err(lp, arr["lp32"]) <= bound(arr["lp32"], arr["lp64"], 5e-6)
    assert rel_err(z, arr["z32"]) <= bound(arr["z32"], arr["z64"], 5e-6)
    assert rel_err(y, arr["y32"]) <= bound(arr["y32"], arr["y64"], 5e-6)
    s = flow.reference_module("sample")(torch.zeros(7, arr["x"].shape[1]))
    assert s.shape == (7, arr["x"].shape[1]) and torch.isfinite(s).all()


@pytest.mark.parametrize("name", EXT_CASES[:5] + EXT_CASES[6:] + IMG_CASES + SOFT_CASES)
def test_reference_module_covers_convnet_radial_and_image_flows(name):
    """ConvNet (vector) / ConvNet2D conditioners, Lp-radial bases, image-shaped events."""
    flow, arr = _flow(name)
    lp = flow.reference_module("log_prob")(arr["x"])
    z = flow.reference_module("backward")(arr["x"])
    y = flow.reference_module("forward")(arr["z0"])
    bound = lambda a32, a64, floor: 3.0 * rel_err(a32, a64) + floor      # noqa: E731
    assert lp.shape == arr["lp32"].shape and z.shape == arr["z32"].shape
    assert rel_err(lp, arr["lp32"]) <= bound(arr["lp32"], arr["lp64"], 5e-6)
    assert rel_err(z, arr["z32"]) <= bound(arr["z32"], arr["z64"], 5e-6)
    assert rel_err(y, arr["y32"]) <= bound(arr["y32"], arr["y64"], 5e-6)
    traced = torch.jit.trace(flow.reference_module("log_prob").eval(), (arr["x"][:4],))
    assert torch.equal(traced(arr["x"][4:]), flow.reference_module("log_prob")(arr["x"][4:]))


@pytest.mark.parametrize("mode", ["log_prob", "backward", "forward"])
def test_reference_module_traces_like_an_exporter_would(mode):
    flow, arr = _flow("d6_hh_normal")
    module = flow.reference_module(mode).eval()
    x = arr["x"][:5] if mode != "forward" else arr["z0"][:5]
    traced = torch.jit.trace(module, (x,))                  # the first stage of the TorchScript ONNX exporter
    other = arr["x"][5:40] if mode != "forward" else arr["z0"][5:40]
    assert torch.equal(traced(other), module(other))        # no shape or value was baked in
    assert not any(p.requires_grad for p in module.parameters())


def test_to_onnx_writes_a_file_or_reports_the_missing_package(tmp_path):
    flow, _ = _flow("c1_d2_laplace")
    path = tmp_path / "flow.onnx"
    try:
        import onnx  # noqa: F401
    except ImportError:
        with pytest.raises(Exception, match="onnx"):
            flow.to_onnx(str(path))
        return
    flow.to_onnx(str(path))
    assert path.stat().st_size > 0


def test_the_product_path_still_has_no_cpu_route():
    flow, arr = _flow("c1_d2_laplace")
    with pytest.raises(RuntimeError, match="CUDA"):
        flow.log_prob(arr["x"])
    with pytest.raises(RuntimeError, match="CUDA"):
        flow.backward(arr["x"])
