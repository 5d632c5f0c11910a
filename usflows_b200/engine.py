"""Host-side execution engine: compiles a stack of transform layers into a launch program of fused
kernels and runs it chunk by chunk over the batch rows.

  data -> latent ("backward", `Flow.log_prob`, reference flows.py:225-245) and
  latent -> data ("forward", `Flow.sample` / `_forward`, flows.py:247-265, 45-55).

Planning folds every per-feature vector op into a neighbouring kernel:
  ScaleTransform.backward + the first affine layer's `y - b`  -> the ingest kernel
  `y - b` of later affine layers                              -> `postsub` of the producing contraction
  ScaleTransform.forward                                      -> `colscale` of the last contraction
  bias / ReLU / coupling mask / residual add-sub              -> contraction epilogue
so a USFlow evaluation is one ingest, 5B+1 contractions and (for log_prob) one base-density reduction.

Precision modes (`set_precision`, or `Flow(precision=...)`):
  "fp32"       tcgen05 kind::tf32 with the 3-term split -- fp32-level accuracy (default)
  "fp32_simt"  CUDA-core FFMA contraction (cross-check engine; also used for tiny / unaligned layers)
  "tf32"       tcgen05 kind::tf32 single pass
  "bf16"       tcgen05 kind::f16 with bf16 operands, fp32 accumulate and fp32 residual stream
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import torch

from . import ops
from .ops import Act, ENGINE_SIMT, ENGINE_TC_3XTF32, ENGINE_TC_BF16, ENGINE_TC_TF32, pad4

PRECISIONS = ("fp32", "fp32_simt", "tf32", "bf16")
_default_precision = "fp32"
_default_chunk_rows = 16384
TC_MIN_DIM = 32          # contractions narrower than this run on the SIMT engine


def set_precision(mode: str) -> None:
    global _default_precision
    if mode not in PRECISIONS:
        raise ValueError(f"unknown precision {mode!r}; choose from {PRECISIONS}")
    _default_precision = mode


def get_precision() -> str:
    return _default_precision


def set_chunk_rows(rows: int) -> None:
    global _default_chunk_rows
    _default_chunk_rows = int(rows)


# --------------------------------------------------------------------------------------------------
# operand preparation per precision mode (cached on the owning tensor's identity)
# --------------------------------------------------------------------------------------------------
class _OperandCache:
    """fp32 weight matrix -> the planes a contraction engine reads (tf32 hi/lo split, bf16 copy)."""

    def __init__(self):
        self._c = {}

    def get(self, w: torch.Tensor, mode: str):
        key = (w.data_ptr(), w._version, tuple(w.shape), mode)
        hit = self._c.get(key)
        if hit is not None:
            return hit[1], hit[2]
        rows, cols = w.shape
        ld = pad4(cols)
        if mode == "fp32":
            buf = torch.zeros(2, rows, ld, dtype=torch.float32, device=w.device)
            hi, lo = buf[0, :, :cols], buf[1, :, :cols]
            ops.split_tf32(w, hi, lo)
            res = (hi, lo)
        elif mode == "bf16":
            b = torch.zeros(rows, pad4(cols), dtype=torch.bfloat16, device=w.device)[:, :cols]
            ops.to_bf16(w, b)
            res = (b, None)
        else:  # "tf32" / "fp32_simt": the fp32 matrix itself, re-laid with a 16-byte-multiple pitch if needed
            if w.stride(0) % 4 == 0 and w.data_ptr() % 16 == 0:
                res = (w, None)
            else:
                b = torch.zeros(rows, ld, dtype=torch.float32, device=w.device)[:, :cols]
                b.copy_(w)
                res = (b, None)
        if len(self._c) > 4096:
            self._c.clear()
        self._c[key] = (w, res[0], res[1])   # keep `w` alive so data_ptr stays unique
        return res


_operands = _OperandCache()


def _engine_for(mode: str, N: int, K: int) -> int:
    if mode == "fp32_simt" or min(N, K) < TC_MIN_DIM:
        return ENGINE_SIMT
    return {"fp32": ENGINE_TC_3XTF32, "tf32": ENGINE_TC_TF32, "bf16": ENGINE_TC_BF16}[mode]


# --------------------------------------------------------------------------------------------------
# program steps
# --------------------------------------------------------------------------------------------------
@dataclass
class Step:
    kind: str                                  # "mm" | "leaky" | "permute" | "vec"
    src: str = "x"                             # slot read:  "x" (stream) or "h" (conditioner hidden)
    dst: str = "x"
    w: Optional[torch.Tensor] = None           # [N, K] fp32 prepared weight
    bias: Optional[torch.Tensor] = None
    relu: bool = False
    resid: bool = False                        # out = x + sign * value
    sign: float = 1.0
    colscale: Optional[torch.Tensor] = None
    postsub: Optional[torch.Tensor] = None
    presub: Optional[torch.Tensor] = None      # planning only: subtract from the input first
    vec_div: Optional[torch.Tensor] = None     # "vec" steps: ((x / div) * mul) - sub
    vec_mul: Optional[torch.Tensor] = None
    vec_sub: Optional[torch.Tensor] = None
    slope: float = 1.0
    perm: Optional[torch.Tensor] = None
    needs: set = field(default_factory=set)    # what the consumer of this step's output reads


def _emit_layer(layer, direction: str, steps: List[Step]) -> None:
    from . import transforms as T
    fwd = direction == "forward"
    if isinstance(layer, T.InverseTransform):
        _emit_layer(layer.transform, "backward" if fwd else "forward", steps)
        return
    if isinstance(layer, T.BlockAffineTransform):
        layer = layer.block_transform
    if isinstance(layer, T.AffineTransform):
        p = layer._prepared()
        if fwd:    # x @ W^T + b   (transforms.py:913-934)
            steps.append(Step("mm", w=p["matrix"], bias=p["bias"]))
        else:      # (y - b) @ Winv^T  (transforms.py:936-962)
            steps.append(Step("mm", w=p["inverse_matrix"], presub=p["bias"]))
        return
    if isinstance(layer, T.MaskedCoupling):
        p = layer._prepared()
        ws, bs = p["weights"], p["biases"]
        n = len(ws)
        for j in range(n):
            last = j == n - 1
            steps.append(Step("mm", src="x" if j == 0 else "h", dst="x" if last else "h", w=ws[j], bias=bs[j],
                              relu=not last, resid=last, sign=1.0 if fwd else -1.0))
        return
    if isinstance(layer, T.ScaleTransform):
        s = layer.scale.detach().reshape(-1)
        ops.require_cuda(s, "ScaleTransform.scale")
        steps.append(Step("vec", vec_mul=s) if fwd else Step("vec", vec_div=s))
        return
    if isinstance(layer, T.LeakyReLUTransform):
        steps.append(Step("leaky", slope=layer.alpha if fwd else 1.0 / layer.alpha))
        return
    if isinstance(layer, T.Permute):
        perm = layer.permutation if fwd else layer.inv_permutation
        steps.append(Step("permute", perm=perm.to(torch.int32)))
        return
    raise NotImplementedError(f"usflows_b200: no kernel path for layer type {type(layer).__name__}")


def _fuse(steps: List[Step]) -> List[Step]:
    """Peephole fusion of per-feature vector ops into neighbouring kernels (see module docstring)."""
    out: List[Step] = []
    for st in steps:
        prev = out[-1] if out else None
        if st.kind == "mm" and st.presub is not None and st.src == "x":
            if prev is not None and prev.kind == "mm" and prev.dst == "x" and prev.postsub is None:
                prev.postsub, st.presub = st.presub, None
            elif prev is not None and prev.kind == "vec" and prev.vec_sub is None:
                prev.vec_sub, st.presub = st.presub, None
            else:
                out.append(Step("vec", vec_sub=st.presub))
                st.presub = None
        elif st.kind == "vec" and st.vec_mul is not None and st.vec_div is None and st.vec_sub is None:
            if prev is not None and prev.kind == "mm" and prev.dst == "x" and prev.colscale is None and prev.postsub is None:
                prev.colscale = st.vec_mul
                continue
        out.append(st)
    return out


# --------------------------------------------------------------------------------------------------
# workspace + execution
# --------------------------------------------------------------------------------------------------
class _Workspace:
    """Two stream buffers (x) and two hidden buffers (h) per device, each with the planes a mode needs."""

    def __init__(self):
        self._bufs = {}

    def planes(self, device, name: str, rows: int, width: int, fmt: str) -> torch.Tensor:
        ld = pad4(width)
        key = (device, name, fmt)
        need = rows * ld
        dtype = torch.bfloat16 if fmt == "bf16" else torch.float32
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < need:
            buf = torch.empty(max(need, 1), dtype=dtype, device=device)
            self._bufs[key] = buf
        return buf[:need].view(rows, ld)[:, :width]

    def act(self, device, name: str, rows: int, width: int, mode: str, needs: set) -> Act:
        a = Act(rows, width)
        want_op = "op" in needs
        want_res = "resid" in needs
        want_f32 = "f32" in needs or "final" in needs
        if mode == "fp32":
            if want_op or want_res:
                a.hi = self.planes(device, name, rows, width, "hi")
                a.lo = self.planes(device, name, rows, width, "lo")
            if want_f32:
                a.f32 = self.planes(device, name, rows, width, "f32")
        elif mode == "bf16":
            if want_op:
                a.bf16 = self.planes(device, name, rows, width, "bf16")
            if want_res or want_f32 or "simt" in needs:
                a.f32 = self.planes(device, name, rows, width, "f32")
        else:
            a.f32 = self.planes(device, name, rows, width, "f32")
        if a.f32 is None and a.hi is None and a.bf16 is None:
            a.f32 = self.planes(device, name, rows, width, "f32")
        return a


_workspace = _Workspace()


def _run_steps(steps: List[Step], x: torch.Tensor, mode: str, final_out: torch.Tensor) -> Act:
    """Run the program over one chunk of rows; the final stream value lands in `final_out` (fp32)."""
    dev, rows = x.device, x.shape[0]
    slots = {}
    flip = {"x": 0, "h": 0}
    last_x = max(i for i, st in enumerate(steps) if st.dst == "x")

    def new_act(slot: str, width: int, needs: set, idx: int) -> Act:
        flip[slot] ^= 1
        a = _workspace.act(dev, f"{slot}{flip[slot]}", rows, width, mode, needs)
        if idx == last_x:
            a.f32 = final_out
        return a

    for i, st in enumerate(steps):
        if st.kind == "vec":
            if i == 0:
                src = x
            else:
                cur = slots["x"]
                if cur.f32 is None:
                    raise RuntimeError("internal: vec step needs an fp32 stream plane")
                src = cur.f32
            out = new_act("x", src.shape[1], st.needs, i)
            ops.ingest(src, out, div=st.vec_div, mul=st.vec_mul, sub=st.vec_sub)
            slots["x"] = out
        elif st.kind == "mm":
            a = slots[st.src]
            N, K = st.w.shape
            eng = _engine_for(mode, N, K)
            wmode = "fp32" if mode == "fp32" else ("bf16" if eng == ENGINE_TC_BF16 else "tf32")
            w, w_lo = _operands.get(st.w, wmode)
            if eng == ENGINE_SIMT and mode == "bf16":
                a_use = Act(a.rows, a.width, f32=a.f32) if a.f32 is not None else None
                if a_use is None:
                    raise RuntimeError("internal: SIMT step in bf16 mode needs an fp32 plane")
                a = a_use
            resid = slots["x"] if st.resid else None
            if st.resid and st.dst == "x" and st.src != "x":
                # in place on the stream buffer is safe (each element is read and written by one thread),
                # but the planes the consumer needs may differ: allocate per `needs` on the same buffer
                out = _workspace.act(dev, f"x{flip['x']}", rows, N, mode, st.needs)
                if i == last_x:
                    out.f32 = final_out
            else:
                out = new_act(st.dst, N, st.needs, i)
            ops.linear(eng, a, w, w_lo, N, K, bias=st.bias, relu=st.relu, resid=resid, resid_sign=st.sign,
                       colscale=st.colscale, postsub=st.postsub, out=out)
            slots[st.dst] = out
        elif st.kind == "leaky":
            cur = slots["x"]
            out = new_act("x", cur.width, {"f32"}, i)
            ops.leaky_relu(cur.f32, st.slope, out.f32)
            slots["x"] = out
        elif st.kind == "permute":
            cur = slots["x"]
            out = new_act("x", cur.width, {"f32"}, i)
            ops.permute(cur.f32, st.perm, out.f32)
            slots["x"] = out
        else:
            raise RuntimeError(st.kind)
    return slots["x"]


def _needs_reingest(steps: List[Step]) -> List[Step]:
    """leaky/permute write only an fp32 plane; a following contraction needs operand planes -> insert ingest."""
    out: List[Step] = []
    for st in steps:
        if st.kind == "mm" and st.src == "x" and out and out[-1].kind in ("leaky", "permute"):
            out.append(Step("vec"))
        out.append(st)
    return out


class Program:
    def __init__(self, layers, direction: str):
        steps: List[Step] = []
        seq = list(layers) if direction == "forward" else list(reversed(list(layers)))
        for layer in seq:
            _emit_layer(layer, direction, steps)
        steps = _needs_reingest(_fuse(steps))
        if not steps or steps[0].kind != "vec":
            steps.insert(0, Step("vec"))
        for i, st in enumerate(steps):
            needs = set()
            for nxt in steps[i + 1:]:
                if nxt.kind == "mm":
                    if nxt.src == st.dst:
                        needs.add("op")
                        if min(nxt.w.shape) < TC_MIN_DIM:
                            needs.add("simt")       # SIMT engine reads fp32-accurate planes
                    if nxt.resid and st.dst == "x":
                        needs.add("resid")
                    if nxt.dst == st.dst:
                        break
                elif st.dst == "x":
                    needs.add("f32")
                    break
            st.needs = needs
        self.steps = steps

    def run(self, x: torch.Tensor, mode: Optional[str] = None, chunk_rows: Optional[int] = None,
            out: Optional[torch.Tensor] = None, sink=None) -> Optional[torch.Tensor]:
        """Evaluate the program on x [rows, d].  With `sink`, the final stream value of each chunk is handed
        to `sink(chunk_f32 [r, width], r0, r1)` from a reused workspace buffer instead of being stored."""
        mode = mode or _default_precision
        ops.require_cuda(x, "input")
        if x.dim() != 2:
            raise RuntimeError("usflows_b200: expected a [rows, d] input")
        x = x.contiguous()
        rows = x.shape[0]
        width = self.out_width(x.shape[1])
        chunk = min(chunk_rows or _default_chunk_rows, max(rows, 1))
        if rows == 0:                       # empty batch: nothing to launch
            return None if sink is not None else torch.empty(0, width, dtype=torch.float32, device=x.device)
        if sink is not None:
            for r0 in range(0, rows, chunk):
                r1 = min(rows, r0 + chunk)
                fin = _workspace.planes(x.device, "final", r1 - r0, width, "f32")
                _run_steps(self.steps, x[r0:r1], mode, fin)
                sink(fin, r0, r1)
            return None
        if out is None:
            out = torch.empty(rows, width, dtype=torch.float32, device=x.device)
        for r0 in range(0, rows, chunk):
            r1 = min(rows, r0 + chunk)
            _run_steps(self.steps, x[r0:r1], mode, out[r0:r1])
        return out

    def out_width(self, d_in: int) -> int:
        w = d_in
        for st in self.steps:
            if st.kind == "mm" and st.dst == "x":
                w = st.w.shape[0]
        return w


# --------------------------------------------------------------------------------------------------
# public helpers used by the layer / flow classes
# --------------------------------------------------------------------------------------------------
def _flatten_rows(x: torch.Tensor, event_ndim: int = 1):
    ops.require_cuda(x, "input")
    batch_shape = x.shape[:x.dim() - event_ndim]
    return x.reshape(max(1, math.prod(batch_shape)), -1), batch_shape


def run_layers(layers, direction: str, x: torch.Tensor, mode: Optional[str] = None,
               chunk_rows: Optional[int] = None) -> torch.Tensor:
    x2, batch_shape = _flatten_rows(x)
    with torch.no_grad():
        y = Program(layers, direction).run(x2, mode, chunk_rows)
    return y.reshape(*batch_shape, y.shape[-1]) if x.dim() != 2 else y


def run_mlp(net, x: torch.Tensor, mode: Optional[str] = None) -> torch.Tensor:
    """Plain evaluation of a DenseNN (no mask, no residual) through the contraction kernels."""
    x2, batch_shape = _flatten_rows(x)
    lin = list(net.layers)
    steps = [Step("vec")]
    for j, l in enumerate(lin):
        last = j == len(lin) - 1
        steps.append(Step("mm", src="x" if j == 0 else "h", dst="x" if last else "h",
                          w=l.weight.detach(), bias=l.bias.detach(), relu=not last))
    prog = Program([], "forward")
    prog.steps = steps
    for i, st in enumerate(steps):
        st.needs = {"op"} if i < len(steps) - 1 else set()
    with torch.no_grad():
        y = prog.run(x2, mode)
    return y.reshape(*batch_shape, y.shape[-1]) if x.dim() != 2 else y


def total_ladj(layers) -> float:
    """Sum of the forward log|det J| of all layers -- a model constant for a USFlow (every layer's log-det
    is data independent).  One device->host read per weight version (callers cache)."""
    parts = [l._ladj_device() for l in layers]
    parts = [p for p in parts if p is not None]
    if not parts:
        return 0.0, 0
    tot = torch.stack(parts).double().sum(0).cpu()
    return float(tot[0]), int(tot[1])


def base_log_prob(base, z: torch.Tensor, add_const: float = 0.0) -> torch.Tensor:
    d = math.prod(base.event_shape)
    z2, batch_shape = _flatten_rows(z, len(base.event_shape))
    if z2.shape[1] != d:
        raise RuntimeError("usflows_b200: event shape mismatch in base log_prob")
    loc, scale = base._prepared()
    z2 = z2.contiguous()
    out = torch.empty(z2.shape[0], dtype=torch.float32, device=z2.device)
    ops.base_logprob(Act(z2.shape[0], d, f32=z2), loc, scale, base.base_kind, add_const, out)
    return out.reshape(batch_shape)


def base_sample(base, sample_shape=None) -> torch.Tensor:
    if sample_shape is None:
        sample_shape = []
    shape = [int(s) for s in sample_shape]
    rows = max(1, math.prod(shape))
    d = math.prod(base.event_shape)
    loc, scale = base._prepared()
    out = torch.empty(rows, d, dtype=torch.float32, device=loc.device)
    seed = int(torch.initial_seed())
    base._seed_offset += 1
    ops.base_sample(Act(rows, d, f32=out), loc, scale, base.base_kind, seed, base._seed_offset)
    return out.reshape(*shape, *base.event_shape)


def leaky_relu_ladj(x: torch.Tensor, alpha: float) -> torch.Tensor:
    x2, batch_shape = _flatten_rows(x)
    x2 = x2.contiguous()
    y = torch.empty_like(x2)
    cnt = torch.empty(x2.shape[0], dtype=torch.float32, device=x2.device)
    ops.leaky_relu(x2, alpha, y, cnt)
    out = cnt * math.log(alpha)
    return out.reshape(batch_shape) if x.dim() > 1 else out.reshape(())


def profile_step(fn) -> dict:
    """Run `fn` once with CUDA events around every hot-path kernel launch; returns milliseconds summed per
    kernel class ("linear NxK", "ingest", "base_logprob") plus "_names" (launch order) and "_total"."""
    records = []
    originals = {name: getattr(ops, name) for name in ("linear", "ingest", "base_logprob")}

    def wrap(name, f):
        def inner(*a, **k):
            if name == "linear":
                label = f"linear {a[4]}x{a[5]}"
            else:
                label = name
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            f(*a, **k)
            e1.record()
            records.append((label, e0, e1))
        return inner

    try:
        for name, f in originals.items():
            setattr(ops, name, wrap(name, f))
        torch.cuda.synchronize()
        fn()
        torch.cuda.synchronize()
    finally:
        for name, f in originals.items():
            setattr(ops, name, f)
    out = {}
    for label, e0, e1 in records:
        out[label] = out.get(label, 0.0) + e0.elapsed_time(e1)
    out["_total"] = sum(v for k, v in out.items() if not k.startswith("_"))
    out["_names"] = [r[0] for r in records]
    return out
