"""Hand-written forward + backward of the density pass for the training step (reference: `Flow.fit`, flows.py:195-207:
`loss = -log_prob(x).mean() - log_prior(); loss.backward()` with torch autograd through F.linear / torch.inverse /
tril / triu, transforms.py:1264-1293).

Every contraction of the step runs on the fp16-split tcgen05 engine of libusflows_b200.so:

  forward   the same chain as `Flow.log_prob` (ingest -> affine / conditioner contractions with bias / ReLU / in-place
            coupling update in the epilogue), with every operand-plane activation KEPT for the backward pass
  dX        dY . M       = usf_linear(a = dY planes, w = M^T planes)            (M^T prepared weight-side)
  dW        dY^T . X     = usf_linear(a = dY^T planes, w = X^T planes, split_k)  (the transposed planes come from
            `usf_planes_glue`, which also applies the ReLU-backward mask and accumulates the bias gradient)
  weights   L, U, W = L U, L^-1, U^-1 (batched recursive triangular inverse), W^-1 = U^-1 L^-1 and their derivative
            dW_tot = dW - W^-T dW^-1 W^-T,  dL = tril(dW_tot U^T, -1),  dU = triu(L^T dW_tot)   -- all as usf_linear calls

Gradients are carried UN-NORMALISED (d(-sum log p)) through the batch side so that they stay in the fp16-split range
(values of order 1, absolute resolution 1e-11; an fp16 overflow raises the device flag `overflow`), and are scaled by
1 / global_batch where they are written into `p.grad`.

Scope: flat events, `BlockAffineTransform` over ONE `LUTransform` (directly or as a one-element
`SequentialAffineTransform`), additive `MaskedCoupling` with a `DenseNN` conditioner whose mask splits the features into two
contiguous halves after re-ordering (the checkerboard / channel masks of `USFlow`), `InverseTransform` of such an affine
layer, a leading `ScaleTransform`, Laplace / Normal base.  `supports(flow)` says whether a flow fits; anything else trains
through `training.log_prob_autograd` (same kernels for the contractions, torch autograd for the bookkeeping).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch

from . import ops
from .ops import Act, ENGINE_SIMT, ENGINE_TC_3XF16, pad4


def _planes(rows: int, cols: int, device) -> Act:
    ld = pad4(cols)
    buf = torch.zeros(2, rows, ld, dtype=torch.float16, device=device)
    return Act(rows, cols, h16=buf[0, :, :cols], l16=buf[1, :, :cols])


def _seg(a: Act, c0: int, w: int) -> Act:
    return Act(a.rows, w, h16=a.h16[:, c0:c0 + w], l16=a.l16[:, c0:c0 + w])


def _rows(a: Act, r0: int, n: int) -> Act:
    return Act(n, a.width, h16=a.h16[r0:r0 + n], l16=a.l16[r0:r0 + n])


def _f32(rows: int, cols: int, device) -> torch.Tensor:
    return torch.zeros(rows, pad4(cols, 4), dtype=torch.float32, device=device)[:, :cols]


def _unwrap_lu(block):
    """The single LUTransform behind a BlockAffineTransform, or None."""
    from . import transforms as T
    t = block.block_transform
    if isinstance(t, T.SequentialAffineTransform):
        if len(t.transforms) != 1:
            return None
        t = t.transforms[0]
    return t if isinstance(t, T.LUTransform) else None


def _ops_of(flow) -> Optional[List[tuple]]:
    """Density-direction op list [(kind, object)], or None if a layer is outside the engine's scope."""
    from . import transforms as T
    from .nn import DenseNN
    out = []
    for layer in reversed(flow.layers):
        if isinstance(layer, T.ScaleTransform):
            out.append(("scale", layer))
        elif isinstance(layer, T.InverseTransform) and isinstance(layer.transform, T.BlockAffineTransform):
            lu = _unwrap_lu(layer.transform)
            if lu is None or layer.transform.n_blocks != 1:
                return None
            out.append(("aff_fwd", lu))
        elif isinstance(layer, T.BlockAffineTransform):
            lu = _unwrap_lu(layer)
            if lu is None or layer.n_blocks != 1:
                return None
            out.append(("aff_bwd", lu))
        elif type(layer) is T.MaskedCoupling and isinstance(layer.conditioner, DenseNN) \
                and layer.conditioner.count_params == 1 and len(layer.conditioner.layers) >= 2:
            out.append(("coupling", layer))
        else:
            return None
    return out


def supports(flow) -> bool:
    from .distributions import DistributionModule, Independent, RadialDistribution, _FrozenBase
    ev = tuple(flow._event_shape())
    if len(ev) != 1 or ev[0] % 8 or ev[0] < 32 or getattr(flow, "soft_training", False):
        return False
    base = flow.base_distribution.base_dist if isinstance(flow.base_distribution, Independent) else flow.base_distribution
    if not isinstance(base, DistributionModule) or isinstance(base, (RadialDistribution, _FrozenBase)) or base.base_kind < 0:
        return False                  # (a frozen torch-object base has no gradients to write: autograd route)
    seq = _ops_of(flow)
    if not seq or any(k == "scale" for k, _ in seq[1:]):
        return False
    kinds = [k for k, _ in seq]
    part = None
    for i, (k, obj) in enumerate(seq):
        if k != "coupling":
            continue
        # a coupling sits between two affine layers (its column re-ordering is folded into their matrices)
        if i == 0 or i + 1 >= len(seq) or not kinds[i - 1].startswith("aff") or not kinds[i + 1].startswith("aff"):
            return False
        m = obj.mask.reshape(-1).float().cpu()
        if part is None:
            part = m
        if not (torch.equal(m, part) or torch.equal(m, 1 - part)):
            return False
        h1 = int((part > 0.5).sum())
        if h1 % 8 or (m.numel() - h1) % 8 or h1 == 0 or h1 == m.numel():
            return False
        dims = [l.weight.shape[0] for l in obj.conditioner.layers[:-1]]
        if any(h % 8 or h < 32 for h in dims):
            return False
    return True


class _LU:
    """Weight-side state of one LUTransform for one step: factors, inverses, products, gradient accumulators."""

    def __init__(self, lu, index: int):
        self.lu, self.index = lu, index
        self.need_w = self.need_winv = False
        self.mult = 0                  # net multiplicity of log|det W| in the log-likelihood (see `step`)


class TrainEngine:
    def __init__(self, flow, rows: int):
        if not supports(flow):
            raise NotImplementedError("usflows_b200.train_engine: this flow is outside the engine's scope")
        from .distributions import Independent
        self.flow, self.rows = flow, int(rows)
        self.base = flow.base_distribution.base_dist if isinstance(flow.base_distribution, Independent) else flow.base_distribution
        dev = next(flow.parameters()).device
        self.dev = dev
        d = self.d = int(flow._event_shape()[0])
        M = self.rows
        seq = _ops_of(flow)
        # ---- column re-ordering of the coupled segments (as engine.Program._compression_plan) ----
        self.order = self.inv_order = None
        coup = [obj for k, obj in seq if k == "coupling"]
        if coup:
            part = coup[0].mask.reshape(-1).float().to(dev)
            idx1 = torch.nonzero(part > 0.5).reshape(-1)
            idx0 = torch.nonzero(part <= 0.5).reshape(-1)
            self.part, self.idx1, self.idx0 = part, idx1.to(torch.int32), idx0.to(torch.int32)
            self.h1, self.h0 = int(idx1.numel()), int(idx0.numel())
            order = torch.cat([idx1, idx0])
            self.order = order.to(torch.int32)
            inv = torch.empty_like(order)
            inv[order] = torch.arange(d, device=dev)
            self.inv_order = inv.to(torch.int32)
        # ---- unique LU layers ----
        self.lus: Dict[int, _LU] = {}
        for k, obj in seq:
            if k.startswith("aff") and id(obj) not in self.lus:
                self.lus[id(obj)] = _LU(obj, len(self.lus))
        n_lu = len(self.lus)
        if 2 * n_lu > 32:
            raise NotImplementedError("usflows_b200.train_engine: more than 16 LU layers")
        self.T = torch.zeros(2 * n_lu, d, d, dtype=torch.float32, device=dev)      # [L_0, U_0^T, L_1, U_1^T, ...]
        self.X = torch.zeros_like(self.T)                                          # their inverses
        self.Ttmp = torch.zeros_like(self.T)
        self.unit_mask = sum(1 << (2 * i) for i in range(n_lu))
        # ---- op plan ----
        self.plan: List[dict] = []
        self.flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self.scale_layer = None
        kinds = [k for k, _ in seq]
        for i, (k, obj) in enumerate(seq):
            if k == "scale":
                self.scale_layer = obj
                continue
            if k.startswith("aff"):
                st = self.lus[id(obj)]
                perm_in = i > 0 and kinds[i - 1] == "coupling"
                perm_out = i + 1 < len(seq) and kinds[i + 1] == "coupling"
                fwd = k == "aff_fwd"
                if fwd:
                    st.need_w = True
                    st.mult += 1
                else:
                    st.need_winv = True
                    st.mult -= 1
                self.plan.append(dict(kind="aff", fwd=fwd, lu=st, perm_in=perm_in, perm_out=perm_out,
                                      M=_planes(d, d, dev), MT=_planes(d, d, dev), c=torch.zeros(d, device=dev),
                                      x=None, y=_planes(M, d, dev), xT=_planes(d, M, dev),
                                      dM=_f32(d, d, dev), dc=torch.zeros(d, device=dev)))
            else:
                first = bool(torch.equal(obj.mask.reshape(-1).float().to(dev), self.part))
                lin = list(obj.conditioner.layers)
                dims = [l.weight.shape[0] for l in lin]
                h_in, h_out = (self.h1, self.h0) if first else (self.h0, self.h1)
                in_seg = (0, self.h1) if first else (self.h1, self.h0)
                out_seg = (self.h1, self.h0) if first else (0, self.h1)
                widths_in = [h_in] + dims[:-1]
                widths_out = dims[:-1] + [h_out]
                layers = []
                for j, l in enumerate(lin):
                    n, kk = widths_out[j], widths_in[j]
                    layers.append(dict(lin=l, n=n, k=kk, W=_planes(n, kk, dev), WT=_planes(kk, n, dev),
                                       b=torch.zeros(n, device=dev), dW=_f32(n, kk, dev), db=torch.zeros(n, device=dev),
                                       h=_planes(M, n, dev) if j < len(lin) - 1 else None,       # post-ReLU activation
                                       hT=_planes(n, M, dev) if j < len(lin) - 1 else None,
                                       g=_planes(M, n, dev), gT=_planes(n, M, dev)))
                self.plan.append(dict(kind="coupling", layer=obj, first=first, in_seg=in_seg, out_seg=out_seg, layers=layers,
                                      idx_in=self.idx1 if first else self.idx0, idx_out=self.idx0 if first else self.idx1))
        if not self.plan or self.plan[0]["kind"] != "aff" or self.plan[-1]["kind"] != "aff":
            raise NotImplementedError("usflows_b200.train_engine: the stack must start (data side) with an affine layer")
        # stream buffers: the input of op i is the output of op i - 1 (a coupling updates its input in place)
        self.x0 = _planes(M, d, dev)                 # x / scale
        self.x0T = _planes(d, M, dev)
        prev = self.x0
        for op in self.plan:
            if op["kind"] == "aff":
                op["x"] = prev
                prev = op["y"]
            else:
                op["x"] = prev                       # in place
        self.z = _f32(M, d, dev)                     # latent (fp32) for the base density
        self.lp = torch.zeros(M, device=dev)
        self.g = _planes(M, d, dev)                  # gradient stream (two buffers in turn)
        self.g2 = _planes(M, d, dev)
        self.gT = _planes(d, M, dev)
        # per-LU work buffers (fp32 [d, d]) and operand planes
        for st in self.lus.values():
            st.W = _f32(d, d, dev)
            st.Winv = _f32(d, d, dev)
            st.pL, st.pUT, st.pU, st.pLT = (_planes(d, d, dev) for _ in range(4))
            st.pUinv, st.pLinvT = _planes(d, d, dev), _planes(d, d, dev)
            st.pWinv, st.pWinvT = _planes(d, d, dev), _planes(d, d, dev)
            st.dW = _f32(d, d, dev)
            st.dWinv = _f32(d, d, dev)
            st.pG, st.pTt = _planes(d, d, dev), _planes(d, d, dev)
            st.dWtot = _f32(d, d, dev)
            st.pdWtot, st.pdWtotT = _planes(d, d, dev), _planes(d, d, dev)
            st.dL, st.dU = _f32(d, d, dev), _f32(d, d, dev)
            st.db = torch.zeros(d, device=dev)
            st.tmp = _f32(d, d, dev)
            st.vec = _f32(1, d, dev)
            st.vec2 = _f32(1, d, dev)
            st.gL = torch.zeros(d, d, device=dev)
            st.gU = torch.zeros(d, d, device=dev)
            st.gb = torch.zeros(d, device=dev)
        self.d_loc = torch.zeros(d, device=dev)
        self.d_scale = torch.zeros(d, device=dev)
        self.cs2 = torch.zeros(d, device=dev)        # sum_r dy0 * x (scale gradient)
        pairs = max(1, _num_sms() // 2)
        self._pairs = pairs
        self._side = [torch.cuda.Stream(device=dev) for _ in range(3)] if dev.type == "cuda" else []
        self._side_used = set()

    # ------------------------------------------------------------------------------------------------------------
    def _split_k(self, n_out: int, k_out: int, rows: int) -> int:
        """Number of K pieces of a dW product: the one that minimises (waves of the persistent CTA-pair grid) x (K-slabs per
        piece + the fixed cost of a work item: accumulator drain + reduction into memory, ~4 slabs' worth)."""
        bn = 256 if k_out % 256 == 0 or k_out > 832 else 208
        tiles = -(-n_out // 256) * -(-k_out // bn)
        k_slabs = -(-rows // 64)
        best, best_cost = 1, None
        for s in range(1, max(1, k_slabs // 2) + 1):
            per = -(-k_slabs // s)
            if (s - 1) * per >= k_slabs:
                continue
            cost = -(-tiles * s // self._pairs) * (per + 4)
            if best_cost is None or cost < best_cost:
                best, best_cost = s, cost
        return best

    def _gemm(self, a: Act, w: Act, N: int, K: int, *, bias=None, relu=False, resid=None, sign=1.0, out: Act) -> None:
        ops.linear(ENGINE_TC_3XF16, a, w.h16, w.l16, N, K, bias=bias, relu=relu, resid=resid, resid_sign=sign, out=out,
                   overflow_flag=self.flag)

    # ------------------------------------------------------------------------------------------------------------
    def _prepare_weights(self) -> None:
        """Weight-side forward: factors, inverses, products and the operand planes of every use (once per step)."""
        d = self.d
        for st in self.lus.values():
            lu, i = st.lu, st.index
            ops.lu_assemble(lu.L_raw.detach(), lu.U_raw.detach(), self.T[2 * i], self.T[2 * i + 1], transpose_u=True)
        ops.tri_inverse_batched(self.T, self.X, self.Ttmp, self.unit_mask)         # L^-1 and (U^T)^-1 = (U^-1)^T of every layer
        # the chains of the layers are independent of each other (16 output tiles each on 74 CTA pairs): they run side by
        # side on a few streams (parallel branches of the captured graph); the conditioner operands are prepared on the
        # main stream meanwhile
        for k, st in enumerate(self.lus.values()):
            with self._on_side(k):
                self._prepare_lu(st)
        self._prepare_conditioners()
        self._join_sides()

    def _on_side(self, k: int):
        """Context: a side stream that has waited for everything issued on the current stream so far (CUDA devices;
        a no-op on the emulated backend)."""
        if not self._side:
            import contextlib
            return contextlib.nullcontext()
        s = self._side[k % len(self._side)]
        s.wait_stream(torch.cuda.current_stream(self.dev))
        self._side_used.add(k % len(self._side))
        return torch.cuda.stream(s)

    def _join_sides(self) -> None:
        main = torch.cuda.current_stream(self.dev) if self._side else None
        for k in sorted(self._side_used):
            main.wait_stream(self._side[k])
        self._side_used.clear()

    def _prepare_lu(self, st: _LU) -> None:
        d = self.d
        i = st.index
        L, UT, Linv, UinvT = self.T[2 * i], self.T[2 * i + 1], self.X[2 * i], self.X[2 * i + 1]
        ops.mat_prep(L, out=st.pL, out_t=st.pLT, overflow_flag=self.flag)
        ops.mat_prep(UT, out=st.pUT, out_t=st.pU, overflow_flag=self.flag)
        if st.need_w:                                                            # W = L U      (transforms.py:1281-1283)
            ops.linear(ENGINE_TC_3XF16, st.pL, st.pUT.h16, st.pUT.l16, d, d, out=Act(d, d, f32=st.W), overflow_flag=self.flag)
        if st.need_winv:                                                         # W^-1 = U^-1 L^-1   (transforms.py:1289-1293)
            ops.mat_prep(UinvT, transpose=True, out=st.pUinv, overflow_flag=self.flag)
            ops.mat_prep(Linv, transpose=True, out=st.pLinvT, overflow_flag=self.flag)
            ops.linear(ENGINE_TC_3XF16, st.pUinv, st.pLinvT.h16, st.pLinvT.l16, d, d, out=Act(d, d, f32=st.Winv),
                       overflow_flag=self.flag)
            ops.mat_prep(st.Winv, out=st.pWinv, out_t=st.pWinvT, overflow_flag=self.flag)
        for op in self.plan:
            if op["kind"] == "aff" and op["lu"] is st:
                P = st.W if op["fwd"] else st.Winv
                r = self.order if op["perm_out"] else None
                c = self.order if op["perm_in"] else None
                ops.mat_prep(P, row_idx=r, col_idx=c, out=op["M"], out_t=op["MT"], overflow_flag=self.flag)
                b = st.lu.bias_vector.detach()
                if op["fwd"]:                                                        # y = x W^T + b   (transforms.py:913-934)
                    ops.mat_prep(b.reshape(1, d), col_idx=r, out_f32=op["c"].reshape(1, d))
                else:                                                                # y = (x - b) W^-T = x W^-T - W^-1 b  (:936-962)
                    ops.rowdot(st.Winv, b, -1.0, op["c"], row_idx=r)

    def _prepare_conditioners(self) -> None:
        for op in self.plan:
            if op["kind"] == "coupling":
                n_l = len(op["layers"])
                for j, L in enumerate(op["layers"]):
                    w = L["lin"].weight.detach()
                    bias = L["lin"].bias.detach()
                    ci = op["idx_in"] if j == 0 else None
                    ri = op["idx_out"] if j == n_l - 1 else None
                    ops.mat_prep(w, row_idx=ri, col_idx=ci, out=L["W"], out_t=L["WT"], overflow_flag=self.flag)
                    ops.mat_prep(bias.reshape(1, -1), col_idx=ri, out_f32=L["b"].reshape(1, -1))

    # ------------------------------------------------------------------------------------------------------------
    def _forward(self, x: torch.Tensor) -> None:
        d, M = self.d, self.rows
        scale = None if self.scale_layer is None else self.scale_layer.scale.detach().reshape(-1)
        ops.ingest(x, self.x0, div=scale, overflow_flag=self.flag)                  # x / scale    (transforms.py:116-125)
        for op in self.plan:
            if op["kind"] == "aff":
                # (the last op hands the latent to the base density as fp32; its backward only needs its INPUT planes)
                out = Act(M, d, f32=self.z) if op is self.plan[-1] else op["y"]
                self._gemm(op["x"], op["M"], d, d, bias=op["c"], out=out)
            else:
                cur = op["x"]
                a = _seg(cur, *op["in_seg"])
                n_l = len(op["layers"])
                for j, L in enumerate(op["layers"]):
                    if j < n_l - 1:                                                  # h = relu(a W^T + b)
                        self._gemm(a, L["W"], L["n"], L["k"], bias=L["b"], relu=True, out=L["h"])
                        a = L["h"]
                    else:                                                            # x_out <- x_out - t (transforms.py:292-306)
                        seg = _seg(cur, *op["out_seg"])
                        self._gemm(a, L["W"], L["n"], L["k"], bias=L["b"], resid=seg, sign=-1.0, out=seg)

    # ------------------------------------------------------------------------------------------------------------
    def step(self, x: torch.Tensor, total_rows: int, reducer=None) -> torch.Tensor:
        """Forward + backward on this rank's rows `x` [rows, d] (fp32, CUDA); fills `p.grad` of every parameter with the
        gradient of  -sum_rows log p(x) / total_rows  and returns that loss share (device scalar)."""
        if x.shape != (self.rows, self.d):
            raise RuntimeError("usflows_b200.train_engine: the engine was built for another batch shape")
        d, M = self.d, self.rows
        inv_total = 1.0 / float(total_rows)
        share = M * inv_total
        with torch.no_grad():
            self._prepare_weights()
            self._forward(x.contiguous())
            loc, sc = self.base._prepared()
            # loss value: -sum log p(z) / total + share * (sum of the weight-only log-determinants)
            ops.base_logprob(Act(M, d, f32=self.z), loc, sc, self.base.base_kind, 0.0, self.lp)
            ladj = None
            for st in self.lus.values():
                if st.mult != 0:
                    term = st.lu.U_raw.detach().diagonal().abs().log().sum() * float(-st.mult)
                    ladj = term if ladj is None else ladj + term
            if self.scale_layer is not None:
                term = self.scale_layer.scale.detach().abs().log().sum()
                ladj = term if ladj is None else ladj + term
            loss = -self.lp.sum() * inv_total
            if ladj is not None:
                loss = loss + ladj * share
            # ---- backward -------------------------------------------------------------------------------------
            for t in (self.d_loc, self.d_scale, self.cs2):
                t.zero_()
            g, g_other = self.g, self.g2
            ops.base_backward(self.z, loc, sc, self.base.base_kind, g, self.gT, self.d_loc, self.d_scale)
            gT_ready = True
            self._base_grads(inv_total)
            if reducer is not None:
                for p in self.base.parameters():
                    reducer.push(p)
            for st in self.lus.values():
                st.dW.zero_()
                st.dWinv.zero_()
                st.db.zero_()
                st.uses_left = (1 if st.need_w else 0) + (1 if st.need_winv else 0)
                st.n_uses = sum(1 for op in self.plan if op["kind"] == "aff" and op["lu"] is st)
            for op in reversed(self.plan):
                if op["kind"] == "aff":
                    st = op["lu"]
                    op["dc"].zero_()
                    if not gT_ready:                      # dY^T planes + bias gradient in one pass
                        ops.planes_glue(g, rows=M, n=d, t=self.gT, colsum=op["dc"])
                    else:
                        ops.planes_glue(g, rows=M, n=d, colsum=op["dc"])
                    ops.planes_glue(op["x"], rows=M, n=d, t=op["xT"])            # X^T planes of the saved input
                    ops.linear_splitk(ENGINE_TC_3XF16, self.gT, op["xT"], d, M, op["dM"], self._split_k(d, d, M))
                    self._gemm(g, op["MT"], d, d, out=g_other)                  # dX = dY . M
                    g, g_other = g_other, g
                    gT_ready = False
                    self._scatter_affine_grads(op, inv_total)
                    st.n_uses -= 1
                    if st.n_uses == 0:               # weight-side derivative of this layer: beside the next block's batch work
                        with self._on_side(st.index):
                            self._lu_backward(st, inv_total, share)
                            if reducer is not None:
                                for p in (st.lu.L_raw, st.lu.U_raw, st.lu.bias_vector):
                                    reducer.push(p)
                else:
                    self._coupling_backward(op, g, inv_total)
                    gT_ready = False
                    if reducer is not None:
                        for p in op["layer"].conditioner.parameters():
                            reducer.push(p)
            # g = d loss / d (x / scale): gradient of the scale layer (transforms.py:116-144)
            if self.scale_layer is not None:
                ops.planes_glue(g, rows=M, n=d, mul=x, colsum2=self.cs2)
                s = self.scale_layer.scale.detach().reshape(-1)
                gs = (-self.cs2 / (s * s)) * inv_total + share / s
                self._set_grad(self.scale_layer.scale, gs.reshape(self.scale_layer.scale.shape))
                if reducer is not None:
                    reducer.push(self.scale_layer.scale)
            self._join_sides()
        return loss

    # ------------------------------------------------------------------------------------------------------------
    def bind_grad_buffers(self, view_of) -> None:
        """Let the large gradients (LU factors, conditioner weights) be written straight into caller-owned buffers
        (`view_of(param)` -> tensor of the parameter's shape: the slices of the all-reduce buckets), so that the exchange
        needs no copy in or out."""
        for st in self.lus.values():
            st.gL, st.gU, st.gb = view_of(st.lu.L_raw), view_of(st.lu.U_raw), view_of(st.lu.bias_vector)
        for op in self.plan:
            if op["kind"] == "coupling":
                for L in op["layers"]:
                    L["gw"], L["gb"] = view_of(L["lin"].weight), view_of(L["lin"].bias)
                    L["gw"].zero_()                       # (only the gathered rows / columns are rewritten each step)
                    L["gb"].zero_()

    def _set_grad(self, p: torch.nn.Parameter, g: torch.Tensor) -> None:
        if p.grad is None or p.grad.shape != g.shape:
            p.grad = g                    # persistent buffers of the engine: no copy
        elif p.grad.data_ptr() != g.data_ptr():
            p.grad.copy_(g)

    def _base_grads(self, inv_total: float) -> None:
        b = self.base
        self._set_grad(b.loc, (self.d_loc * inv_total).reshape(b.loc.shape))
        raw = b.scale_unconstrained.detach()
        gs = self.d_scale * inv_total
        if raw.dim() == 0:                                # scalar scale expanded over the event (distributions.py:228-232)
            gr = (gs * torch.sigmoid(raw)).sum().reshape(raw.shape)
        else:
            gr = (gs.reshape(raw.shape) * torch.sigmoid(raw))
        self._set_grad(b.scale_unconstrained, gr)

    def _scatter_affine_grads(self, op: dict, inv_total: float) -> None:
        """dM (in the use's row / column order) and dc -> the layer's dW / dW^-1 / db accumulators (true order).  The
        1 / global_batch normalisation is applied HERE: the un-normalised sums over the batch rows are fp32 (split-K
        output), everything weight-side after this point runs on fp16-split planes again and must stay in their range."""
        st, d = op["lu"], self.d
        r = self.inv_order if op["perm_out"] else None
        c = self.inv_order if op["perm_in"] else None
        ops.mat_prep(op["dM"], row_idx=r, col_idx=c, scale=inv_total, out_f32=st.tmp)
        if op["fwd"]:
            st.dW.add_(st.tmp)
            ops.mat_prep(op["dc"].reshape(1, d), col_idx=r, scale=inv_total, out_f32=st.vec)   # b' = b[order] -> db = dc'[inv_order]
            st.db.add_(st.vec.reshape(-1))
        else:
            st.dWinv.add_(st.tmp)
            # c = -W^-1 b (rows in the use's order):  dW^-1 -= dc (x) b ,  db -= dc . W^-1
            ops.mat_prep(op["dc"].reshape(1, d), col_idx=r, scale=inv_total, out_f32=st.vec)   # dc in the true row order
            dcv = st.vec.reshape(-1)
            ops.rank1(st.dWinv, dcv, st.lu.bias_vector.detach(), -1.0)
            ops.colcomb(st.Winv, dcv, -1.0, st.db)

    def _lu_backward(self, st: _LU, inv_total: float, share: float) -> None:
        """dW, dW^-1 -> dL_raw, dU_raw (transforms.py:1271-1293 differentiated; masks of :1209-1213)."""
        d = self.d
        E = ENGINE_TC_3XF16
        if st.need_winv:
            ops.mat_prep(st.dWinv, out=st.pG, overflow_flag=self.flag)
            ops.linear(E, st.pWinv, st.pG.h16, st.pG.l16, d, d, out=st.pTt, overflow_flag=self.flag)     # T^T = W^-1 G^T
            # dW_tot = dW - W^-T (G W^-T)
            ops.linear(E, st.pWinvT, st.pTt.h16, st.pTt.l16, d, d, resid=Act(d, d, f32=st.dW), resid_sign=-1.0,
                       out=Act(d, d, f32=st.dWtot), overflow_flag=self.flag)
            src = st.dWtot
        else:
            src = st.dW
        ops.mat_prep(src, out=st.pdWtot, out_t=st.pdWtotT, overflow_flag=self.flag)
        ops.linear(E, st.pdWtot, st.pU.h16, st.pU.l16, d, d, out=Act(d, d, f32=st.dL), overflow_flag=self.flag)     # dW_tot U^T
        ops.linear(E, st.pLT, st.pdWtotT.h16, st.pdWtotT.l16, d, d, out=Act(d, d, f32=st.dU), overflow_flag=self.flag)  # L^T dW_tot
        ops.tri_mask(st.dL, 0, 1.0, st.gL)
        # log-likelihood term: -mult * share * sum log|U_kk|
        ops.tri_mask(st.dU, 1, 1.0, st.gU, diag_src=st.lu.U_raw.detach(), coef=-st.mult * share)
        st.gb.copy_(st.db)
        self._set_grad(st.lu.L_raw, st.gL)
        self._set_grad(st.lu.U_raw, st.gU)
        self._set_grad(st.lu.bias_vector, st.gb)

    def _coupling_backward(self, op: dict, g: Act, inv_total: float) -> None:
        """x_out' = x_out - t(x_in): d x_out unchanged, d x_in += dt . dt/dx_in with dt = -d x_out'; parameter gradients of
        the conditioner (pyro.nn.DenseNN restated: Linear / ReLU stack)."""
        M = self.rows
        layers = op["layers"]
        n_l = len(layers)
        cur = op["x"]
        for L in layers:
            L["db"].zero_()
        # gradient wrt the conditioner output: -(gradient of the updated segment)
        last = layers[-1]
        gy = last["g"]
        ops.planes_glue(_seg(g, *op["out_seg"]), rows=M, n=last["n"], sign=-1.0, out=gy, t=last["gT"], colsum=last["db"])
        for j in range(n_l - 1, -1, -1):
            L = layers[j]
            a_in = layers[j - 1]["h"] if j > 0 else _seg(cur, *op["in_seg"])
            a_inT = layers[j - 1]["hT"] if j > 0 else None
            if j > 0:
                ops.planes_glue(a_in, rows=M, n=L["k"], t=a_inT)
            else:                                     # rows in_seg of the transposed stream planes
                a_inT = self._stream_T(cur, op["in_seg"])
            ops.linear_splitk(ENGINE_TC_3XF16, L["gT"], a_inT, L["k"], M, L["dW"], self._split_k(L["n"], L["k"], M))
            if j > 0:                                 # dh = (dy . W) * (h > 0)
                prev = layers[j - 1]
                self._gemm(L["g"], L["WT"], L["k"], L["n"], out=prev["g"])
                ops.planes_glue(prev["g"], rows=M, n=prev["n"], mask_h=prev["h"].h16, out=prev["g"], t=prev["gT"],
                                colsum=prev["db"])
            else:                                     # d x_in += dy . W1   (in place on the gradient stream segment)
                seg = _seg(g, *op["in_seg"])
                self._gemm(L["g"], L["WT"], L["k"], L["n"], resid=seg, sign=1.0, out=seg)
        # parameter gradients (scatter the gathered rows / columns back; everything else of the parameter stays zero)
        for j, L in enumerate(layers):
            lin = L["lin"]
            w = lin.weight
            if "gw" not in L:
                L["gw"] = torch.zeros_like(w)
                L["gb"] = torch.zeros_like(lin.bias)
            if j == 0 and j == n_l - 1:
                raise NotImplementedError
            if j == 0:
                L["gw"][:, op["idx_in"].long()] = L["dW"] * inv_total
                L["gb"].copy_(L["db"]).mul_(inv_total)
            elif j == n_l - 1:
                L["gw"][op["idx_out"].long(), :] = L["dW"] * inv_total
                L["gb"][op["idx_out"].long()] = L["db"] * inv_total
            else:
                torch.mul(L["dW"], inv_total, out=L["gw"])
                L["gb"].copy_(L["db"]).mul_(inv_total)
            self._set_grad(w, L["gw"])
            self._set_grad(lin.bias, L["gb"])

    def _stream_T(self, cur: Act, seg) -> Act:
        """Transposed planes of a column segment of a saved stream activation (built on demand, one buffer)."""
        if not hasattr(self, "_sT"):
            self._sT = _planes(self.d, self.rows, self.dev)
        c0, w = seg
        t = _rows(self._sT, 0, w)
        ops.planes_glue(_seg(cur, c0, w), rows=self.rows, n=w, t=t)
        return t


def _num_sms() -> int:
    import ctypes
    from . import _lib
    try:
        sm = ctypes.c_int(0)
        _lib.check(_lib.load().usf_device_info(ctypes.byref(sm), None, None, None))
        return max(2, sm.value)
    except Exception:                                  # noqa: BLE001  (CPU tests with the emulated backend)
        return 148
