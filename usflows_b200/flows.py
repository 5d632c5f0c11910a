"""`Flow` and `USFlow` -- drop-in mirrors of the reference's model API (src/usflows/flows.py:22-605).

Same constructor signatures, attributes (`layers`, `trainable_layers`, `base_distribution`, `device`,
`export`), state-dict layout and methods (`log_prob`, `sample`, `_forward`, `backward`, `forward`, `to`,
`is_feasible`, `add_jitter`, `log_prior`).  Evaluation runs through `engine.Program`: fused sm_100a kernels
behind the C ABI, prepared weights cached per weight version, the total log|det J| (a model constant for
every USFlow layer) computed once per weight version.  CUDA tensors only.
"""
from __future__ import annotations

import math
import threading
from typing import Any, Dict, Iterable, List, Literal, Optional, Type

import torch

from . import engine, image_engine, ops
from .distributions import DistributionModule, Independent, _FrozenBase
from .transforms import MaskedAffineCoupling
from .transforms import (BaseTransform, Bijective1x1Conv2d, BlockAffineTransform, HouseholderTransform, InverseTransform, LUTransform,
                         MaskedCoupling, ScaleTransform, SequentialAffineTransform)


SMALL_BATCH_GRAPH_ROWS = 4096   # `log_prob` on at most this many rows replays one captured CUDA graph (0 = off)
SMALL_BATCH_GRAPH_IMAGES = 1024 # the same for image-shaped events (one chunk; ~15 us of host time per launch otherwise)
USE_C_PLAN = True             # device-resident `log_prob` / `backward` / `_forward`: ONE C call per chunk (usf_flow_logprob /
                              # usf_flow_apply on a library-owned plan) where the program is contractions only
HOST_CUDA_GRAPHS = True       # `log_prob_host`: replay one captured CUDA graph per full-size chunk
HOST_CHUNK_ROWS = 16384      # rows per H2D copy / kernel batch of `log_prob_host` (copy i+1 overlaps compute i)
HOST_CHUNK_UNITS = (1, 1, 2, 3)   # `log_prob_host` chunk sizes in wave-aligned units (the last entry repeats).  Only the first
                             # copy is exposed, later chunks may grow as fast as the copy stays ahead of the kernels: a C2
                             # row copies 1.4x faster than it computes, so chunk k may hold ~1.4x the rows of chunk k-1.
                             # Measured on C2 (65 536 rows, B200): uniform 7.11 ms, (1, 2, 2, ..) 6.68 ms, (1, 2, 4, ..) 7.0 ms
HOST_STAGE_BYTES = 1 << 30   # the whole batch is staged on the device when it fits (copies run back to back); else a 2-buffer ring
HOST_MAX_STAGE_SLICES = 8    # ... and when it needs at most this many chunks (one captured graph + result buffer per slice)


_capture_lock = threading.Lock()     # one graph capture at a time per process (the launch-mode switch below is global)
_capture_streams: Dict[int, Any] = {}


def capture_stream(device) -> "torch.cuda.Stream":
    """The side stream graph captures of `device` run on.  torch.cuda.graph's default is ONE process-wide stream created on
    whichever device captured first; a capture for another device would then launch on the wrong device (peer access, or an
    illegal address), which is what the row-shard driver (parallel.py) hit with two GPUs in one process."""
    device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    s = _capture_streams.get(idx)
    if s is None:
        s = _capture_streams[idx] = torch.cuda.Stream(device=idx)
    return s


class _plain_stream_order:
    """Kernels captured into a CUDA graph are launched without programmatic dependent launch: inside a graph the
    programmatic edges bought nothing and cost 4% on the chunked host path (measured); on a plain stream they hide the
    launch latency and the next kernel's prologue (-1.7% per step).  Holds the process-wide capture lock: the row-shard
    driver (parallel.py) warms several devices up from several threads.

    Also keeps the cyclic garbage collector off for the duration of the capture: a collection that starts inside it may
    finalise CUDA graphs and plans of an earlier weight version, and `cudaGraphExecDestroy` / `cudaFree` issued by the
    capturing thread invalidate a thread-local capture ("operation not permitted when stream is capturing"; torch >= 2.9
    no longer collects before a capture by itself).  The garbage is collected after the capture instead."""

    def __enter__(self):
        import gc
        from . import _lib
        _capture_lock.acquire()
        self._gc_was_on = gc.isenabled()
        gc.disable()
        _lib.check(_lib.load().usf_debug_set_pdl(0))

    def __exit__(self, *exc):
        import gc
        from . import _lib
        try:
            _lib.check(_lib.load().usf_debug_set_pdl(1))
        finally:
            if self._gc_was_on:
                gc.enable()
            _capture_lock.release()
        return False


def _wave_aligned_chunk_rows(prog, device) -> int:
    """Rows per H2D copy / kernel batch of `log_prob_host`: the smallest multiple of 256-row tile rows >= HOST_CHUNK_ROWS / 2
    for which the widest layer's tile count is a whole number of waves of the persistent CTA-pair kernel (one tile per
    SM pair and wave).  A small first copy keeps the pipeline fill short (the copy of chunk i+1 hides under the kernels
    of chunk i); whole waves keep the tail of every launch full."""
    import ctypes
    from . import _lib
    if getattr(prog, "small", None) is not None:                 # one launch per chunk: copy granularity only
        return max(HOST_CHUNK_ROWS, (32 << 20) // (4 * prog.small["d"]))
    widest = max([st.N for st in prog.steps if st.kind == "mm"] or [0])
    if widest < 32:
        return HOST_CHUNK_ROWS
    sm = ctypes.c_int(0)
    _lib.check(_lib.load().usf_device_info(ctypes.byref(sm), None, None, None))
    pairs = max(1, sm.value // 2)
    n_blocks = -(-widest // 256) if widest % 256 == 0 else -(-widest // 208)
    unit = pairs // math.gcd(pairs, n_blocks)                 # tile rows (of 256 rows) per whole number of waves
    k = max(1, -(-(HOST_CHUNK_ROWS // 2) // (256 * unit)))
    return 256 * unit * k


class Flow(torch.nn.Module):
    """Base flow: a list of bijective layers over a base distribution (flows.py:22-292)."""

    export_modes = Literal["log_prob", "sample", "forward", "backward"]
    export: str = "log_prob"
    device = "cpu"
    _context_needs_soft_training = False

    def __init__(self, base_distribution, layers, soft_training: bool = False, training_noise_prior=None,
                 device: str = "cpu", precision: Optional[str] = None, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        if training_noise_prior is None:                           # flows.py:79-80
            training_noise_prior = torch.distributions.Uniform(0, 1e-6)
        self.soft_training = soft_training
        self.training_noise_prior = training_noise_prior
        self.layers = layers
        self.trainable_layers = torch.nn.ModuleList([l for l in layers if isinstance(l, torch.nn.Module)])
        if isinstance(base_distribution, torch.distributions.Distribution):      # a plain torch Laplace / Normal object
            base_distribution = _FrozenBase(base_distribution)
        self.base_distribution = base_distribution
        # ConditionalDenseNN conditioners tell "no context" (its context layer is skipped: `backward` / `_forward`,
        # flows.py:45-67) from the zero context that a soft-training USFlow substitutes in `log_prob` / `sample`
        # (flows.py:559-565, 580-590: the context layer's bias stays) -- two launch programs per direction
        self._cond_dense_nets = []
        self._layer_route = False             # a conditioner without a fused lowering (networks.BottleneckConv)
        for l in layers:
            while isinstance(l, InverseTransform):
                l = l.transform
            if isinstance(l, MaskedCoupling) and hasattr(l.conditioner, "zero_context_default"):
                self._cond_dense_nets.append(l.conditioner)
            if isinstance(l, MaskedCoupling) and getattr(l.conditioner, "layer_route_only", False):
                self._layer_route = True
        self._zero_ctx = bool(soft_training and self._context_needs_soft_training)
        self.precision = precision
        self.to(device)
        self.device = device
        batch_shape = self.base_distribution.batch_shape
        if len(batch_shape) > 0:                                   # flows.py:97-101
            self.base_distribution = Independent(self.base_distribution, len(batch_shape))
        self._programs: Dict[str, Any] = {}
        ev = tuple(self._event_shape())
        if len(ev) == 3:                    # a Bijective1x1Conv2d's log|det| counts the pixels of the event (its own
            for l in layers:                # forward sees them on the tensor; the launch program needs them up front)
                while isinstance(l, InverseTransform):
                    l = l.transform
                if isinstance(l, Bijective1x1Conv2d) and l.n_blocks is None:
                    l.n_blocks = ev[1] * ev[2]

    # -- reference API ---------------------------------------------------------------------------
    def simplify(self) -> "Flow":
        """The same flow with every LU / Householder / sequential affine layer replaced by its plain matrix form
        (`PlaneBijectiveLinearTransform`, `Bijective1x1Conv2d` for image-shaped events): flows.py:600-606.  The result
        stays on this flow's device and keeps its precision mode (the reference's lands on its default device)."""
        return Flow(self.base_distribution, [l.simplify() for l in self.layers], device=self.device,
                    precision=self.precision)

    def forward(self, x: torch.Tensor):
        """Export-mode dispatch (flows.py:30-43)."""
        if self.export == "log_prob":
            return self.log_prob(x)
        if self.export == "sample":
            return self.sample()
        if self.export == "forward":
            return self._forward(x)
        if self.export == "backward":
            return self.backward(x)
        raise ValueError(f"Unknown export mode {self.export}")

    def reference_module(self, export_mode: str = "log_prob") -> torch.nn.Module:
        """Frozen pure-PyTorch module with the reference's semantics of this flow (for ONNX export / inspection on any
        device; never used by `log_prob` / `sample`, which run on the CUDA kernels only)."""
        from .export import ReferenceSemantics
        return ReferenceSemantics(self, export_mode)

    def to_onnx(self, path: str, export_mode: str = "log_prob", **export_kwargs) -> None:
        """Saves the model as an ONNX file (flows.py:212-223); `export_mode` as `Flow.export`."""
        from .export import to_onnx
        to_onnx(self, path, export_mode, **export_kwargs)

    def _forward(self, x: torch.Tensor) -> torch.Tensor:
        """latent -> data through every layer's `forward` (flows.py:45-55)."""
        return self._run("forward", x)

    def backward(self, x: torch.Tensor) -> torch.Tensor:
        """data -> latent through every layer's `backward` in reverse order (flows.py:57-67)."""
        return self._run("backward", x)

    def log_prob(self, x: torch.Tensor, context: Optional[torch.Tensor] = None) -> torch.Tensor:
        """log p(x) = base.log_prob(z) - sum_k log|det J_k|  (flows.py:225-245)."""
        if self.soft_training:
            self._require_conditional()
        elif self._context_needs_soft_training:
            context = None                    # USFlow.log_prob drops the context unless soft_training (flows.py:559-567)
        if context is not None or self._layer_route:
            return self._with_context("log_prob", x, context)
        with ops.on_device(x):                # no context = context 0 (flows.py:559-565): the conditional conditioners
            return self._log_prob(x)          # are lowered without their context input, which is exact

    def _require_conditional(self) -> None:
        """Soft training hands a context to every coupling's conditioner (transforms.py:284-289), which only the
        conditional networks accept: the reference fails with a TypeError on the first call otherwise."""
        for l in self.layers:
            while isinstance(l, InverseTransform):
                l = l.transform
            if isinstance(l, MaskedCoupling) and not getattr(l.conditioner, "context_channels", 0):
                raise TypeError(f"soft_training passes a context to the conditioner; {type(l.conditioner).__name__} takes "
                                "none (use ConditionalDenseNN / CondConvNet / CondConvNet2D)")

    def _with_context(self, what: str, x: torch.Tensor, context: torch.Tensor) -> torch.Tensor:
        """Evaluation with an explicit context (flows.py:235-238, 257-263): the context channel of the conditional
        conditioners is data, so this runs the layer-by-layer route of the training pass (training.py: every contraction
        on the library's kernels, no gradient bookkeeping)."""
        from . import training
        ops.require_cuda(x, "input")
        ev = len(self._event_shape())
        batch_shape = x.shape[:x.dim() - ev]
        x2 = x.reshape(-1, *x.shape[x.dim() - ev:])
        if context is not None:
            context = torch.as_tensor(context, dtype=torch.float32, device=x.device)
        with torch.no_grad(), ops.on_device(x):
            if what == "log_prob":
                return training.log_prob_autograd(self, x2, context).reshape(batch_shape)
            return training.apply_autograd(self, x2, what, context).reshape(*batch_shape, *x2.shape[1:])

    def _log_prob(self, x: torch.Tensor) -> torch.Tensor:
        prog, ladj = self._program("backward", self._zero_ctx)
        base = self._base_module()
        base._prepared()                  # parameter-side work happens here, outside any graph capture
        x2, batch_shape = engine._flatten_rows(x, len(self._event_shape()))
        d = x2.shape[1]
        rows = x2.shape[0]
        if isinstance(prog, image_engine.ImageProgram):      # image-shaped event: the latent arrives channels-last
            if x2.is_cuda and 0 < rows <= SMALL_BATCH_GRAPH_IMAGES and not prog.force_fallback \
                    and not torch.cuda.is_current_stream_capturing():
                lp = self._log_prob_small_batch(prog, ladj, x2)     # ~100 launches per step: one graph replay instead
                if lp is not None:
                    return lp.reshape(batch_shape)
            out = torch.empty(rows, dtype=torch.float32, device=x2.device)
            perm = self._channels_last_perm(x2.device)
            with torch.no_grad():
                prog.run(x2, sink=lambda z, r0, r1: base._density_into(ops.Act(r1 - r0, d, f32=z), -ladj, out[r0:r1],
                                                                       perm=perm))
            return out.reshape(batch_shape)
        if x2.is_cuda and 0 < rows <= SMALL_BATCH_GRAPH_ROWS and getattr(prog, "small", None) is None \
                and not prog.force_fallback and not prog.has_row_ladj and not torch.cuda.is_current_stream_capturing():
            lp = self._log_prob_small_batch(prog, ladj, x2)
            if lp is not None:
                return lp.reshape(batch_shape)
        if USE_C_PLAN and x2.is_cuda and rows > 0 and prog.plan_able() and getattr(base, "base_kind", -1) >= 0 \
                and not torch.cuda.is_current_stream_capturing():
            return self._log_prob_plan(prog, base, ladj, x2).reshape(batch_shape)
        out = torch.empty(rows, dtype=torch.float32, device=x2.device)

        def sink(z_chunk, r0, r1):
            base._density_into(ops.Act(r1 - r0, d, f32=z_chunk), -ladj, out[r0:r1])

        with torch.no_grad():
            if prog.has_row_ladj:          # affine couplings: per-row log-determinants next to the model constant
                row_ladj = torch.zeros(rows, dtype=torch.float32, device=x2.device)
                prog.run(x2, sink=sink, ladj_rows=row_ladj)
                if rows:
                    ops.sub_rows(out, row_ladj)
            else:
                prog.run(x2, sink=sink)
        return out.reshape(batch_shape)

    @staticmethod
    def _plan_chunks(rows: int):
        """(chunk rows, plan capacity) as `Program.run` would cut the batch; the capacity is rounded up so that a few plans
        serve every batch size."""
        cap = engine._default_chunk_rows
        n_chunks = (rows + cap - 1) // cap
        chunk = min(rows, ((rows + n_chunks - 1) // n_chunks + 255) // 256 * 256)
        size = 4096
        while size < chunk:
            size *= 2
        return chunk, min(size, max(cap, chunk))

    def _log_prob_plan(self, prog, base, ladj: float, x2: torch.Tensor) -> torch.Tensor:
        """`log_prob` through the whole-stack C entry: one `usf_flow_logprob` call per chunk (ingest, every contraction and
        the base density are launched by the library).  Same kernels, same order, same bits as the launch-by-launch route."""
        rows, d = x2.shape
        x2 = x2.contiguous()
        chunk, capacity = self._plan_chunks(rows)
        with torch.no_grad():
            plan = prog.c_plan(d, capacity, base, -ladj)
            out = torch.empty(rows, dtype=torch.float32, device=x2.device)
            starts = list(range(0, rows, chunk))
            flags = torch.zeros(len(starts), dtype=torch.int32, device=x2.device) if prog.mode == "fp32" else None
            for ci, r0 in enumerate(starts):
                r1 = min(rows, r0 + chunk)
                plan.log_prob(x2[r0:r1], out[r0:r1], None if flags is None else flags[ci:ci + 1])
            if flags is not None:                       # chunks that left the fp16 range: tf32-split engine, launch by launch
                for ci in torch.nonzero(flags).reshape(-1).tolist():
                    r0 = starts[ci]
                    r1 = min(rows, r0 + chunk)
                    prog._fallback().run(x2[r0:r1], sink=lambda z, a, b, r0=r0: base._density_into(
                        ops.Act(b - a, z.shape[1], f32=z), -ladj, out[r0 + a:r0 + b]))
        return out

    def _log_prob_small_batch(self, prog, ladj: float, x2: torch.Tensor):
        """Small batches are bound by the host-side cost of ~23 kernel launches (~40 us each through ctypes + tensor-map
        encoding), not by the kernels: the launch program + base density of a given row count is captured ONCE into a CUDA
        graph over fixed staging buffers and replayed (input copied in, result cloned out).  Returns None when a value
        left the fp16 range (the caller then takes the launch-by-launch route, which handles the tf32-split re-run)."""
        # the cache lives ON the Program (one Program per weight version and mode): graphs of a replaced weight version die
        # with their Program instead of surviving under a recycled id() with dangling operand pointers
        cache = prog.__dict__.setdefault("_lp_graphs", {})
        rows, d = x2.shape
        dev = x2.device
        key = (rows, str(dev))
        ent = cache.get(key)
        if ent is None or ent["gen"] != engine._workspace.generation:
            base = self._base_module()
            base._prepared()                  # parameter-side work happens here, outside any graph capture
            xin = torch.empty(rows, d, dtype=torch.float32, device=dev)
            out = torch.empty(rows, dtype=torch.float32, device=dev)
            flag = torch.zeros(1, dtype=torch.int32, device=dev) if getattr(prog, "uses_range_flag", prog.mode == "fp32") else None
            width = prog.out_width(d)
            if isinstance(prog, image_engine.ImageProgram):
                fin = image_engine._dense(dev, "img_final_small", rows, d)
                perm = self._channels_last_perm(dev)

                def body():
                    prog._run_chunk(xin, fin, False, flag)
                    base._density_into(ops.Act(rows, d, f32=fin), -ladj, out, perm=perm)
            else:
                fin = engine._workspace.planes(dev, "final_small", rows, width, "f32")

                def body():
                    prog._run_chunk(xin, fin, flag)
                    base._density_into(ops.Act(rows, width, f32=fin), -ladj, out)
            with torch.no_grad():
                xin.copy_(x2)
                body()                                        # eager pass: sizes every workspace buffer before capture
                torch.cuda.synchronize(dev)
                graph = torch.cuda.CUDAGraph()
                with _plain_stream_order(), torch.cuda.graph(graph, stream=capture_stream(dev), capture_error_mode="thread_local"):
                    body()
            if len(cache) >= 8:                               # a few row counts per weight version; drop the oldest
                cache.pop(next(iter(cache)))
            ent = cache[key] = dict(graph=graph, x=xin, out=out, flag=flag, gen=engine._workspace.generation)
        with torch.no_grad():
            ent["x"].copy_(x2)
            if ent["flag"] is not None:
                ent["flag"].zero_()
            ent["graph"].replay()
            if ent["flag"] is not None and int(ent["flag"].item()) != 0:
                return None
            return ent["out"].clone()

    def _chunk_graph(self, prog, slot: int, buf: torch.Tensor, d: int) -> dict:
        """CUDA graph of the launch program of one full-size chunk reading host-staging buffer `slot`."""
        cache = prog.__dict__.setdefault("_host_graphs", {})     # on the Program: see _log_prob_small_batch
        key = (slot, buf.data_ptr(), tuple(buf.shape))
        ent = cache.get(key)
        if ent is not None and ent["gen"] == engine._workspace.generation:
            return ent
        dev = buf.device
        rows, width = buf.shape[0], prog.out_width(d)
        flag = torch.zeros(1, dtype=torch.int32, device=dev) if prog.mode == "fp32" else None
        fin = engine._workspace.planes(dev, f"final_host{slot}", rows, width, "f32")
        prog._run_chunk(buf, fin, flag)                   # eager pass: sizes every workspace buffer before capture
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with _plain_stream_order(), torch.cuda.graph(graph, stream=capture_stream(dev), capture_error_mode="thread_local"):
            prog._run_chunk(buf, fin, flag)
        ent = dict(graph=graph, fin=fin, flag=flag, gen=engine._workspace.generation)
        for k in [k for k in cache if k[0] == slot and (k[1] != key[1] or cache[k]["gen"] != ent["gen"])]:
            del cache[k]                                  # graphs of a replaced staging buffer / of stale workspaces
        while sum(1 for k in cache if k[0] == slot) >= 6:  # a few chunk sizes per staging buffer; drop the oldest
            del cache[next(k for k in cache if k[0] == slot)]
        cache[key] = ent
        return ent

    def log_prob_host(self, x_host: torch.Tensor, out_host: Optional[torch.Tensor] = None,
                      chunk_rows: Optional[int] = None) -> torch.Tensor:
        """`log_prob` for rows living in HOST memory (pinned for full copy speed): the batch is streamed
        through the device in chunks -- H2D copy of chunk i+1 on a side stream overlaps the kernels of
        chunk i -- and the log-probs are copied back into `out_host`.  Returns `out_host`."""
        if x_host.is_cuda:
            raise RuntimeError("log_prob_host expects a host tensor; use log_prob for device tensors")
        with ops.on_device(next(self.parameters())):
            return self._log_prob_host(x_host, out_host, chunk_rows)

    def _log_prob_host(self, x_host, out_host, chunk_rows):
        dev = next(self.parameters()).device
        if self._layer_route:                 # no launch program to overlap the copies with: plain chunks
            out = torch.empty(x_host.shape[0], dtype=torch.float32) if out_host is None else out_host
            step = max(1, chunk_rows or 8192)
            for r0 in range(0, x_host.shape[0], step):
                out[r0:r0 + step].copy_(self.log_prob(x_host[r0:r0 + step].to(dev)))
            return out
        prog, ladj = self._program("backward", self._zero_ctx)
        base = self._base_module()
        base._prepared()                  # parameter-side work happens here, outside any graph capture
        x2 = x_host.reshape(-1, math.prod(self._event_shape()))
        rows, d = x2.shape
        if out_host is None:
            out_host = torch.empty(rows, dtype=torch.float32, pin_memory=True)
        if isinstance(prog, image_engine.ImageProgram):
            # image-shaped events: two staging buffers in turn, the H2D copy of chunk i + 1 on the copy stream under the
            # kernels of chunk i, one D2H copy of all log-probs at the end
            step = max(1, chunk_rows or HOST_CHUNK_ROWS)
            main = torch.cuda.current_stream(dev)
            if getattr(self, "_img_stage", None) is None or self._img_stage.shape[1:] != (min(step, max(rows, 1)), d) \
                    or self._img_stage.device != dev:
                self._img_stage = torch.empty(2, min(step, max(rows, 1)), d, dtype=torch.float32, device=dev)
                self._copy_stream = torch.cuda.Stream(device=dev)
            out_dev = torch.empty(rows, dtype=torch.float32, device=dev)
            starts = list(range(0, rows, step))
            copied = [torch.cuda.Event() for _ in range(2)]
            consumed = [torch.cuda.Event() for _ in range(2)]
            self._copy_stream.wait_stream(main)

            def issue_copy(i):
                r0 = starts[i]
                n = min(step, rows - r0)
                with torch.cuda.stream(self._copy_stream):
                    if i >= 2:
                        self._copy_stream.wait_event(consumed[i % 2])
                    self._img_stage[i % 2, :n].copy_(x2[r0:r0 + n], non_blocking=True)
                    copied[i % 2].record(self._copy_stream)

            if starts:
                issue_copy(0)
            for i, r0 in enumerate(starts):
                n = min(step, rows - r0)
                if i + 1 < len(starts):
                    issue_copy(i + 1)
                main.wait_event(copied[i % 2])
                out_dev[r0:r0 + n] = self._log_prob(self._img_stage[i % 2, :n].reshape(n, *self._event_shape()))
                consumed[i % 2].record(main)
            out_host.reshape(-1)[:rows].copy_(out_dev, non_blocking=True)
            main.synchronize()
            return out_host
        # Chunk schedule.  The H2D copy of chunk i+1 hides under the kernels of chunk i (a row copies faster than it
        # computes), so only the FIRST copy is exposed: start with the smallest wave-aligned chunk and let the later
        # ones grow (fewer kernel boundaries, fuller tails).  An explicit `chunk_rows` gives uniform chunks.
        unit = min(chunk_rows or _wave_aligned_chunk_rows(prog, dev), max(rows, 1))
        sizes, left = [], rows
        while left > 0:
            mult = 1 if chunk_rows else HOST_CHUNK_UNITS[min(len(sizes), len(HOST_CHUNK_UNITS) - 1)]
            take = min(unit * mult, left)
            if left - take < unit // 2:                  # do not leave a sliver for a last launch sequence
                take = left
            sizes.append(take)
            left -= take
        chunk = max(sizes) if sizes else unit
        starts = [sum(sizes[:i]) for i in range(len(sizes))]
        # staging: one device buffer per chunk (slices of one allocation holding the whole batch) when that fits -- the
        # copies then run back to back on the copy stream, independent of the kernels --, else two buffers used in turn
        whole = rows * d * 4 <= HOST_STAGE_BYTES and len(sizes) <= HOST_MAX_STAGE_SLICES
        need = max(rows, 1) if whole else 2 * chunk
        if getattr(self, "_host_stage", None) is None or self._host_stage.shape[0] < need \
                or self._host_stage.shape[1] != d or self._host_stage.device != dev:
            self._host_stage = torch.empty(need, d, dtype=torch.float32, device=dev)
            self._host_out = torch.empty(0, dtype=torch.float32, device=dev)
            self._copy_stream = torch.cuda.Stream(device=dev)
        n_buf = max(1, len(sizes)) if whole else 2
        if whole:
            bufs = [self._host_stage[starts[i]:starts[i] + sizes[i]] for i in range(len(sizes))] or [self._host_stage[:0]]
        else:
            bufs = [self._host_stage[:chunk], self._host_stage[chunk:2 * chunk]]
        if self._host_out.numel() < rows:
            self._host_out = torch.empty(rows, dtype=torch.float32, device=dev)
        out_dev = self._host_out[:rows]
        main = torch.cuda.current_stream(dev)
        copied = [torch.cuda.Event() for _ in range(n_buf)]
        consumed = [torch.cuda.Event() for _ in range(n_buf)]
        guarded = prog.mode == "fp32" and not prog.force_fallback       # fp16-split engine: range flag per chunk
        flags = torch.zeros(len(starts), dtype=torch.int32, device=dev) if guarded else None

        def make_sink(r0):
            def sink(z_chunk, a, b):
                base._density_into(ops.Act(b - a, d, f32=z_chunk), -ladj, out_dev[r0 + a:r0 + b])
            return sink

        use_graphs = HOST_CUDA_GRAPHS and not prog.force_fallback and rows >= unit and getattr(prog, "small", None) is None \
            and not prog.has_row_ladj
        row_ladj = torch.zeros(rows, dtype=torch.float32, device=dev) if prog.has_row_ladj else None
        with torch.no_grad():
            graphs = None
            if use_graphs:                                # one graph per (staging buffer, chunk size), largest first so
                graphs = {}                               # that the shared workspaces are sized once
                for _ in range(2):                        # a workspace that grew while capturing invalidates earlier graphs
                    for i in sorted(range(len(sizes)), key=lambda j: -sizes[j]):
                        if sizes[i] >= unit // 2:
                            graphs[(i % n_buf, sizes[i])] = self._chunk_graph(prog, i % n_buf, bufs[i % n_buf][:sizes[i]], d)
                    if all(g["gen"] == engine._workspace.generation for g in graphs.values()):
                        break
                for g in graphs.values():
                    if g["flag"] is not None:
                        g["flag"].zero_()
            self._copy_stream.wait_stream(main)
            for i, r0 in enumerate(starts):
                r1 = r0 + sizes[i]
                slot = i % n_buf
                buf = bufs[slot][: r1 - r0]
                with torch.cuda.stream(self._copy_stream):
                    if i >= n_buf:
                        self._copy_stream.wait_event(consumed[slot])
                    buf.copy_(x2[r0:r1], non_blocking=True)
                    copied[slot].record(self._copy_stream)
                main.wait_event(copied[slot])
                g = None if graphs is None else graphs.get((slot, sizes[i]))
                if g is not None:
                    # one graph launch replaces the ~23 kernel launches of the chunk: the host-side launch cost
                    # (~40 us per launch through ctypes + tensor-map encoding) is what bounds small chunks otherwise
                    g["graph"].replay()
                    make_sink(r0)(g["fin"], 0, sizes[i])
                else:
                    prog.run(buf, chunk_rows=chunk, sink=make_sink(r0),
                             flag_out=flags[i:i + 1] if guarded else None,
                             ladj_rows=None if row_ladj is None else row_ladj[r0:r1])
                consumed[slot].record(main)
            if guarded:                                   # one sync; out-of-range chunks go through the tf32 split
                redo = set(torch.nonzero(flags).reshape(-1).tolist())
                if graphs and any(int(g["flag"].item()) != 0 for g in graphs.values() if g["flag"] is not None):
                    redo = set(range(len(starts)))        # a graph's flag is not per chunk: recompute all of them
                for i in sorted(redo):
                    r0 = starts[i]
                    r1 = r0 + sizes[i]
                    buf = bufs[i % n_buf][: r1 - r0]
                    buf.copy_(x2[r0:r1])
                    if row_ladj is not None:
                        row_ladj[r0:r1].zero_()
                    prog._fallback().run(buf, chunk_rows=chunk, sink=make_sink(r0),
                                         ladj_rows=None if row_ladj is None else row_ladj[r0:r1])
            if row_ladj is not None and rows:
                ops.sub_rows(out_dev, row_ladj)
            out_host.reshape(-1)[:rows].copy_(out_dev, non_blocking=True)
            main.synchronize()
        return out_host

    def sample(self, sample_shape: Iterable[int] = None, context: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Draw from the base and push through the layers (flows.py:247-265)."""
        if sample_shape is None:
            sample_shape = [1]
        shape = [int(s) for s in sample_shape]
        with ops.on_device(next(self.parameters())):
            z = self.base_distribution.sample(shape)
            ev = len(self._event_shape())
            z = z.reshape(-1, *z.shape[z.dim() - ev:])
            y = self._run("forward", z, self._zero_ctx) if context is None else self._with_context("forward", z, context)
        return y.reshape(*shape, *y.shape[1:])

    def fit(self, data_train, optim=None, optim_params: Optional[Dict[str, Any]] = None, batch_size: int = 32,
            shuffle: bool = True, gradient_clip: Optional[float] = None, device=None, epochs: int = 1, **kw):
        """Maximum-likelihood training loop with the reference's signature (flows.py:113-210); returns the epoch
        losses.  Default optimiser: `usflows_b200.optim.SophiaG`.  Under an initialised `torch.distributed`
        process group the batch is sharded over the ranks and the gradients are all-reduced (see training.py)."""
        from . import training
        return training.fit(self, data_train, optim=optim, optim_params=optim_params, batch_size=batch_size,
                            shuffle=shuffle, gradient_clip=gradient_clip, device=device, epochs=epochs, **kw)

    def to(self, device):
        self.device = device
        self.trainable_layers = torch.nn.ModuleList([l.to(device) for l in self.trainable_layers])
        for l in self.layers:                                    # plain-attribute tensors (masks)
            if isinstance(l, BaseTransform):
                l.to(device)
        self._distribution_to(device)
        self._programs = {}
        return super().to(device)

    def is_feasible(self) -> bool:
        return all(l.is_feasible() for l in self.layers if isinstance(l, BaseTransform))

    def add_jitter(self, jitter: float = 1e-6) -> None:
        for l in self.layers:
            if isinstance(l, BaseTransform) and not l.is_feasible():
                l.add_jitter(jitter)

    def log_prior(self):
        return 0

    def _distribution_to(self, device) -> None:
        pass

    # -- engine glue -----------------------------------------------------------------------------
    def _base_module(self) -> DistributionModule:
        b = self.base_distribution
        return b.base_dist if isinstance(b, Independent) else b

    def _event_shape(self):
        return self.base_distribution.event_shape

    def _channels_last_perm(self, device) -> torch.Tensor:
        C, H, W = self._event_shape()
        key = (C, H, W, str(device))
        if getattr(self, "_cl_perm_key", None) != key:
            self._cl_perm, self._cl_perm_key = image_engine.channels_last_index(C, H * W, device), key
        return self._cl_perm

    def _weights_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _program(self, direction: str, zero_ctx: bool = False):
        """(Program, total forward log|det J|) for the current weight version; `zero_ctx`: lowered for the zero context of a
        soft-training `log_prob` / `sample` (only ConditionalDenseNN conditioners tell it from no context)."""
        mode = self.precision or engine.get_precision()
        key = (mode,) + self._weights_key()
        slot = direction
        if self._cond_dense_nets:
            slot = (direction, bool(zero_ctx))
        hit = self._programs.get(slot)
        if hit is None or hit[0] != key:
            for net in self._cond_dense_nets:
                net.zero_context_default = bool(zero_ctx)
            with torch.no_grad():
                ev = tuple(self._event_shape())
                if len(ev) == 3:
                    prog = image_engine.ImageProgram(self.layers, direction, mode, ev)
                elif len(ev) == 1:
                    prog = engine.Program(self.layers, direction, mode)
                else:
                    raise NotImplementedError("usflows_b200: events of shape [d] and [C, H, W] are built")
            ladj, n_bad = engine.total_ladj(self.layers)
            hit = (key, prog, ladj, n_bad)
            self._programs[slot] = hit
        return hit[1], hit[2]

    def _apply_plan(self, prog, x2: torch.Tensor) -> torch.Tensor:
        rows, d = x2.shape
        x2 = x2.contiguous()
        chunk, capacity = self._plan_chunks(rows)
        plan = prog.c_plan(d, capacity)
        width = prog.out_width(d)
        y = torch.empty(rows, width, dtype=torch.float32, device=x2.device)
        starts = list(range(0, rows, chunk))
        flags = torch.zeros(len(starts), dtype=torch.int32, device=x2.device) if prog.mode == "fp32" else None
        for ci, r0 in enumerate(starts):
            r1 = min(rows, r0 + chunk)
            plan.apply(x2[r0:r1], y[r0:r1], None if flags is None else flags[ci:ci + 1])
        if flags is not None:
            for ci in torch.nonzero(flags).reshape(-1).tolist():
                r0 = starts[ci]
                r1 = min(rows, r0 + chunk)
                prog._fallback().run(x2[r0:r1], out=y[r0:r1])
        return y

    def _run(self, direction: str, x: torch.Tensor, zero_ctx: bool = False) -> torch.Tensor:
        if self._layer_route:
            return self._with_context(direction, x, None)
        x2, batch_shape = engine._flatten_rows(x, len(self._event_shape()))
        with torch.no_grad(), ops.on_device(x2):
            prog, _ = self._program(direction, zero_ctx)
            if USE_C_PLAN and x2.is_cuda and x2.shape[0] > 0 and isinstance(prog, engine.Program) and prog.plan_able() \
                    and not torch.cuda.is_current_stream_capturing():
                y = self._apply_plan(prog, x2)
            else:
                y = prog.run(x2)
        return y.reshape(*batch_shape, *self._event_shape())


class USFlow(Flow):
    """Uniformly scaling flow: [BlockAffine(LU..[, Householder]) -> additive MaskedCoupling
    [-> BlockAffine^-1]] x coupling_blocks -> BlockAffine(LU) -> Scale  (flows.py:380-491)."""

    MASKTYPE = Literal["checkerboard", "channel"]

    _context_needs_soft_training = True

    def __init__(self, base_distribution, in_dims: List[int], coupling_blocks: int,
                 conditioner_cls: Type[torch.nn.Module], conditioner_args: Dict[str, Any], soft_training=False,
                 prior_scale: Optional[float] = None, training_noise_prior=None, affine_conjugation: bool = False,
                 nonlinearity: Optional[torch.nn.Module] = None, lu_transform: int = 1, householder: int = 1,
                 masktype: str = "checkerboard", *args, **kwargs):
        self.coupling_blocks = coupling_blocks
        self.in_dims = in_dims
        self.conditioner_cls = conditioner_cls
        self.conditioner_args = conditioner_args
        self.prior_scale = prior_scale
        if masktype == "checkerboard":
            self.mask_Generator = USFlow.create_checkerboard_mask
        elif masktype == "channel":
            self.mask_Generator = USFlow.create_channel_mask
        else:
            raise ValueError(f"Unknown mask type {masktype}")
        if lu_transform < 0:
            raise ValueError("Number of LU transforms must be non-negative")
        if householder < 0:
            raise ValueError("Number of Householder vectors transforms must be non-negative")
        self.lu_transform = lu_transform
        self.householder = householder
        # EXTENSION (not in the reference, whose coupling is additive only): coupling="affine" builds scale-and-shift
        # couplings; the conditioner then has to emit 2*d values (DenseNN: param_dims=[d, d])
        coupling = kwargs.pop("coupling", "additive")
        if coupling not in ("additive", "affine"):
            raise ValueError(f"Unknown coupling type {coupling}")
        self.coupling = coupling

        layers = []
        mask = self.mask_Generator(in_dims)
        for _ in range(coupling_blocks):
            affine_layers = [LUTransform(in_dims[0], prior_scale) for _ in range(lu_transform)]
            if householder > 0:
                affine_layers.append(HouseholderTransform(dim=in_dims[0], nvs=householder))
            block_affine_layer = None
            if affine_layers:
                block_affine_layer = BlockAffineTransform(in_dims, SequentialAffineTransform(affine_layers))
                layers.append(block_affine_layer)
            coupling_cls = MaskedAffineCoupling if coupling == "affine" else MaskedCoupling
            layers.append(coupling_cls(mask, conditioner_cls(**conditioner_args)))
            if affine_conjugation and block_affine_layer is not None:
                layers.append(InverseTransform(block_affine_layer))   # shares parameters (flows.py:469-470)
            mask = 1 - mask
        layers.append(BlockAffineTransform(in_dims, LUTransform(in_dims[0], prior_scale)))
        layers.append(ScaleTransform(in_dims))
        super().__init__(base_distribution, layers, soft_training=soft_training,
                         training_noise_prior=training_noise_prior, *args, **kwargs)

    @classmethod
    def create_checkerboard_mask(cls, in_dims, invert: bool = False) -> torch.Tensor:
        """(sum of index coordinates) mod 2, float32, shape (1, *in_dims)  (flows.py:494-514)."""
        axes = [torch.arange(d, dtype=torch.int32) for d in in_dims]
        idx = torch.stack(torch.meshgrid(*axes, indexing="ij"))
        mask = torch.fmod(idx.sum(dim=0), 2).to(torch.float32).view(1, *in_dims)
        return 1 - mask if invert else mask

    @classmethod
    def create_channel_mask(cls, in_dims, invert: bool = False) -> torch.Tensor:
        """(first index coordinate) mod 2  (flows.py:516-536)."""
        axes = [torch.arange(d, dtype=torch.int32) for d in in_dims]
        idx = torch.stack(torch.meshgrid(*axes, indexing="ij"))
        mask = torch.fmod(idx[0], 2).to(torch.float32).view(1, *in_dims)
        return 1 - mask if invert else mask

    def log_prior(self):
        """Sum of the top-level layers' log_prior -- identically 0 in the reference because
        BlockAffineTransform does not forward to its LU layer (flows.py:538-549)."""
        if self.prior_scale is not None:
            return sum(l.log_prior() for l in self.layers)
        return 0

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        """Accepts reference checkpoints: InverseTransform layers alias block i's parameters and appear a
        second time under `trainable_layers.{i+2}.transform...` -- both copies are accepted."""
        return super().load_state_dict(state_dict, strict=strict, **kw)
