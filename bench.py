#!/usr/bin/env python
"""Benchmark of the USFlows hot path on B200:  `log_prob` throughput (samples/s) of a flat USFlow.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c4|c5|c1] [--precision fp32|tf32|bf16]
    python bench.py --impl reference ...      # the reference algorithm's CPU path (oracle port) on host cores

Contract (one JSON line on stdout, printed by rank 0): metric/value/unit/n_gpus/steps/warmup/ms_per_step/
higher_is_better/scaling/vs_baseline/dtype/data/config + roofline, cpu_baseline, e2e, clocks, gpu_launches.

A "step" is one `log_prob` pass over one batch of synthetic rows per GPU (weak scaling: every rank evaluates
its own `rows` rows, no data-path collective).  `value` is measured with the batch resident in HBM; `e2e` goes
through `Flow.log_prob_host` with pinned HOST buffers (H2D of the batch and D2H of the log-probs inside the
timed region).  Workloads follow BASELINE.json / SURVEY 8d:
  c2: d=784,  B=4, MLP [1024,1024], Laplace base, 65536 rows   <- the configuration the metric is quoted on
  c4: d=3072, B=8, MLP [1024,1024], Normal base,  32768 rows
  c5: d=3072, B=4, MLP [1024,1024], Laplace base, 32768 rows
  c1: d=2,    B=10, MLP [32,32],    Laplace base, 1048576 rows
Inputs are larger than L2 (c2: 205 MB per batch vs 126 MB), so no explicit L2 flush is needed between steps.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c2": dict(spec=dict(in_dims=[784], coupling_blocks=4, hidden_dims=[1024, 1024], affine_conjugation=True,
                         lu_transform=1, householder=0, base="laplace"), rows=65536, cpu_rows=8192,
               name="C2 MNIST-shaped 784-D USFlow (B=4, MLP 1024x1024, Laplace) batch log_prob"),
    "c4": dict(spec=dict(in_dims=[3072], coupling_blocks=8, hidden_dims=[1024, 1024], affine_conjugation=True,
                         lu_transform=1, householder=0, base="normal"), rows=32768, cpu_rows=1024,
               name="C4 CIFAR-shaped 3072-D USFlow (B=8, MLP 1024x1024, Normal) batch log_prob"),
    "c5": dict(spec=dict(in_dims=[3072], coupling_blocks=4, hidden_dims=[1024, 1024], affine_conjugation=True,
                         lu_transform=1, householder=0, base="laplace"), rows=32768, cpu_rows=2048,
               name="C5 3072-D USFlow (B=4, MLP 1024x1024, Laplace) batch log_prob"),
    # C2's shape with the reference's own MLP-style conditioner (networks.ConvNet, vector branch: 2 gated blocks of width
    # 1024 with LayerNorm) and the base of its live MNIST configuration (experiments/mnist/mnist.yaml:79-92: L1-radial,
    # LogNormal radius) -- SURVEY 8f rows 2 and 4
    "c2cn": dict(spec=dict(in_dims=[784], coupling_blocks=4, conditioner="convnet", c_hidden=[1024, 1024], gating=True,
                           normalize_layers=True, affine_conjugation=True, lu_transform=1, householder=0, base="radial",
                           p=1, norm="lognormal"), rows=65536, cpu_rows=4096,
                 name="C2-shaped 784-D USFlow (B=4, ConvNet conditioner 2 gated blocks x 1024 + LayerNorm, L1-radial "
                      "LogNormal base) batch log_prob"),
    # the reference's live MNIST configuration (experiments/mnist/mnist.yaml:44-92): 28x28 images squeezed to [16, 7, 7],
    # 15 coupling blocks, ConvNet2D conditioner (32 hidden channels, 3 gated 3x3 blocks, LayerNormChannels), conjugated
    # 1x1-convolution LU layers, L1-radial LogNormal base -- SURVEY 8f row 3
    "mnist_img": dict(spec=dict(in_dims=[16, 7, 7], coupling_blocks=15, conditioner="convnet2d", c_hidden=32, num_layers=3,
                                kernel_size=3, gating=True, normalize_layers=True, affine_conjugation=True, lu_transform=1,
                                householder=0, base="radial", p=1, norm="lognormal"), rows=16384, cpu_rows=256,
                      name="MNIST [16,7,7] USFlow (B=15, ConvNet2D 32 ch x 3 gated 3x3 blocks + LayerNorm, 1x1-conv LU, "
                           "L1-radial LogNormal base) batch log_prob"),
    "c1": dict(spec=dict(in_dims=[2], coupling_blocks=10, hidden_dims=[32, 32], affine_conjugation=True,
                         lu_transform=1, householder=0, base="laplace"), rows=1 << 20, cpu_rows=65536,
               name="C1 2-D USFlow (B=10, MLP 32x32, Laplace) batch log_prob"),
}


def algorithmic_flops_per_sample(spec) -> float:
    """(2B+1) * 2 d^2 + B * 2 (d H + H^2 + H d) with conjugation (B+1 affine layers without)."""
    d, B = spec["in_dims"][0], spec["coupling_blocks"]
    if spec.get("conditioner") == "convnet2d":     # per pixel: 1x1-conv affine layers 2 C^2; k x k convolutions 2 k^2 Cin Cout
        C, hw = spec["in_dims"][0], spec["in_dims"][1] * spec["in_dims"][2]
        ch, kk, L = spec["c_hidden"], spec.get("kernel_size", 3) ** 2, spec["num_layers"]
        per_block = 2 * kk * ch * ch + (2 * ch * 2 * ch if spec.get("gating", True) else 0)
        cond = 2 * kk * C * ch + L * per_block + 2 * kk * ch * C
        n_aff = 2 * B + 1 if spec.get("affine_conjugation") else B + 1
        return hw * (n_aff * 2 * C * C + B * cond)
    if spec.get("conditioner") == "convnet":       # Linear(d,h0) + per block [h_in h + h 2h (+ proj)] + Linear(h_last, d)
        ch = list(spec["c_hidden"])
        mlp = 2 * d * ch[0] + 2 * ch[-1] * d
        for i, oc in enumerate(ch):
            ic = ch[i - 1] if i > 0 else ch[0]
            mlp += 2 * ic * oc + (2 * oc * 2 * oc if spec.get("gating", True) else 0) + (2 * ic * oc if ic != oc else 0)
    else:
        dims = [d] + list(spec["hidden_dims"]) + [d]
        mlp = sum(2 * dims[i] * dims[i + 1] for i in range(len(dims) - 1))
    n_aff = 2 * B + 1 if spec.get("affine_conjugation") else B + 1
    return n_aff * 2 * d * d + B * mlp


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    # MEASURED_PEAKS.json is driver-written; when it is absent use the fallback /opt/skills/guides/B200_PROFILING.md
    # states (6.65 TB/s, 1.59 PFLOP/s burst, ~1.4 PFLOP/s sustained under the power cap) and say so
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons, power = [], None, set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2]); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "power_w_max": max(power) if power else None, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def cpu_reference_run(wl, steps: int, warmup: int, rows: int):
    """The reference algorithm's own CPU path (oracle port = same ATen ops in the same order, including the
    per-call weight re-preparation) on all host cores.  Returns (as-is samples/s, amortised samples/s, cores)."""
    import torch
    from oracle import flow_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = wl["spec"]
    params = O.random_params(spec, 0)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(rows, *spec["in_dims"], generator=g)
    with torch.no_grad():
        for _ in range(warmup):
            O.flow_log_prob(x, spec, params)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.flow_log_prob(x, spec, params)
        t_asis = (time.perf_counter() - t0) / steps
        _, prepared = O.flow_log_prob_amortised(x, spec, params)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.flow_log_prob_amortised(x, spec, params, prepared=prepared)
        t_am = (time.perf_counter() - t0) / steps
    return rows / t_asis, rows / t_am, cores, t_asis


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="usflows_b200", choices=["usflows_b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp32_tf32", "fp32_simt", "tf32", "bf16"])
    ap.add_argument("--rows", type=int, default=0, help="rows per GPU per step (default: the workload's)")
    ap.add_argument("--chunk-rows", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-modes", action="store_true", help="skip the extra tf32 / bf16 mode measurements")
    ap.add_argument("--only-logprob", action="store_true",
                    help="profiling aid: warm-up + timed log_prob steps only, then exit (no JSON line)")
    ap.add_argument("--train", action="store_true",
                    help="also time the MLE training step (C3: global batch 8192 x n_gpus, gradient all-reduce)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = WORKLOADS[args.workload]
    spec = wl["spec"]
    d = 1
    for v in spec["in_dims"]:
        d *= v
    flops_per_sample = algorithmic_flops_per_sample(spec)

    base_line = dict(metric="log_prob_samples_per_sec", unit="samples/s", n_gpus=args.gpus, steps=args.steps,
                     warmup=args.warmup, higher_is_better=True, scaling="weak", vs_baseline=None,
                     data="synthetic")

    # ---------------------------------------------------------------- reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        rows = args.rows or wl["cpu_rows"]
        v_asis, v_am, cores, t = cpu_reference_run(wl, max(1, args.steps), max(1, min(args.warmup, 1)), rows)
        line = dict(base_line, impl="reference", value=v_asis, ms_per_step=t * 1e3, dtype="f32",
                    config=dict(workload=wl["name"], rows_per_step=rows, d=d, hidden=spec.get("hidden_dims", spec.get("c_hidden")),
                                coupling_blocks=spec["coupling_blocks"], engine="torch CPU (oracle port of the reference)"),
                    cpu_baseline=dict(value=v_asis, unit="samples/s", cores=cores, kind="port",
                                      sample=f"{rows} rows x {max(1, args.steps)} steps of the same workload; "
                                             f"as-is (per-call weight re-preparation, as the reference does); "
                                             f"amortised = {v_am:.1f} samples/s", amortised_value=v_am),
                    e2e=dict(value=v_asis, unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line), flush=True)
        return

    # ---------------------------------------------------------------- B200 arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import usflows_b200 as U
    from usflows_b200 import engine, ops
    from oracle import flow_oracle as O      # parameters only (deterministic synthetic model); timed code is ours
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import build_flow

    if args.chunk_rows:
        U.set_chunk_rows(args.chunk_rows)
        if len(spec["in_dims"]) > 1:                 # image-shaped events: channels-last rows (N*H*W) per chunk
            from usflows_b200 import image_engine
            image_engine.IMAGE_CHUNK_ROWS = args.chunk_rows
    rows = args.rows or wl["rows"]
    params = O.random_params(spec, 0)
    flow = build_flow(spec, params, device=dev, precision=args.precision)
    g = torch.Generator().manual_seed(1 + rank)
    x_host = torch.rand(rows, *spec["in_dims"], generator=g).pin_memory()
    x = x_host.to(dev)
    out_host = torch.empty(rows, dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    # preparation (once per weight version) timed separately
    t0 = time.perf_counter()
    lp = flow.log_prob(x[:256])
    torch.cuda.synchronize()
    prep_ms = (time.perf_counter() - t0) * 1e3

    step = lambda: flow.log_prob(x)                    # noqa: E731
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ops.LAUNCHES = 0
    ms_step = timed(step, args.steps)
    launches = ops.LAUNCHES
    if args.only_logprob:
        if rank == 0:
            sampler.stop()
            print(f"log_prob: {ms_step:.3f} ms/step, {launches} launches in {args.steps} steps", flush=True)
        return
    value = world * rows / (ms_step * 1e-3)

    # the sampling pass (latent -> data) of the same model and batch size: base draws + forward layer stack
    sample_step = lambda: flow.sample([rows])          # noqa: E731
    for _ in range(max(1, args.warmup // 2)):
        sample_step()
    ms_sample = timed(sample_step, max(2, args.steps // 2))

    # end to end: pinned host rows in, host log-probs out, copies inside the timed region
    e2e_step = lambda: flow.log_prob_host(x_host, out_host)   # noqa: E731
    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # kernel-class breakdown of one step (events around every launch; after the timed region)
    breakdown = engine.profile_step(lambda: flow.log_prob(x))
    gemm_ms = sum(v for k, v in breakdown.items() if k.startswith("linear"))
    if len(spec["in_dims"]) > 1:            # image path: the convolutions (implicit GEMM, or gather + contraction)
        gemm_ms += breakdown.get("im2col", 0.0) + breakdown.get("conv2d_rows", 0.0)
    peaks, peak_kind = measured_peaks()
    # DRAM traffic of the dominant kernel per launch, from the committed ncu --set full capture (profiles/): the mean of
    # dram__bytes_read.sum + dram__bytes_write.sum over the captured launches of one step
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath) and args.workload == "c2" and args.precision == "fp32":
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    achieved_tf = rows * flops_per_sample / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    peak_tf = peaks["bf16_tflops_sustained"]
    n_gemm = sum(1 for k in breakdown.get("_names", []) if k.startswith("linear")) or 1

    extra_modes = {}
    if not args.no_modes and rank == 0 and world == 1:
        for mode in ("fp32_tf32", "tf32", "bf16"):
            if mode == args.precision:
                continue
            f2 = build_flow(spec, params, device=dev, precision=mode)
            f2.log_prob(x[:256])
            for _ in range(2):
                f2.log_prob(x)
            ms = timed(lambda: f2.log_prob(x), max(3, args.steps // 2))
            ref = lp.double().cpu()
            err = float(((f2.log_prob(x[:256]).double().cpu() - ref).abs() / ref.abs().clamp(min=1)).max())
            extra_modes[mode] = dict(value=rows / (ms * 1e-3), ms_per_step=ms,
                                     tflops=rows * flops_per_sample / (ms * 1e-3) / 1e12,
                                     max_rel_diff_vs_fp32_mode=err)
            del f2

    train = None
    if args.train:
        # C3: Fashion-MNIST-shaped MLE training, weak scaling: 8192 rows per GPU per step, one NCCL all-reduce of the
        # flat gradient buffer per step (usflows_b200/training.py), SophiaG lr=1e-3 wd=0 as the reference config
        import numpy as np
        from usflows_b200 import training
        per_gpu = 8192
        tflow = build_flow(spec, O.random_params(spec, 0), device=dev, precision=args.precision)
        tparams = list(tflow.parameters())
        opt = U.SophiaG(tparams, lr=1e-3, weight_decay=0.0)
        gt = torch.Generator().manual_seed(100 + rank)
        xt = torch.rand(per_gpu, d, generator=gt).to(dev)

        def train_step():
            opt.zero_grad()
            loss = -training.log_prob_autograd(tflow, xt).sum() / (per_gpu * world)
            loss.backward()
            if world > 1:
                training.allreduce_gradients(tparams)
            opt.step()
            return loss
        for _ in range(3):
            train_step()
        ms_train = timed(train_step, max(3, args.steps))
        train = dict(metric="train_samples_per_sec", value=world * per_gpu / (ms_train * 1e-3), unit="samples/s",
                     ms_per_step=ms_train, global_batch=per_gpu * world, grad_floats=sum(p.numel() for p in tparams),
                     note="forward + backward (all batch-side GEMMs on the tcgen05 tf32-split engine) + gradient "
                          "all-reduce + SophiaG step; feasibility check not in the timed region")
        del tflow, opt

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        v_asis, v_am, cores, _ = cpu_reference_run(wl, 3, 1, wl["cpu_rows"])
        cpu = dict(value=v_asis, unit="samples/s", cores=cores, kind="port",
                   sample=f"{wl['cpu_rows']} rows x 3 steps of the same workload through the oracle port of the "
                          f"reference (torch CPU, per-call weight re-preparation as the reference does)",
                   amortised_value=v_am)

    roofline = dict(bound="tensor", achieved=achieved_tf, peak=peak_tf, unit="TFLOP/s",
                    frac=achieved_tf / peak_tf, traffic=traffic,
                    peak_source=f"bf16_tflops_sustained, of {peak_kind}",
                    kernel="tc2::gemm_tc2_kernel (CTA-pair tcgen05, all launches of one step)" if len(spec["in_dims"]) == 1
                    else "convtc::conv_tc_kernel (implicit-GEMM tcgen05 convolution) + tc2::gemm_tc2_kernel / SIMT 1x1 "
                         "convolutions, all launches of one step",
                    kernel_ms_per_step=gemm_ms,
                    launches_per_step=n_gemm,
                    note="algorithmic fp32 FLOPs over the summed CUDA-event time of the GEMM launches of one step; "
                         "the fp32 mode spends 3 fp16 MMAs per algorithmic MAC (fp16 runs at the bf16 rate), so its "
                         "ceiling is 1/3 of this peak (1/6 for fp32_tf32)",
                    frac_of_mode_ceiling=(achieved_tf / (peak_tf / {"fp32": 3.0, "fp32_tf32": 6.0}[args.precision]))
                    if args.precision in ("fp32", "fp32_tf32") else None)
    if breakdown.get("flow_small"):
        # tiny event size: the whole stack is one FP32-FMA bound launch (usf_flow_small); neither HBM nor the tensor pipe
        # bounds it, so the figure is set against the nominal FP32 FMA rate of the part (stated, not measured)
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        fma_peak = sm_count * 128 * 2 * 1.965e9 / 1e12
        ach = rows * flops_per_sample / (breakdown["flow_small"] * 1e-3) / 1e12
        roofline = dict(bound="fp32_fma", achieved=ach, peak=fma_peak, unit="TFLOP/s", frac=ach / fma_peak, traffic=None,
                        peak_source="nominal: SMs x 128 FMA lanes x 2 x 1965 MHz", kernel="flow_small_kernel",
                        kernel_ms_per_step=breakdown["flow_small"], launches_per_step=1,
                        hbm_gbs_for_context=rows * (4 * d + 4) / (breakdown["flow_small"] * 1e-3) / 1e9)

    line = dict(
        base_line, impl="usflows_b200", value=value, ms_per_step=ms_step,
        dtype={"fp32": "f32 (fp16-split x3 on tcgen05 kind::f16, fp32 accumulate/promote; tf32-split fallback)",
               "fp32_tf32": "f32 (tf32-split x3 on tcgen05 kind::tf32)", "fp32_simt": "f32", "tf32": "tf32",
               "bf16": "bf16"}[args.precision],
        config=dict(workload=wl["name"], rows_per_gpu_per_step=rows, d=d, hidden=spec.get("hidden_dims", spec.get("c_hidden")),
                    coupling_blocks=spec["coupling_blocks"], precision=args.precision,
                    chunk_rows=engine._default_chunk_rows, l2="inputs larger than L2, no flush",
                    flops_per_sample=flops_per_sample, prep_ms_once_per_weight_version=prep_ms),
        roofline=roofline,
        cpu_baseline=cpu,
        e2e=dict(value=world * rows / (ms_e2e * 1e-3), unit="samples/s", ms_per_step=ms_e2e,
                 h2d_bytes_per_step=rows * d * 4, d2h_bytes_per_step=rows * 4),
        sample=dict(metric="sample_samples_per_sec", value=world * rows / (ms_sample * 1e-3), unit="samples/s",
                    ms_per_step=ms_sample, tflops=rows * flops_per_sample / (ms_sample * 1e-3) / 1e12,
                    note="Flow.sample([rows]): Philox base draws + forward pass; same algorithmic FLOPs per sample"),
        train=train, gpu_launches=launches, clocks=clocks, modes=extra_modes,
        breakdown_ms={k: v for k, v in breakdown.items() if not k.startswith("_")})
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
