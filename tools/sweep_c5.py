"""C5: batch sweep of the 3072-D flow (B=4, MLP 1024x1024, Laplace) for log_prob and sample, 1 K .. 64 M rows.

    python tools/sweep_c5.py [--max-log2 24] [--precision fp32]         (one process per GPU under torchrun)

Rows are synthetic and generated on the device: batches above 65 536 rows re-use one resident 65 536-row block
(64 M x 3072 x 4 B = 805 GB does not exist anywhere), so the figure is the kernel-path throughput with inputs in
HBM.  Prints one JSON line per batch size: rows/s (all ranks), algorithmic TFLOP/s per GPU and the fraction of the
measured bf16 peak / of the mode's ceiling.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from usflows_b200.builders import build_flow  # noqa: E402
from oracle import flow_oracle as O  # noqa: E402   (parameters only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-log2", type=int, default=10)
    ap.add_argument("--max-log2", type=int, default=24)
    ap.add_argument("--precision", default="fp32")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    spec = bench.WORKLOADS["c5"]["spec"]
    flops = bench.algorithmic_flops_per_sample(spec)
    peaks, kind = bench.measured_peaks()
    flow = build_flow(spec, O.random_params(spec, 0), device=dev, precision=args.precision)
    block = torch.rand(65536, 3072, device=dev, generator=torch.Generator(device=dev).manual_seed(1 + rank))
    flow.log_prob(block[:256])
    ceil_div = {"fp32": 3.0, "fp32_tf32": 6.0}.get(args.precision, 1.0)
    for lg in range(args.min_log2, args.max_log2 + 1, 2):
        n = 1 << lg                                   # rows per GPU (weak scaling: every rank evaluates n rows)
        full, rest = divmod(n, block.shape[0])
        res = {}
        for what in ("log_prob", "sample"):
            def run():
                for _ in range(full):
                    flow.log_prob(block) if what == "log_prob" else flow.sample([block.shape[0]])
                if rest:
                    flow.log_prob(block[:rest]) if what == "log_prob" else flow.sample([rest])
            run()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = max(1, min(20, (1 << 17) // n))
            e0.record()
            for _ in range(reps):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            res[what] = float(ms)
        if rank == 0:
            tf = n * flops / (res["log_prob"] * 1e-3) / 1e12
            print(json.dumps(dict(rows_per_gpu=n, n_gpus=world, precision=args.precision,
                                  log_prob_rows_per_s=world * n / (res["log_prob"] * 1e-3),
                                  sample_rows_per_s=world * n / (res["sample"] * 1e-3),
                                  log_prob_ms=res["log_prob"], sample_ms=res["sample"], alg_tflops_per_gpu=tf,
                                  frac_of_bf16_peak=tf / peaks["bf16_tflops_sustained"],
                                  frac_of_mode_ceiling=tf / (peaks["bf16_tflops_sustained"] / ceil_div),
                                  peak_source=kind)), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
