"""Per-chain timeline of the CTA-pair tcgen05 kernel (run on the B200 box).

    python tools/gemm_timeline.py [--engine 3xf16] [--shape 16384x1024x1024] [--bn 0] [--flags 0]

Uses usf_debug_gemm_timeline: cluster 0's MMA issuer and epilogue warp 4 stamp clock64() at the hand-over points
of every accumulation chain.  Prints the median cycle counts of each phase and the time of one launch with the
debug flags (1 = no TMEM drain, 2 = no store phase) so the limiter of the pipeline can be read off directly.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

from usflows_b200 import _lib, ops  # noqa: E402
from gemm_bench import ENG, PASSES, make_case  # noqa: E402


def med(v):
    v = sorted(v)
    return v[len(v) // 2] if v else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", default="3xf16")
    ap.add_argument("--shape", default="16384x1024x1024")
    ap.add_argument("--bn", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=-1)
    ap.add_argument("--flags", default="0,1,2,3")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--lead", type=int, default=-1)
    args = ap.parse_args()
    lib = _lib.load()
    lib.usf_debug_set_block_n(args.bn)
    if args.chunk >= 0:
        lib.usf_set_accum_chunk(args.chunk)
    if args.lead >= 0:
        lib.usf_set_accum_lead(args.lead)
    M, N, K = (int(v) for v in args.shape.split("x"))
    eng = args.engine
    act, wt, wl, bias, out, _, _ = make_case(eng, M, N, K, 0, False)
    run = lambda: ops.linear(ENG[eng], act, wt, wl, N, K, bias=bias, out=out)   # noqa: E731
    for flags in (int(f) for f in args.flags.split(",")):
        buf = torch.zeros(512 * 8, dtype=torch.int64, device="cuda")
        lib.usf_debug_gemm_timeline(None, flags)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / args.iters
        lib.usf_debug_gemm_timeline(buf.data_ptr(), flags)
        run()
        torch.cuda.synchronize()
        lib.usf_debug_gemm_timeline(None, 0)
        t = buf.cpu().view(512, 8).tolist()
        rows = [r for r in t if r[0] and r[3]]
        rec = dict(engine=eng, M=M, N=N, K=K, flags=flags, us=round(us, 1),
                   mma_tflops=round(PASSES[eng] * 2 * M * N * K / us / 1e6, 1), chains=len(rows))
        st = [r for r in rows if r[5]]
        if st:
            rec["store"] = med([r[5] - r[4] for r in st])
            rec["store_math"] = med([r[6] >> 32 for r in st])
            rec["store_convert_sts"] = med([r[6] & 0xffffffff for r in st])
            rec["store_copy_out"] = med([r[7] for r in st])
        if len(rows) > 40:
            body = rows[8:-8]
            idx = {id(r): i for i, r in enumerate(rows)}
            rec["period"] = med([rows[idx[id(r)] + 1][2] - r[2] for r in body])            # issue -> next issue
            rec["wait_empty_to_full"] = med([r[1] - r[0] for r in body])                   # operands late?
            rec["issue"] = med([r[2] - r[1] for r in body])
            rec["full_seen_after_issue"] = med([r[3] - r[2] for r in body])               # chain queue + execution + signal
            rec["drain"] = med([r[4] - r[3] for r in body])
            rec["drain_to_next_issuer_wake"] = med([rows[idx[id(r)] + 2][0] - r[4] for r in body if idx[id(r)] + 2 < len(rows)])
            rec["full_to_full"] = med([rows[idx[id(r)] + 1][3] - r[3] for r in body])
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
