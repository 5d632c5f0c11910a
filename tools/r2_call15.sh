#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -x -k "one_tma_operation or split_k or tma_store_path_equals or all_engines" > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2p_pytest.log
tail -6 gpurun_out/r2p_pytest.log
python - <<'PY'
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
import torch, json
from usflows_b200 import _lib, ops
from gemm_bench import ENG, make_case
lib = _lib.load()
for shp in ("65536x784x784", "65536x1024x1024", "65536x392x1024", "65536x1024x392"):
    M, N, K = (int(v) for v in shp.split("x"))
    act, wt, wl, bias, out, _, _ = make_case("3xf16", M, N, K, 0, False)
    for rep in range(2):
        for on in (0, 1):
            lib.usf_debug_set_planes3d(on)
            for _ in range(3): ops.linear(ENG["3xf16"], act, wt, wl, N, K, bias=bias, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): ops.linear(ENG["3xf16"], act, wt, wl, N, K, bias=bias, out=out)
            e1.record(); torch.cuda.synchronize()
            print(shp, "planes3d", on, round(e0.elapsed_time(e1) * 50, 1), "us", flush=True)
lib.usf_debug_set_planes3d(1)
PY
for i in 1 2; do python bench.py --only-logprob --steps 30 --warmup 5 2>/dev/null | tail -1; done
