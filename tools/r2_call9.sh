#!/bin/bash
# A/B of the C2 log_prob step: round-1 tree (_old/) vs the current tree, alternating on the same box
for i in 1 2 3; do
  (cd _old && python bench.py --only-logprob --steps 30 --warmup 5 2>/dev/null | tail -1 | sed 's/^/old: /')
  python bench.py --only-logprob --steps 30 --warmup 5 2>/dev/null | tail -1 | sed 's/^/new: /'
done
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_training.py -m gpu -q -p no:cacheprovider -k "split_k or tma_store_path or hand_written" 2>&1 | tail -3
