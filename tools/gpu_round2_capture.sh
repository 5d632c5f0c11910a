#!/bin/bash
# End-of-round-2 capture on one B200: tests, smoke, bench lines, ncu launch lists and full captures (summarised under profiles/).
T=${1:-r2z}
mkdir -p gpurun_out
rm -f gpurun_out/parity_elementwise.jsonl
if [ -z "$SKIP_TESTS" ]; then
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/${T}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1
fi
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
python bench.py --workload c1 --no-modes --no-extra --no-train > gpurun_out/${T}_bench_c1.json 2> gpurun_out/${T}_bench_c1.err
python bench.py --workload c2cn --no-modes --no-extra --no-train > gpurun_out/${T}_bench_c2cn.json 2> gpurun_out/${T}_bench_c2cn.err
python bench.py --workload mnist_img --no-modes --no-extra --no-train > gpurun_out/${T}_bench_mnist_img.json 2> gpurun_out/${T}_bench_mnist_img.err
python bench.py --workload c5 --sweep-full --no-modes --no-train --no-cpu-baseline > gpurun_out/${T}_bench_c5_sweep.json 2> gpurun_out/${T}_bench_c5_sweep.err
python tools/small_batch_probe.py > gpurun_out/${T}_small_batch.log 2>&1
python tools/train_breakdown.py > gpurun_out/${T}_train_breakdown.log 2>&1
USF_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_c2.csv python bench.py --steps 2 --warmup 3 --only-logprob > gpurun_out/${T}_ncu1.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_train.csv python tools/train_one_step.py > gpurun_out/${T}_ncu2.log 2>&1
USF_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc2 -c 21 -o /tmp/${T}_full_gemm python bench.py --steps 1 --warmup 3 --only-logprob > gpurun_out/${T}_ncu3.log 2>&1
USF_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc2 -c 21 -o /tmp/${T}_full_gemm_bf16 python bench.py --precision bf16 --steps 1 --warmup 3 --only-logprob > gpurun_out/${T}_ncu4.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:planes_glue|mat_prep|tri_combine|base_backward' -c 12 -o /tmp/${T}_full_train python tools/train_one_step.py > gpurun_out/${T}_ncu5.log 2>&1
# the reports stay on the box (gpurun copies at most 64 MiB back): their raw pages travel as CSV, one small report for source-level reading
for n in full_gemm full_gemm_bf16 full_train; do ncu -i /tmp/${T}_${n}.ncu-rep --page raw --csv > gpurun_out/${T}_${n}_raw.csv 2>/dev/null; done
USF_PROFILE_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 2 -c 3 -o gpurun_out/${T}_full_gemm_3launches python bench.py --steps 1 --warmup 3 --only-logprob > gpurun_out/${T}_ncu6.log 2>&1
ls -la gpurun_out/${T}_* | awk '{print $5, $9}' | sort -n | tail -5
python tools/show_bench.py gpurun_out/${T}_bench.json 2>/dev/null | head -5
