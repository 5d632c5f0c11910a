#!/bin/bash
# Final round-2 capture on one B200 (after the pixel-plane image path): tests, smoke, the bench lines, launch lists of the C2 and
# the image step, one full ncu capture of usf_conv2d_pix (plain + gated).  Summarised under profiles/ by tools/r02_summarise.py.
T=${1:-r2f}
mkdir -p gpurun_out
rm -f gpurun_out/parity_elementwise.jsonl
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/${T}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
python bench.py --workload c1 --no-modes --no-extra --no-train > gpurun_out/${T}_bench_c1.json 2> gpurun_out/${T}_bench_c1.err
python bench.py --workload c2cn --no-modes --no-extra --no-train > gpurun_out/${T}_bench_c2cn.json 2> gpurun_out/${T}_bench_c2cn.err
python bench.py --workload mnist_img --no-extra --no-train > gpurun_out/${T}_bench_mnist_img.json 2> gpurun_out/${T}_bench_mnist_img.err
python tools/conv_probe.py > gpurun_out/${T}_conv_probe.log 2>&1
python tools/narrow_probe.py >> gpurun_out/${T}_conv_probe.log 2>&1
USF_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_c2.csv python bench.py --steps 2 --warmup 3 --only-logprob > gpurun_out/${T}_ncu1.log 2>&1
USF_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_mnist_img.csv python bench.py --workload mnist_img --steps 1 --warmup 3 --only-logprob > gpurun_out/${T}_ncu2.log 2>&1
for mode in plain gated; do
  ncu --set full --clock-control none --import-source on -k regex:conv_pix_kernel -s 2 -c 1 -o /tmp/${T}_pix_$mode -f python tools/pix_ncu.py $mode > gpurun_out/${T}_ncu_pix_$mode.log 2>&1
  ncu -i /tmp/${T}_pix_$mode.ncu-rep --page raw --csv > gpurun_out/${T}_pix_${mode}_raw.csv 2>/dev/null
done
cp /tmp/${T}_pix_gated.ncu-rep gpurun_out/ 2>/dev/null
tail -3 gpurun_out/${T}_pytest.log; tail -1 gpurun_out/${T}_smoke.log
ls -la gpurun_out/${T}_* | awk '{print $5, $9}' | sort -n | tail -4
