# End-of-round capture on one B200: tests, smoke, bench lines, ncu launch lists and full captures (see profiles/).
T=${1:-r1x}
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/${T}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
python bench.py --workload c2cn --no-modes > gpurun_out/${T}_bench_c2cn.json 2> gpurun_out/${T}_bench_c2cn.err
python bench.py --workload mnist_img --no-modes > gpurun_out/${T}_bench_mnist_img.json 2> gpurun_out/${T}_bench_mnist_img.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches_c2.csv python bench.py --steps 2 --warmup 1 --only-logprob > gpurun_out/${T}_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${T}_launches_c2cn.csv python bench.py --workload c2cn --steps 2 --warmup 1 --only-logprob > gpurun_out/${T}_ncu2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches_mnist_img.csv python bench.py --workload mnist_img --rows 4096 --steps 1 --warmup 1 --only-logprob > gpurun_out/${T}_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 24 -c 5 -o gpurun_out/${T}_full_gemm python bench.py --steps 1 --warmup 1 --only-logprob > gpurun_out/${T}_ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:gate_norm|radial_logprob' -s 26 -c 3 -o gpurun_out/${T}_full_glue python bench.py --workload c2cn --steps 1 --warmup 1 --only-logprob > gpurun_out/${T}_ncu5.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:im2col|gate_norm|masked_add|layout' -s 200 -c 4 -o gpurun_out/${T}_full_img python bench.py --workload mnist_img --rows 16384 --steps 1 --warmup 1 --only-logprob > gpurun_out/${T}_ncu6.log 2>&1
tail -3 gpurun_out/${T}_pytest.log; tail -4 gpurun_out/${T}_smoke.log
