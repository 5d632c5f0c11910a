python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r1s_pytest.log
python bench.py > gpurun_out/r1s_bench.json 2> gpurun_out/r1s_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1s_bench_ref.json 2> gpurun_out/r1s_bench_ref.err
python bench.py --workload c2cn --no-modes > gpurun_out/r1s_bench_c2cn.json 2> gpurun_out/r1s_bench_c2cn.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1s_launches.csv python bench.py --steps 2 --warmup 1 --only-logprob > gpurun_out/r1s_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r1s_launches_c2cn.csv python bench.py --workload c2cn --steps 2 --warmup 1 --only-logprob > gpurun_out/r1s_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:gate_norm|radial_logprob' -s 12 -c 4 -o gpurun_out/r1s_glue python bench.py --workload c2cn --steps 1 --warmup 1 --only-logprob > gpurun_out/r1s_ncu3.log 2>&1
tail -3 gpurun_out/r1s_pytest.log; cat gpurun_out/r1s_bench_c2cn.json | cut -c1-1500; tail -3 gpurun_out/r1s_bench_c2cn.err
