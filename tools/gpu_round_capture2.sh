# Final capture of the round (after the implicit-GEMM convolution): tests, smoke, bench lines, launch list of the image path.
T=${1:-r1f}
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/${T}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
python bench.py --workload mnist_img --no-modes > gpurun_out/${T}_bench_mnist_img.json 2> gpurun_out/${T}_bench_mnist_img.err
python tools/conv_probe.py > gpurun_out/${T}_conv_probe.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches_mnist_img.csv python bench.py --workload mnist_img --rows 4096 --steps 1 --warmup 1 --only-logprob > gpurun_out/${T}_ncu3.log 2>&1
tail -3 gpurun_out/${T}_pytest.log; tail -2 gpurun_out/${T}_smoke.log; python tools/show_bench.py gpurun_out/${T}_bench_mnist_img.json | sed -n "1,3p;5p"; python tools/show_bench.py gpurun_out/${T}_bench.json | sed -n "1,3p"; tail -12 gpurun_out/${T}_conv_probe.log
