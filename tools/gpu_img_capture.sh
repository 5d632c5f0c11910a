python bench.py --workload mnist_img --no-modes > gpurun_out/r1v_bench_mnist_img.json 2> gpurun_out/r1v_bench_mnist_img.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r1v_launches_mnist_img.csv python bench.py --workload mnist_img --rows 4096 --steps 1 --warmup 1 --only-logprob > gpurun_out/r1v_ncu1.log 2>&1
cut -c1-2500 gpurun_out/r1v_bench_mnist_img.json; tail -5 gpurun_out/r1v_bench_mnist_img.err
