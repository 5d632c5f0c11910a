python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/f64_probe.py 2>&1 | tee gpurun_out/r2_f64_probe.log
python tools/prep_probe.py 2>&1 | tee gpurun_out/r2_prep_probe_tri.log
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_prep_launches_c5_tri.csv python tools/prep_ncu.py c5 > /dev/null 2>&1
python tools/launch_agg.py gpurun_out/r2_prep_launches_c5_tri.csv 10
