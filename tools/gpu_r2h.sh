set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest_gpu.log
tail -5 gpurun_out/r2h_pytest_gpu.log
QB_SKIP_SMALL=1 QB_ENGINES=tile QB_REPS=3 python tools/tile_probe.py > gpurun_out/r2h_probe_p01.log 2>&1
QB_SKIP_SMALL=1 QB_ENGINES=tile QB_REPS=3 QB_P=0.5 python tools/tile_probe.py > gpurun_out/r2h_probe_p05.log 2>&1
cat gpurun_out/r2h_probe_p01.log gpurun_out/r2h_probe_p05.log
QB_SKIP_SMALL=1 QB_ENGINES=tile QB_REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2h_launches.csv python tools/tile_probe.py > gpurun_out/r2h_ncu_list.log 2>&1
QB_SKIP_SMALL=1 QB_ENGINES=tile QB_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_step --launch-skip 120 --launch-count 8 -o gpurun_out/r2h_k_tile_step python tools/tile_probe.py > gpurun_out/r2h_ncu_full.log 2>&1
tail -3 gpurun_out/r2h_ncu_full.log
