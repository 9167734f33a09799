set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest_gpu.log
tail -5 gpurun_out/r2i_pytest_gpu.log
QB_SKIP_SMALL=1 QB_ENGINES=tile QB_REPS=3 python tools/tile_probe.py > gpurun_out/r2i_probe_p01.log 2>&1
QB_SKIP_SMALL=1 QB_ENGINES=tile QB_REPS=3 QB_P=0.5 python tools/tile_probe.py > gpurun_out/r2i_probe_p05.log 2>&1
cat gpurun_out/r2i_probe_p01.log gpurun_out/r2i_probe_p05.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
cat gpurun_out/r2i_bench.json; tail -5 gpurun_out/r2i_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2i_bench_ref.json 2> gpurun_out/r2i_bench_ref.err; echo "ref rc=$?"
cat gpurun_out/r2i_bench_ref.json
