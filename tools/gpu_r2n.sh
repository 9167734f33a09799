# round-2 final: N=1 bench line, then the whole GPU suite, on the round's final code
cd $GRAFT_REPO_ROOT
timeout 70 python bench.py > gpurun_out/r2n_bench_n1.json 2> gpurun_out/r2n_bench_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2n_bench_n1.json
timeout 110 python -m pytest tests -m gpu -q --durations=8 2>&1 | tee gpurun_out/r2n_pytest_gpu.log | tail -14; echo "pytest rc=${PIPESTATUS[0]}"
