# round-2 late: engine + hostq GPU tests, e2e breakdown with the root-width plan, N=1 bench line
cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_engine.py tests/test_hostq.py tests/test_full_size_gpu.py -m gpu -x -q > gpurun_out/r2l_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2l_pytest_gpu.log
timeout 100 python tools/e2e_probe.py > gpurun_out/r2l_e2e_plan.log 2>&1; echo "e2e rc=$?"; tail -7 gpurun_out/r2l_e2e_plan.log
DEVA_B200_ROOT_PLAN=0 timeout 100 python tools/e2e_probe.py > gpurun_out/r2l_e2e_noplan.log 2>&1; tail -7 gpurun_out/r2l_e2e_noplan.log | head -4
timeout 400 python bench.py > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_bench_n1.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2l_bench_n1.json
