# round-2 late: the device pool's one-block pop (small pools) - pool tests against brute force, the host-executor apps on it
cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_hostq.py tests/test_dropin.py -m gpu -x -q --durations=14 > gpurun_out/r2m_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -18 gpurun_out/r2m_pytest_gpu.log
