# round-2 late validation: GPU suite (new host-executor / hostq / heat tests included), e2e breakdown with and without chunked roots
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r2k_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2k_pytest_gpu.log
timeout 120 python tools/e2e_probe.py > gpurun_out/r2k_e2e_chunked.log 2>&1; echo "e2e rc=$?"
DEVA_B200_ROOT_CHUNK=0 timeout 120 python tools/e2e_probe.py > gpurun_out/r2k_e2e_onepiece.log 2>&1; echo "e2e0 rc=$?"
DEVA_B200_ROOT_CHUNK=4194304 timeout 120 python tools/e2e_probe.py > gpurun_out/r2k_e2e_chunk4m.log 2>&1
tail -4 gpurun_out/r2k_e2e_chunked.log gpurun_out/r2k_e2e_onepiece.log gpurun_out/r2k_e2e_chunk4m.log
