# usage: bash tools/gpu_multi.sh N [steps] [warmup]   (on a box with N GPUs)
N=$1; K=${2:-3}; W=${3:-3}
cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/r2_multi_gpu_check_n$N.jsonl 2> gpurun_out/r2_multi_gpu_check_n$N.err; echo "check rc=$?"
cat gpurun_out/r2_multi_gpu_check_n$N.jsonl; tail -3 gpurun_out/r2_multi_gpu_check_n$N.err
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps $K --warmup $W > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"
cat gpurun_out/r2_bench_n$N.json; tail -5 gpurun_out/r2_bench_n$N.err
