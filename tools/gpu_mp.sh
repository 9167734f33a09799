# usage: bash tools/gpu_mp.sh N "ENV=.. ENV=.." ...   : one multi_probe run per env set
N=$1; shift
cd $GRAFT_REPO_ROOT
for envs in "$@"; do
  echo "== $envs"
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/multi_probe.py 2>&1 | grep -E "^\{|tile prof|rror" | cut -c1-400
done
