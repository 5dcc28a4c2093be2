set -x
cd $GRAFT_REPO_ROOT
TAG=$1; N=$2
run() {
  name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --quick "$@" > gpurun_out/${TAG}_n${N}_${name}.json 2> gpurun_out/${TAG}_n${N}_${name}.err
  echo "bench $name rc=$?"
  tail -c 700 gpurun_out/${TAG}_n${N}_${name}.json
}
run auto_sparse --allreduce auto
run auto_dense --allreduce auto --dense-exchange
