set -x
cd $GRAFT_REPO_ROOT
TAG=$1; N=$2
run() {
  name=$1; shift
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --quick "$@" > gpurun_out/${TAG}_n${N}_${name}.json 2> gpurun_out/${TAG}_n${N}_${name}.err
  echo "bench $name rc=$?"
  tail -c 900 gpurun_out/${TAG}_n${N}_${name}.json
}
run cfg4_strong_sparse --workload config4_1m_256_sh3_b32 --views-total 32
run cfg4_strong_dense --workload config4_1m_256_sh3_b32 --views-total 32 --dense-exchange
