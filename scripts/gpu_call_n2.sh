# multi-GPU check of the exchange kernels: bash scripts/gpu_call_n2.sh TAG N
set -x
cd $GRAFT_REPO_ROOT
TAG=$1; N=$2
( time timeout 900 python -m pytest tests/test_p2p_gpu.py -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
run() {  # name, extra args
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --quick "$@" > gpurun_out/${TAG}_n${N}_${name}.json 2> gpurun_out/${TAG}_n${N}_${name}.err
  echo "bench $name rc=$?"
  tail -c 700 gpurun_out/${TAG}_n${N}_${name}.json
}
run p2p_sparse --allreduce p2p
run p2p_dense --allreduce p2p --dense-exchange
run mc_sparse --allreduce mc
run mc_dense --allreduce mc --dense-exchange
