# multi-GPU check: bash scripts/gpu_call_n.sh TAG N
set -x
cd $GRAFT_REPO_ROOT
TAG=$1; N=$2
( time timeout 900 python -m pytest tests/test_p2p_gpu.py tests/test_extra_gpu.py tests/test_renderer_gpu.py -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
for ar in auto nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --quick --allreduce $ar > gpurun_out/${TAG}_n${N}_${ar}.json 2> gpurun_out/${TAG}_n${N}_${ar}.err
  echo "bench $ar rc=$?"
  tail -c 600 gpurun_out/${TAG}_n${N}_${ar}.json
done
