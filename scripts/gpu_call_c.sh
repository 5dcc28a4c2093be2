# tests + bench + quick A/B runs of environment switches (run through gpurun): bash scripts/gpu_call_c.sh TAG "ENV1=.. ENV2=.." ...
set -x
cd $GRAFT_REPO_ROOT
TAG=${1:-r2c}; shift
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
( time python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"
i=0
for envs in "$@"; do
  i=$((i+1))
  echo "== $envs" >> gpurun_out/${TAG}_ab.log
  env $envs python bench.py --quick --steps 30 --warmup 5 >> gpurun_out/${TAG}_ab.log 2>&1
done
cat gpurun_out/${TAG}_ab.log | cut -c1-200
