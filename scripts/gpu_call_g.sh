# quick parity subset + A/B list + targeted ncu: bash scripts/gpu_call_g.sh TAG "kernel regex" "ENV=.." ...
set -x
cd $GRAFT_REPO_ROOT
TAG=${1:-r2g}; shift
KRE=$1; shift
( time timeout 1200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
for envs in "$@"; do
  echo "== $envs" >> gpurun_out/${TAG}_ab.log
  env $envs timeout 300 python bench.py --quick --steps 30 --warmup 5 >> gpurun_out/${TAG}_ab.log 2>&1
done
cut -c1-200 gpurun_out/${TAG}_ab.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$KRE" -o gpurun_out/${TAG}_full -f \
   python scripts/profile_batch.py headline_1m_512_sh3 3 4 > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu rc=$?"
