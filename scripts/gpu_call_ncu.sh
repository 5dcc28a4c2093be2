# targeted ncu capture of one batched step: bash scripts/gpu_call_ncu.sh TAG "kernel regex" [ENV=..]
set -x
cd $GRAFT_REPO_ROOT
TAG=$1; KRE=$2; shift; shift
env "$@" timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$KRE" -o gpurun_out/${TAG}_full -f \
   python scripts/profile_batch.py headline_1m_512_sh3 3 4 > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu rc=$?"
