# tests + bench + one ncu --set full capture of a batched step (run through gpurun)
set -x
cd $GRAFT_REPO_ROOT
TAG=${1:-r2b}
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
( time python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/${TAG}_full -f \
   python scripts/profile_batch.py headline_1m_512_sh3 3 4 > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ls -la gpurun_out/
