set -x
cd $GRAFT_REPO_ROOT
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=25 ) > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
( time python bench.py ) > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2_ncu_bench.log 2>&1
echo "ncu rc=$?"
tail -5 gpurun_out/r2_pytest_gpu.log
