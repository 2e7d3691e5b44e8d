set -x
mkdir -p gpurun_out/r2k
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2k/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k/pytest.log
tail -30 gpurun_out/r2k/pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2k/bench.json 2> gpurun_out/r2k/bench.err; tail -3 gpurun_out/r2k/bench.err
