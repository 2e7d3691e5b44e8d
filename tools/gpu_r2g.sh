set -x
mkdir -p gpurun_out/r2g
timeout 1500 python -m pytest tests/test_gpu_packed.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/r2g/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g/pytest.log
tail -4 gpurun_out/r2g/pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2g/bench20.json 2> gpurun_out/r2g/bench20.err
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2g/bench200.json 2> gpurun_out/r2g/bench200.err
timeout 600 python bench.py --steps 20 --warmup 5 --parts 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2g/bench20_p3.json 2> gpurun_out/r2g/bench20_p3.err
tail -2 gpurun_out/r2g/*.err
