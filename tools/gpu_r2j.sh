set -x
mkdir -p gpurun_out/r2j
timeout 1500 python -m pytest tests/test_gpu_policy.py tests/test_gpu_verbs.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2j/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j/pytest.log
tail -40 gpurun_out/r2j/pytest.log
