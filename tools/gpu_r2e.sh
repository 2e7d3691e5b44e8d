set -x
mkdir -p gpurun_out/r2e
timeout 1200 python -m pytest tests/test_gpu_packed.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/r2e/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e/pytest.log
tail -5 gpurun_out/r2e/pytest.log
timeout 900 python tools/bench_variants.py air_hockey_training_environment_b200/libah_b200.so air_hockey_training_environment_b200/variants/*.so > gpurun_out/r2e/variants.txt 2>&1
cat gpurun_out/r2e/variants.txt
timeout 300 python tools/bench_kscale.py > gpurun_out/r2e/kscale.txt 2>&1; cat gpurun_out/r2e/kscale.txt
