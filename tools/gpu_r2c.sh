set -x
mkdir -p gpurun_out/r2c
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c/pytest.log
tail -5 gpurun_out/r2c/pytest.log
timeout 900 python tools/bench_variants.py air_hockey_training_environment_b200/libah_b200.so air_hockey_training_environment_b200/variants/*.so > gpurun_out/r2c/variants.txt 2>&1
cat gpurun_out/r2c/variants.txt
