set -x
mkdir -p gpurun_out/r2d
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2d/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d/pytest.log
tail -5 gpurun_out/r2d/pytest.log
timeout 300 python tools/bench_kscale.py > gpurun_out/r2d/kscale.txt 2>&1; cat gpurun_out/r2d/kscale.txt
timeout 900 python tools/bench_variants.py air_hockey_training_environment_b200/variants/libah_t64_*.so air_hockey_training_environment_b200/variants/libah_t128_o9.so > gpurun_out/r2d/variants.txt 2>&1
cat gpurun_out/r2d/variants.txt
