mkdir -p gpurun_out/variants
timeout 900 python -m pytest tests/test_gpu_packed.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python tools/bench_variants.py air_hockey_training_environment_b200/libah_b200.so air_hockey_training_environment_b200/variants/*.so > gpurun_out/variants/variants.txt 2>&1
cat gpurun_out/variants/variants.txt
