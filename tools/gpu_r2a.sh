set -x
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a/smi.txt
nproc >> gpurun_out/r2a/smi.txt; lscpu | head -20 >> gpurun_out/r2a/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a/bench_ref.json 2> gpurun_out/r2a/bench_ref.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 330 -c 1 -o gpurun_out/r2a/step_steady python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2a/ncu.log 2>&1
tail -3 gpurun_out/r2a/pytest.log
