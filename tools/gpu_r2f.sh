set -x
mkdir -p gpurun_out/r2f
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f/pytest.log
tail -4 gpurun_out/r2f/pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f/bench.json 2> gpurun_out/r2f/bench.err
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2f/bench200.json 2> gpurun_out/r2f/bench200.err
timeout 600 python bench.py --steps 200 --warmup 5 --parts 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2f/bench200_p1.json 2> gpurun_out/r2f/bench200_p1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_packed -s 330 -c 1 -o gpurun_out/r2f/packed_steady python bench.py --steps 20 --warmup 5 --parts 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2f/ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2f/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2f/ncu_launches.log 2>&1
