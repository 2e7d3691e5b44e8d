set -x
mkdir -p gpurun_out/r2b
ls oracle/_ref | head -5 > gpurun_out/r2b/ref_ls.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2b/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b/pytest.log
tail -30 gpurun_out/r2b/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2b/smoke.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2b/bench_ref.json 2> gpurun_out/r2b/bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b/bench.json 2> gpurun_out/r2b/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_packed -s 330 -c 1 -o gpurun_out/r2b/packed_steady python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2b/ncu.log 2>&1
tail -2 gpurun_out/r2b/smoke.log
