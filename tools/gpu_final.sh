# Final-tree evidence run (one B200): tests, smoke, the driver's two bench commands, the launch list, then an A/B of
# library variants (tools/bench_variants.py).  usage: gpurun -- 'bash tools/gpu_final.sh <tag> [variant .so ...]'
tag=${1:-r2y}; shift
mkdir -p gpurun_out/$tag
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/$tag/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/$tag/pytest.log
tail -3 gpurun_out/$tag/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/$tag/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/$tag/smoke.log
tail -2 gpurun_out/$tag/smoke.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/$tag/bench_ref.json 2> gpurun_out/$tag/bench_ref.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/$tag/bench.json 2> gpurun_out/$tag/bench.err
CMD="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extras"
timeout 600 $CMD > gpurun_out/$tag/bench_plain.json 2> gpurun_out/$tag/bench_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/$tag/launches.csv $CMD > gpurun_out/$tag/ncu_launches.log 2>&1
if [ $# -gt 0 ]; then timeout 600 python tools/bench_variants.py air_hockey_training_environment_b200/libah_b200.so "$@" > gpurun_out/$tag/variants.txt 2>&1; cat gpurun_out/$tag/variants.txt; fi
head -c 600 gpurun_out/$tag/bench.json
