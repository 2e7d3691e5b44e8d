# Final evidence of the round: all GPU tests, smoke, the driver's two bench commands, the launch list, K scaling.
tag=${1:-r2zz}
mkdir -p gpurun_out/$tag
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/$tag/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/$tag/pytest.log
tail -3 gpurun_out/$tag/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/$tag/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/$tag/smoke.log
tail -2 gpurun_out/$tag/smoke.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/$tag/bench_ref.json 2> gpurun_out/$tag/bench_ref.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/$tag/bench.json 2> gpurun_out/$tag/bench.err
timeout 300 python tools/bench_kscale.py > gpurun_out/$tag/kscale.txt 2>&1; cat gpurun_out/$tag/kscale.txt
head -c 300 gpurun_out/$tag/bench.json
