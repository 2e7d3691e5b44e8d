# Round-2 evidence run (one B200): tests, the driver's bench commands, then the ncu captures -- each only after the same
# command exited 0 without ncu.  Outputs under gpurun_out/<tag>/; summaries are produced from them by
# tools/make_profile_summary.py, tools/ncu_regions.py and tools/sass_path.py and committed under profiles/.
tag=${1:-r2z}
mkdir -p gpurun_out/$tag
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/$tag/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/$tag/pytest.log
tail -3 gpurun_out/$tag/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/$tag/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/$tag/smoke.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/$tag/bench_ref.json 2> gpurun_out/$tag/bench_ref.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/$tag/bench.json 2> gpurun_out/$tag/bench.err
CMD="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extras"
timeout 600 $CMD > gpurun_out/$tag/bench_plain.json 2> gpurun_out/$tag/bench_plain.err && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_packed -s 660 -c 1 -o gpurun_out/$tag/packed_steady $CMD > gpurun_out/$tag/ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/$tag/launches.csv $CMD > gpurun_out/$tag/ncu_launches.log 2>&1
tail -2 gpurun_out/$tag/smoke.log
