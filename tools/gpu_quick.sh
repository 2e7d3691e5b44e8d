# usage: gpurun -- 'bash tools/gpu_quick.sh "<pytest args>"'
mkdir -p gpurun_out/quick
timeout 1500 python -m pytest $1 -m gpu -x -q > gpurun_out/quick/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/quick/pytest.log
tail -25 gpurun_out/quick/pytest.log
