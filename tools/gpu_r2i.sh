set -x
mkdir -p gpurun_out/r2i
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i/pytest.log
tail -4 gpurun_out/r2i/pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2i/bench20.json 2> gpurun_out/r2i/bench20.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > gpurun_out/r2i/bench2.json 2> gpurun_out/r2i/bench2.err
tail -3 gpurun_out/r2i/bench2.err
