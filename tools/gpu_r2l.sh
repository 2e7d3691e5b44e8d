set -x
mkdir -p gpurun_out/r2l
timeout 600 python -m pytest tests/test_gpu_peer.py -m gpu -x -q > gpurun_out/r2l/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l/pytest.log
tail -25 gpurun_out/r2l/pytest.log
for c in p2p nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-e2e --collective $c > gpurun_out/r2l/bench2_$c.json 2> gpurun_out/r2l/bench2_$c.err
tail -2 gpurun_out/r2l/bench2_$c.err
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r2l/bench1.json 2> gpurun_out/r2l/bench1.err
