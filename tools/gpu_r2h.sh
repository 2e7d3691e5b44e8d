set -x
mkdir -p gpurun_out/r2h
nvidia-smi topo -m > gpurun_out/r2h/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2h/bench2.json 2> gpurun_out/r2h/bench2.err
tail -5 gpurun_out/r2h/bench2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2h/bench2_ref.json 2> gpurun_out/r2h/bench2_ref.err
