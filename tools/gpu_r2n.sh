mkdir -p gpurun_out/r2n
nvidia-smi topo -m > gpurun_out/r2n/topo.txt 2>&1; nproc >> gpurun_out/r2n/topo.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2n/bench8.json 2> gpurun_out/r2n/bench8.err
tail -3 gpurun_out/r2n/bench8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2n/bench8_ref.json 2> gpurun_out/r2n/bench8_ref.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 5 --no-extras > gpurun_out/r2n/bench4.json 2> gpurun_out/r2n/bench4.err
