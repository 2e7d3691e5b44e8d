mkdir -p gpurun_out/r2p
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --warmup 5 --no-extras --no-e2e $2 > gpurun_out/r2p/$3.json 2> gpurun_out/r2p/$3.err; }
run 29531 "--steps 20" b8_p2p_1
run 29532 "--steps 20 --collective nccl" b8_nccl_1
run 29533 "--steps 20" b8_p2p_2
run 29534 "--steps 20 --collective nccl" b8_nccl_2
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r2p/b1.json 2>/dev/null
