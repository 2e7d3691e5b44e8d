mkdir -p gpurun_out/r2o
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv > gpurun_out/r2o/smi_before.txt
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --warmup 5 --no-extras --no-e2e $2 > gpurun_out/r2o/$3.json 2> gpurun_out/r2o/$3.err; }
run 29531 "--steps 20" b8_s20
run 29532 "--steps 20 --collective nccl" b8_s20_nccl
run 29533 "--steps 200" b8_s200
run 29534 "--steps 20 --parts 1" b8_s20_p1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv > gpurun_out/r2o/smi_after.txt
