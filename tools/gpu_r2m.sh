mkdir -p gpurun_out/r2m
for i in 1 2 3; do for c in p2p nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-e2e --collective $c > gpurun_out/r2m/b_${c}_$i.json 2> /dev/null
done; done
for i in 1 2 3; do timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r2m/b_one_$i.json 2>/dev/null; done
