set -x
mkdir -p gpurun_out/c16
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29761 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/c16/bench_n4.log 2>&1; echo "bench rc=$?" >> gpurun_out/c16/bench_n4.log
