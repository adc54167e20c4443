set -x
mkdir -p gpurun_out/c15
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/c15/bench_n8.log 2>&1; echo "bench rc=$?" >> gpurun_out/c15/bench_n8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29752 tools/prfl_step.py --blocks 40 --nograd 2 --steps 1 --profile > gpurun_out/c15/prfl_n8_profile.json 2> gpurun_out/c15/prfl_n8_profile.err; echo "rc=$?" >> gpurun_out/c15/prfl_n8_profile.err
