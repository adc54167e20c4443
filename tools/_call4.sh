set -x
mkdir -p gpurun_out/c4
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c4/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4/pytest.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c4/bench_n2.log 2>&1; echo "bench rc=$?" >> gpurun_out/c4/bench_n2.log
