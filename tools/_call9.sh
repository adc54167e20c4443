set -x
mkdir -p gpurun_out/c9
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 tests/sp_check.py > gpurun_out/c9/sp_check_n2.log 2>&1; echo "rc=$?" >> gpurun_out/c9/sp_check_n2.log
