set -x
mkdir -p gpurun_out/c11
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731 tests/sp_check.py > gpurun_out/c11/sp_check_n8.log 2>&1; echo "rc=$?" >> gpurun_out/c11/sp_check_n8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29732 tools/prfl_step.py --blocks 40 --nograd 2 --steps 1 --profile > gpurun_out/c11/prfl_n8_profile.json 2> gpurun_out/c11/prfl_n8_profile.err; echo "rc=$?" >> gpurun_out/c11/prfl_n8_profile.err
