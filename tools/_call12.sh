set -x
mkdir -p gpurun_out/c12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 tests/sp_check.py > gpurun_out/c12/sp_check_n2.log 2>&1; echo "rc=$?" >> gpurun_out/c12/sp_check_n2.log
for mode in p2p nccl p2p nccl; do
PRFL_ULYSSES=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29742 tools/prfl_step.py --blocks 6 --nograd 0 --steps 2 > gpurun_out/c12/prfl_n2_$mode.json 2>> gpurun_out/c12/prfl.err; echo "$mode rc=$?" >> gpurun_out/c12/prfl.err
mv gpurun_out/c12/prfl_n2_$mode.json gpurun_out/c12/prfl_n2_${mode}_$(date +%s).json
done
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "merge" > gpurun_out/c12/pytest_merge.log 2>&1
