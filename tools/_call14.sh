set -x
mkdir -p gpurun_out/c14
timeout 900 python -m pytest tests/test_backward_gpu.py tests/test_sharding_gpu.py tests/test_parity_14b_gpu.py tests/test_scheduler_gpu.py tests/test_model_gpu.py -m gpu -q > gpurun_out/c14/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c14/pytest.log
for mode in selective full; do
PRFL_CKPT=$mode timeout 600 python tools/prfl_step.py --blocks 8 --nograd 2 --steps 2 > gpurun_out/c14/prfl_n1_$mode.json 2>> gpurun_out/c14/prfl.err; echo "$mode rc=$?" >> gpurun_out/c14/prfl.err
done
timeout 600 python tools/prfl_step.py --blocks 8 --nograd 2 --steps 1 --profile > gpurun_out/c14/prfl_n1_profile.json 2>> gpurun_out/c14/prfl.err
