set -x
mkdir -p gpurun_out/c7
for v in quadp quadp12; do
PRFL_ATTN_FWD=$v timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -m gpu -q -k "attn_fwd or attention_forward" > gpurun_out/c7/pytest_$v.log 2>&1; echo "rc=$?" >> gpurun_out/c7/pytest_$v.log
done
for rep in 1 2; do
  for v in base quad12 quadp quadp12; do
    PRFL_ATTN_FWD=$v timeout 300 python tools/fwd_ab.py 32760 40 >> gpurun_out/c7/fwd_ab.log 2>&1
  done
done
for v in quad12 quadp quadp12; do
  PRFL_ATTN_FWD=$v timeout 300 python tools/fwd_ab.py 75600 5 >> gpurun_out/c7/fwd_ab.log 2>&1
done
timeout 600 python tools/sample_step.py --steps 3 > gpurun_out/c7/sample_step.json 2> gpurun_out/c7/sample_step.err
