set -x
mkdir -p gpurun_out/c6
for v in quad quad12 quad0; do
PRFL_ATTN_FWD=$v timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attn_fwd" > gpurun_out/c6/pytest_$v.log 2>&1; echo "rc=$?" >> gpurun_out/c6/pytest_$v.log
done
for rep in 1 2; do
  for v in base quadw quad quad12 quad0 quad37; do
    PRFL_ATTN_FWD=$v timeout 300 python tools/fwd_ab.py 32760 40 >> gpurun_out/c6/fwd_ab.log 2>&1
  done
done
for v in base quad quad12; do
  PRFL_ATTN_FWD=$v timeout 300 python tools/fwd_ab.py 75600 5 >> gpurun_out/c6/fwd_ab.log 2>&1
done
