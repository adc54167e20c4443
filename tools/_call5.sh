set -x
mkdir -p gpurun_out/c5
# correctness of the new variants first
PRFL_ATTN_FWD=quad timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -m gpu -q -k "attn" > gpurun_out/c5/pytest_fwd_quad.log 2>&1; echo "rc=$?" >> gpurun_out/c5/pytest_fwd_quad.log
PRFL_ATTN_FWD=quad37 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attn_fwd" > gpurun_out/c5/pytest_fwd_quad37.log 2>&1; echo "rc=$?" >> gpurun_out/c5/pytest_fwd_quad37.log
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py tests/test_backward_gpu.py -m gpu -q -k "attn or gradients" > gpurun_out/c5/pytest_bwd_quad.log 2>&1; echo "rc=$?" >> gpurun_out/c5/pytest_bwd_quad.log
# timing, alternating variants (same box, back to back)
for rep in 1 2; do
  for v in base quad quad37; do
    PRFL_ATTN_FWD=$v timeout 300 python tools/fwd_ab.py 32760 40 >> gpurun_out/c5/fwd_ab.log 2>&1
  done
done
for v in base quad quad37; do
  PRFL_ATTN_FWD=$v timeout 300 python tools/fwd_ab.py 75600 5 >> gpurun_out/c5/fwd_ab.log 2>&1
  PRFL_ATTN_FWD=$v timeout 300 python tools/fwd_ab.py 32760 5 >> gpurun_out/c5/fwd_ab.log 2>&1
done
for rep in 1 2; do
  PRFL_ATTN_BWD_DQ=pair timeout 300 python tools/bwd_ab.py 32760 40 >> gpurun_out/c5/bwd_ab.log 2>&1
  PRFL_ATTN_BWD_DQ=quad timeout 300 python tools/bwd_ab.py 32760 40 >> gpurun_out/c5/bwd_ab.log 2>&1
done
PRFL_ATTN_BWD_DQ=pair timeout 300 python tools/bwd_ab.py 75600 5 >> gpurun_out/c5/bwd_ab.log 2>&1
PRFL_ATTN_BWD_DQ=quad timeout 300 python tools/bwd_ab.py 75600 5 >> gpurun_out/c5/bwd_ab.log 2>&1
