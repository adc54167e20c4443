set -x
mkdir -p gpurun_out/c2
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c2/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2/pytest.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/c2/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/c2/bench.log
for rep in 1 2; do
  timeout 300 python tools/gemm_sweep.py >> gpurun_out/c2/gemm_sweep.log 2>&1
  PRFL_GEMM_GROUP_M=8 timeout 300 python tools/gemm_sweep.py >> gpurun_out/c2/gemm_sweep.log 2>&1
done
timeout 300 python tools/gemm_sweep.py --M 9450 >> gpurun_out/c2/gemm_sweep.log 2>&1
PRFL_GEMM_GROUP_M=8 timeout 300 python tools/gemm_sweep.py --M 9450 >> gpurun_out/c2/gemm_sweep.log 2>&1
timeout 600 python tools/kernel_sweep.py --sections cross > gpurun_out/c2/cross_sweep.json 2> gpurun_out/c2/cross_sweep.err
