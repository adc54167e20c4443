set -x
mkdir -p gpurun_out/c1
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/c1/smi.txt
free -g | head -2 >> gpurun_out/c1/smi.txt; nproc >> gpurun_out/c1/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1/pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/c1/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/c1/bench.log
for g in 8 16 32 4 -8 -16 -4; do PRFL_GEMM_GROUP_M=$g timeout 300 python tools/gemm_sweep.py >> gpurun_out/c1/gemm_sweep.log 2>&1; done
timeout 300 python tools/gemm_sweep.py --cublas >> gpurun_out/c1/gemm_sweep.log 2>&1
for g in 8 16 -8; do PRFL_GEMM_GROUP_M=$g timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:gemm_bf16 --csv --log-file gpurun_out/c1/ncu_gemm_g$g.csv python tools/gemm_sweep.py --once > gpurun_out/c1/ncu_gemm_g$g.log 2>&1; done
