set -x
mkdir -p gpurun_out/c8
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c8/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c8/pytest.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/c8/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/c8/bench.log
# launch list of the same scoring step (after the plain run exited 0), then one full capture of the top kernels in isolation
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/c8/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-parity --no-gpu-baseline --no-prfl > gpurun_out/c8/ncu_launches.log 2>&1
timeout 900 python tools/prof_kernels.py > gpurun_out/c8/prof_plain.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'attn_fwd|attn_bwd|gemm_bf16|ln_mod|rmsnorm_rope' -c 12 -o gpurun_out/c8/prof python tools/prof_kernels.py > gpurun_out/c8/ncu_full.log 2>&1
ls -la gpurun_out/c8
