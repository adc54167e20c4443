set -x
mkdir -p gpurun_out/c3
timeout 300 python tools/_dbg_adamw.py > gpurun_out/c3/dbg_adamw.log 2>&1
timeout 600 python -m pytest tests/test_plugin_gpu.py tests/test_scheduler_gpu.py tests/test_sp_gpu.py -m gpu -q > gpurun_out/c3/pytest_new.log 2>&1; echo "rc=$?" >> gpurun_out/c3/pytest_new.log
# sanitizer over the kernel tests (small shapes): memcheck then racecheck, each bounded
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "test_ln_mod or test_rmsnorm_rope or test_gemm_epilogues or test_attn_fwd or test_attn_bwd or test_sq_pool or test_colsum or test_adamw" -p no:cacheprovider > gpurun_out/c3/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/c3/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "test_ln_mod_bwd or test_rmsnorm_rope_bwd or test_attn_fwd or test_attn_bwd or test_gemm_epilogues or test_sq_pool" -p no:cacheprovider > gpurun_out/c3/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/c3/sanitizer_racecheck.log
