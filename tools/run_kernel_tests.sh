#!/bin/bash
# runs each GPU kernel test group in its own process so one faulting kernel cannot poison the rest
mkdir -p gpurun_out
for k in ${TESTS:-test_ln_mod test_rmsnorm_rope test_gemm_layouts test_gemm_epilogues test_attn_fwd test_patchify_unpatchify_cast test_a2a_pack test_sq_pool test_attn_bwd test_colsum_gate_bwd test_gemm_aux_outputs}; do
  echo "=== $k"
  timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "$k" --timeout 300 -p no:cacheprovider 2>&1 | tail -25
done
