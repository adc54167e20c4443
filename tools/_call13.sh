set -x
mkdir -p gpurun_out/c13
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/c13/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c13/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c13/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/c13/smoke.log
timeout 1200 python bench.py > gpurun_out/c13/bench_default.log 2>&1; echo "bench rc=$?" >> gpurun_out/c13/bench_default.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c13/bench_ref.log 2>&1; echo "ref rc=$?" >> gpurun_out/c13/bench_ref.log
