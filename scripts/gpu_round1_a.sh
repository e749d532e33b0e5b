set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tests/golden/make_golden.py > gpurun_out/golden.log 2>&1; echo "golden rc=$?"; tail -3 gpurun_out/golden.log
timeout 600 python benchmarks/micro_index.py > gpurun_out/micro_index.json 2> gpurun_out/micro_index.err; echo "micro rc=$?"; cat gpurun_out/micro_index.json | head -80
