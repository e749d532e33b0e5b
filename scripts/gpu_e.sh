set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python benchmarks/micro_index.py > gpurun_out/micro_index3.json 2> gpurun_out/micro_index3.err; echo "micro rc=$?"; tail -3 gpurun_out/micro_index3.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; echo "bench rc=$?"; cat gpurun_out/bench_r1c.json; tail -5 gpurun_out/bench_r1c.err
