mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_backbone.py tests/test_gpu_ops.py -q -m gpu --tb=short 2>&1 | tail -15
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err; echo "bench rc=$?"; cat gpurun_out/bench_r1e.json; tail -5 gpurun_out/bench_r1e.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pipeline > gpurun_out/bench_r1e_nopipe.json 2> gpurun_out/bench_r1e_nopipe.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r1e_nopipe.json')); print('no-pipeline: ms_per_step', d['ms_per_step'], 'value', d['value'])"
