mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_golden.py tests/test_backbone.py -q -m gpu --tb=short 2>&1 | tail -6
EDA_FPS_DIRECT=0 timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_golden.py -q -m gpu --tb=short 2>&1 | tail -3
for d in 1 0; do EDA_FPS_DIRECT=$d timeout 600 python benchmarks/micro_index.py 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin)
print('direct=$d', {k: round(v['med_ms'],4) for k,v in d.items() if isinstance(v,dict) and k.startswith('fps_') and 'cl' not in k})"; done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench ms_per_step', d['ms_per_step'], 'value', d['value'], 'fps_kernel_ms', d['roofline']['kernel_ms'])"
