mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -5
timeout 600 python benchmarks/micro_index.py 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin)
print({k: round(v['med_ms'],4) for k,v in d.items() if isinstance(v,dict) and k.startswith('fps_') and 'cl' not in k})"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'fps_kernel_ms', d['roofline']['kernel_ms'])"
