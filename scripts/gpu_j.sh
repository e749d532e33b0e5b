mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_attention.py -q -m gpu --tb=short 2>&1 | tail -30
timeout 600 python benchmarks/micro_attn.py > gpurun_out/micro_attn2.json 2> gpurun_out/micro_attn2.err; echo "micro_attn rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/micro_attn2.json'))
for k,v in d.items(): print(k, v)
"; tail -5 gpurun_out/micro_attn2.err
