mkdir -p gpurun_out
python scripts/lin_ts.py
timeout 900 python -m pytest tests/test_attention.py -q -m gpu --tb=short 2>&1 | tail -15
timeout 600 python benchmarks/micro_attn.py > gpurun_out/micro_attn3.json 2> gpurun_out/micro_attn3.err; echo "micro_attn rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/micro_attn3.json'))
for k,v in d.items(): print(k, v)
"; tail -5 gpurun_out/micro_attn3.err
