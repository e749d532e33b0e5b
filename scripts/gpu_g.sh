mkdir -p gpurun_out
timeout 300 python scripts/debug_attn.py > gpurun_out/debug_attn.log 2>&1; cat gpurun_out/debug_attn.log | tail -30
timeout 600 python -m pytest tests/test_attention.py tests/test_backbone.py -q -m gpu --tb=line 2>&1 | tail -30
