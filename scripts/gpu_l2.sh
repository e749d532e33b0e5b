python scripts/lin_ts.py
timeout 900 python -m pytest tests/test_attention.py -q -m gpu --tb=short 2>&1 | tail -5
timeout 600 python benchmarks/micro_attn.py 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin)
for k in ['ours_enc3_dec6','ours_enc3_dec6_graphed','linear_8192x288x288','linear_ln_8192','linear_2048x288x288']: print(k, d[k])"
EDA_LINEAR_SPLIT=1 python scripts/lin_ts.py
EDA_LINEAR_SPLIT=1 timeout 900 python -m pytest tests/test_attention.py -q -m gpu --tb=short 2>&1 | tail -3
EDA_LINEAR_SPLIT=1 timeout 600 python benchmarks/micro_attn.py 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin)
for k in ['ours_enc3_dec6_graphed','linear_ln_8192','linear_2048x288x288']: print('split', k, d[k])"
