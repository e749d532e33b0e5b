set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention.py tests/test_backbone.py -q -m gpu 2>&1 | tail -80 > gpurun_out/attn_tests.log; cat gpurun_out/attn_tests.log
