mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -25
