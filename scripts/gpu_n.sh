mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -6
timeout 900 python benchmarks/micro_forward.py > gpurun_out/micro_forward.json 2> gpurun_out/micro_forward.err; echo "rc=$?"; cat gpurun_out/micro_forward.json; tail -5 gpurun_out/micro_forward.err
