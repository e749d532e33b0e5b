set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_modules.py -x -q -s -m gpu > gpurun_out/modules.log 2>&1; echo "modules rc=$?"; tail -40 gpurun_out/modules.log
