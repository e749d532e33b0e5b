set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_umma.py -x -q -s > gpurun_out/umma.log 2>&1; echo "umma rc=$?"; tail -25 gpurun_out/umma.log
