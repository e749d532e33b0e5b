set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench rc=$?"; cat gpurun_out/bench_r1.json; tail -5 gpurun_out/bench_r1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_cluster -s 3 -c 1 -f -o gpurun_out/prof_fps python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_fps.log 2>&1; echo "ncu fps rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sa_mlp_kernel -s 6 -c 2 -f -o gpurun_out/prof_sa python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_sa.log 2>&1; echo "ncu sa rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ball_query -s 6 -c 2 -f -o gpurun_out/prof_bq python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bq.log 2>&1; echo "ncu bq rc=$?"
ls -la gpurun_out
