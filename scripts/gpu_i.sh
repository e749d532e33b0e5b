mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; echo "bench rc=$?"; cat gpurun_out/bench_r1d.json; tail -5 gpurun_out/bench_r1d.err
timeout 600 python benchmarks/micro_attn.py > gpurun_out/micro_attn.json 2> gpurun_out/micro_attn.err; echo "micro_attn rc=$?"; cat gpurun_out/micro_attn.json; tail -5 gpurun_out/micro_attn.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_cluster -s 2 -c 2 -f -o gpurun_out/prof_fps_r1d python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_fps.log 2>&1; echo "ncu fps rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_kernel|linear_kernel" -s 20 -c 12 -f -o gpurun_out/prof_attn_r1d python benchmarks/micro_attn.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
ls -la gpurun_out | tail -12
