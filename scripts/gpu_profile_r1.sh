set -x
mkdir -p gpurun_out
M="gpu__time_duration.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tc.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread"
# 1. launch list of the bench command (every launch, device time)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
# 2. full capture of the dominant kernel (SA1 FPS)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_cluster -s 2 -c 2 -f -o gpurun_out/r1_prof_fps python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-pipeline > gpurun_out/ncu_fps.log 2>&1; echo "ncu fps rc=$?"
# 3. key metrics of the other hot kernels in the bench
timeout 600 ncu --metrics $M --clock-control none -k regex:"sa_mlp_kernel|ball_query_kernel" -s 8 -c 12 --csv --log-file gpurun_out/r1_sa_bq_metrics.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-pipeline > gpurun_out/ncu_sa.log 2>&1; echo "ncu sa rc=$?"
# 4. attention stack kernels
timeout 600 ncu --metrics $M --clock-control none -k regex:"attention_kernel|linear_kernel" -s 120 -c 110 --csv --log-file gpurun_out/r1_attn_metrics.csv python benchmarks/micro_attn.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_kernel" -s 40 -c 1 -f -o gpurun_out/r1_prof_attn python benchmarks/micro_attn.py > gpurun_out/ncu_attn2.log 2>&1; echo "ncu attn full rc=$?"
# 5. numbers (never under a profiler)
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; echo "bench rc=$?"; cat gpurun_out/r1_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_reference.json 2> gpurun_out/r1_bench_reference.err; echo "ref rc=$?"; cat gpurun_out/r1_bench_reference.json
timeout 600 python benchmarks/micro_attn.py > gpurun_out/r1_micro_attn.json 2>/dev/null; echo "micro_attn rc=$?"
timeout 600 python benchmarks/micro_index.py > gpurun_out/r1_micro_index.json 2>/dev/null; echo "micro_index rc=$?"
ls -la gpurun_out | tail -20
