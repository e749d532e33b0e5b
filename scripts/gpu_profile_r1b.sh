set -x
mkdir -p gpurun_out
M="gpu__time_duration.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread"
# 1. launch list of ONE eager training step (the second of two), every kernel, device time
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r1_launches_train.csv python benchmarks/train_once.py 1 > gpurun_out/ncu_train_launch.log 2>&1; echo "launch list rc=$?"
# 2. key metrics of the backward kernels
timeout 900 ncu --metrics $M --clock-control none -k regex:"attention_backward|wgrad_kernel|rows_gemm|layernorm_backward|sa_pool|bn_relu|sa_gather|sa_scatter" -c 400 --csv --log-file gpurun_out/r1_bwd_metrics.csv python benchmarks/train_once.py 1 > gpurun_out/ncu_bwd.log 2>&1; echo "ncu bwd rc=$?"
# 3. one full capture of the largest attention-backward launch (vis self-attention, rows = keys)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_backward" -s 3 -c 1 -f -o gpurun_out/r1_prof_attn_bwd python benchmarks/train_once.py 1 > gpurun_out/ncu_attn_bwd.log 2>&1; echo "ncu attn bwd full rc=$?"
# 4. numbers (never under a profiler)
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; echo "bench rc=$?"; cat gpurun_out/r1_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_reference.json 2> gpurun_out/r1_bench_reference.err; echo "ref rc=$?"; cat gpurun_out/r1_bench_reference.json
timeout 600 python benchmarks/micro_train.py 2>/dev/null | grep -v "^NCCL" > gpurun_out/r1_micro_train.json; echo "micro_train rc=$?"
timeout 600 python benchmarks/micro_bwd.py 2>/dev/null > gpurun_out/r1_micro_bwd.json; echo "micro_bwd rc=$?"
timeout 600 python benchmarks/micro_forward.py 2>/dev/null > gpurun_out/r1_micro_forward.json; echo "micro_forward rc=$?"
timeout 600 python benchmarks/profile_train.py gpurun_out/r1_profile_train.json > /dev/null 2>&1; echo "profile_train rc=$?"
ls -la gpurun_out | tail -20
