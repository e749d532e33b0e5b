#!/bin/bash
# One parametrised gpurun job (replaces the per-call scripts of round 1):  scripts/gpu_job.sh TAG STEP [STEP ...]
#   test       pytest -m gpu (whole suite, not stopping at the first failure)
#   bench      python bench.py (both arms), JSON lines into gpurun_out/
#   launches   ncu launch list (gpu__time_duration) of one eager training step
#   metrics    ncu key metrics of every kernel family of one eager training step
#   sweepncu   ncu DRAM / L2 / tensor-pipe counters of the configs[4] sweep points
# Numbers printed by anything run under ncu are never bench values.
set -x
TAG=$1; shift
mkdir -p gpurun_out
M="gpu__time_duration.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread"
for step in "$@"; do
  case $step in
    test)
      timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -60 gpurun_out/${TAG}_pytest.log ;;
    testx)
      timeout 1200 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/${TAG}_pytest.log ;;
    smoke)
      timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/${TAG}_smoke.log ;;
    bench)
      timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
      timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"; tail -c 600 gpurun_out/${TAG}_bench_reference.json ;;
    benchquick)
      timeout 600 python bench.py --steps 10 --warmup 3 --quick-sweep --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches_train.csv python benchmarks/train_once.py 1 > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "launch list rc=$?"
      python scripts/ncu_csv_summary.py gpurun_out/${TAG}_launches_train.csv gpurun_out/${TAG}_launches_train_summary.json ;;
    metrics)
      timeout 1200 ncu --metrics $M --clock-control none -k regex:"_kernel" -c 1500 --csv --log-file gpurun_out/${TAG}_kernel_metrics.csv python benchmarks/train_once.py 1 > gpurun_out/${TAG}_ncu_metrics.log 2>&1; echo "metrics rc=$?"
      python scripts/ncu_csv_summary.py gpurun_out/${TAG}_kernel_metrics.csv gpurun_out/${TAG}_kernel_metrics.json ;;
    sweepncu)
      timeout 1200 ncu --metrics $M --clock-control none -k regex:"fps_cluster|ball_query|attention_kernel" -c 400 --csv --log-file gpurun_out/${TAG}_sweep_ncu.csv python benchmarks/kernels.py > gpurun_out/${TAG}_sweep_under_ncu.log 2>&1; echo "sweep ncu rc=$?"
      python scripts/ncu_csv_summary.py gpurun_out/${TAG}_sweep_ncu.csv gpurun_out/${TAG}_sweep_ncu.json ;;
    full)
      # one `--set full` capture per roofline kernel (the dominant launch of each), source-correlated
      for k in linear wgrad rows_gemm fps ballq attention sa_mlp; do
        cmd="python scripts/kernels_once.py $k"
        case $k in linear) re=linear_kernel;; wgrad) re=wgrad_tc_kernel;; rows_gemm) re=rows_gemm_tc_kernel; cmd="python scripts/rows_gemm_compare.py";; fps) re=fps_cluster_kernel;; ballq) re=ball_query_kernel;; attention) re=attention_kernel;; sa_mlp) re=sa_mlp_kernel;; esac
        timeout 600 ncu --set full --clock-control none --import-source on -k regex:$re -s 3 -c 1 -f -o gpurun_out/${TAG}_full_$k $cmd > gpurun_out/${TAG}_full_$k.log 2>&1; echo "full $k rc=$?"
      done ;;
    benchlaunches)
      # launch list of the bench command itself (graph replays show up as their kernels)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "bench launch list rc=$?"
      python scripts/ncu_csv_summary.py gpurun_out/${TAG}_launches_bench.csv gpurun_out/${TAG}_launches_bench_summary.json ;;
    *) echo "unknown step $step" ;;
  esac
done
ls -la gpurun_out | tail -30
