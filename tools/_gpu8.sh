set -x
mkdir -p gpurun_out
# 1. hess_lag / LDL: what bounds them?  plans x rematerialisation, then launch lists of the automatic plans
for W in 0 32; do
  CCU_JIT_REMAT=$W SWEEP_PLANS="1200:256:2,2500:128:2,2500:128:3,1800:128:2" timeout 200 python tools/sweep_plans.py rocket_hess >> gpurun_out/g8_sweep.jsonl 2>> gpurun_out/g8_sweep.err
done
for W in 0 24; do
  CCU_JIT_REMAT=$W SWEEP_PLANS="600:256:2,1200:256:2,1200:128:2,1200:128:4" timeout 200 python tools/sweep_plans.py kkt_ldl >> gpurun_out/g8_sweep.jsonl 2>> gpurun_out/g8_sweep.err
done
cut -c1-220 gpurun_out/g8_sweep.jsonl
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
for T in rocket_hess quad_adj; do
  timeout 200 ncu --metrics $M --clock-control none -k regex:ccu_seg --csv --log-file gpurun_out/g8_launches_$T.csv python tools/prof_one.py $T 1 0 0 0 1048576 2 > gpurun_out/g8_prof_$T.json 2> gpurun_out/g8_prof_$T.err
done
timeout 200 ncu --set full --clock-control none -k regex:ccu_seg -s 13 -c 2 -o gpurun_out/g8_full_rocket -f python tools/prof_one.py rocket_hess 1 0 0 0 1048576 2 > /dev/null 2> gpurun_out/g8_full_rocket.err
# 2. the whole GPU suite
timeout 900 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/g8_pytest.txt 2>&1
tail -14 gpurun_out/g8_pytest.txt
# 3. the bench as the driver runs it
timeout 800 python bench.py > gpurun_out/g8_bench.json 2> gpurun_out/g8_bench.err
tail -25 gpurun_out/g8_bench.err
timeout 300 python bench.py --impl reference > gpurun_out/g8_bench_ref.json 2> gpurun_out/g8_bench_ref.err
