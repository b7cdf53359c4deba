set -x
mkdir -p gpurun_out
export CASADI_CUDA_LIB=$PWD/casadi_b200/lib/libcasadi_cuda.so
# 0. the plugin binaries must exit (static thread pool used to hang them at exit)
( time timeout 300 tests/integration/_build/bin/test_cuda_map ) > gpurun_out/g7_integration.txt 2>&1; echo "rc=$?" >> gpurun_out/g7_integration.txt
tail -4 gpurun_out/g7_integration.txt
( time timeout 300 tests/integration/_build/bin/cuda_bench quad_ms 2000000 2 1 pageable ) > gpurun_out/g7_cuda_bench.txt 2>&1; echo "rc=$?" >> gpurun_out/g7_cuda_bench.txt
tail -5 gpurun_out/g7_cuda_bench.txt
# 1. rematerialisation on the scratch-bound tapes
for W in 24 32; do
  CCU_JIT_REMAT=$W SWEEP_PLANS="2500:128:2" timeout 300 python tools/sweep_plans.py quad_adj rocket_hess quad_fwd kkt_qr >> gpurun_out/g7_remat.jsonl 2>> gpurun_out/g7_remat.err
done
CCU_JIT_REMAT=16 SWEEP_PLANS="2500:128:2,4400:128:2" timeout 300 python tools/sweep_plans.py rocket_hess quad_adj >> gpurun_out/g7_remat.jsonl 2>> gpurun_out/g7_remat.err
CCU_JIT_REMAT=24 SWEEP_PLANS="4400:128:2" timeout 200 python tools/sweep_plans.py quad_jac >> gpurun_out/g7_remat.jsonl 2>> gpurun_out/g7_remat.err
cat gpurun_out/g7_remat.jsonl | cut -c1-200
# 2. launch list with DRAM traffic of the headline Jacobian (automatic plan)
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
timeout 300 ncu --metrics $M --clock-control none -k regex:ccu_seg --csv --log-file gpurun_out/g7_launches_quad_jac.csv python tools/prof_one.py quad_jac 1 0 0 0 1048576 2 > gpurun_out/g7_prof_quad_jac.json 2> gpurun_out/g7_prof_quad_jac.err
# 3. one full capture of two Jacobian segments in the middle of the chain
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ccu_seg -s 30 -c 2 -o gpurun_out/g7_full_jac -f python tools/prof_one.py quad_jac 1 0 0 0 1048576 2 > /dev/null 2> gpurun_out/g7_full_jac.err
# 4. compute-sanitizer: memcheck + racecheck on a multi-segment specialised tape (ring + spill rows) and on the interpreter
S=/usr/local/cuda/bin/compute-sanitizer
timeout 300 $S --tool memcheck --print-limit 5 python tools/prof_one.py quad_jac 1 0 0 0 2048 1 > gpurun_out/g7_memcheck_jit.txt 2>&1
timeout 300 $S --tool racecheck --print-limit 5 python tools/prof_one.py quad_jac 1 0 0 0 2048 1 > gpurun_out/g7_racecheck_jit.txt 2>&1
timeout 300 $S --tool memcheck --print-limit 5 python tools/prof_interp.py quad1_jac 8192 > gpurun_out/g7_memcheck_interp.txt 2>&1
timeout 300 $S --tool racecheck --print-limit 5 python tools/prof_interp.py quad1_jac 8192 > gpurun_out/g7_racecheck_interp.txt 2>&1
tail -3 gpurun_out/g7_*check*.txt
