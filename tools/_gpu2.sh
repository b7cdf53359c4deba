set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fast_path or flagged or every_specialisation or operator_set or agree_bitwise or full_size" > gpurun_out/g2_pytest.txt 2>&1
tail -5 gpurun_out/g2_pytest.txt
N=1048576
run() { echo "## $*" >> gpurun_out/g2_perf.txt; env "$@" >> gpurun_out/g2_perf.txt 2>&1; }
for t in quad_jac quad cartpole quad_fwd quad_adj rocket_hess mc; do
  NN=$N; if [ $t = cartpole ]; then NN=8388608; fi
  run python tools/prof_one.py $t 1 0 0 0 $NN 3
  run CCU_JIT_FASTOPS=0 python tools/prof_one.py $t 1 0 0 0 $NN 3
done
