set -x
mkdir -p gpurun_out
N=1048576
run() { echo "## $*" >> gpurun_out/g3_perf.txt; env "$@" >> gpurun_out/g3_perf.txt 2>&1; }
for w in 4000 5500 9000 12000; do run CCU_JIT_SEGWEIGHT=$w python tools/prof_one.py quad_jac 1 0 0 0 $N 3; done
for s in 1500 3500; do run CCU_JIT_SEG=$s python tools/prof_one.py quad_jac 1 0 0 0 $N 3; done
run CCU_JIT_INTERLEAVE=8 python tools/prof_one.py quad_jac 1 0 0 0 $N 3
run CCU_JIT_INTERLEAVE=32 python tools/prof_one.py quad_jac 1 0 0 0 $N 3
run CCU_JIT_REGVALS=90 python tools/prof_one.py quad_jac 1 0 0 0 $N 3
run CCU_JIT_REGVALS=110 python tools/prof_one.py quad_jac 1 0 0 0 $N 3
run CCU_JIT_SEGWEIGHT=0 python tools/prof_one.py quad 1 0 0 0 $N 3
run CCU_JIT_SEGWEIGHT=14000 python tools/prof_one.py quad 1 0 0 0 $N 3
run CCU_JIT_SEGWEIGHT=3500 python tools/prof_one.py quad 1 0 0 0 $N 3
run CCU_JIT_INTERLEAVE=16 python tools/prof_one.py quad 1 0 0 0 $N 3
run CCU_JIT_INTERLEAVE=16 python tools/prof_one.py cartpole 1 0 0 0 8388608 3
run python tools/prof_one.py cartpole 1 8000 128 4 8388608 3
run python tools/prof_one.py cartpole 1 8000 256 1 8388608 3
export CASADI_CUDA_LIB=$PWD/casadi_b200/lib/libcasadi_cuda.so
B=tests/integration/_build/bin/cuda_bench
for m in pageable pinned; do
  echo "## cuda_bench quad_ms 4000000 $m" >> gpurun_out/g3_e2e.txt
  $B quad_ms 4000000 2 1 $m >> gpurun_out/g3_e2e.txt 2>&1
done
$B cartpole 1000000 3 1 pageable >> gpurun_out/g3_e2e.txt 2>&1
$B mc 1000000 2 1 pageable reduce >> gpurun_out/g3_e2e.txt 2>&1
$B kkt_ldl 1000000 2 1 pageable >> gpurun_out/g3_e2e.txt 2>&1
