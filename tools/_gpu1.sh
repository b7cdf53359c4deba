set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/g1_smi.txt
./tools/fp64_ilp > gpurun_out/g1_fp64_ilp.jsonl 2>&1
N=1048576
run() { echo "## $*" >> gpurun_out/g1_jac.txt; env "$@" >> gpurun_out/g1_jac.txt 2>&1; }
run python tools/prof_one.py quad_jac 1 0 0 0 $N 3
run CCU_JIT_EXPERIMENT=1 python tools/prof_one.py quad_jac 1 0 0 0 $N 3
run CCU_JIT_EXPERIMENT=3 python tools/prof_one.py quad_jac 1 0 0 0 $N 3
run python tools/prof_one.py quad_jac 1 2500 128 3 $N 3
run python tools/prof_one.py quad_jac 1 2500 64 4 $N 3
run CCU_JIT_SPILL=-1 python tools/prof_one.py quad_jac 1 2500 64 4 $N 3
run CCU_JIT_INTERLEAVE=8 python tools/prof_one.py quad_jac 1 0 0 0 $N 3
run CCU_JIT_INTERLEAVE=16 python tools/prof_one.py quad_jac 1 0 0 0 $N 3
run TILE=75776 python tools/prof_one.py quad_jac 1 2500 128 2 $N 3
run TILE=151552 python tools/prof_one.py quad_jac 1 2500 128 2 $N 3
run CCU_JIT_STREAMS=2 TILE=75776 python tools/prof_one.py quad_jac 1 2500 128 2 $N 3
