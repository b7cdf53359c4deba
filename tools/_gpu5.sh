set -x
mkdir -p gpurun_out
N=1048576
run() { echo "## $*" >> gpurun_out/g5_perf.txt; timeout 300 env "$@" >> gpurun_out/g5_perf.txt 2>&1; }
timeout 600 python tools/table.py > gpurun_out/g5_table.jsonl 2>&1
run CCU_CSE=0 python tools/prof_one.py quad_jac 1 0 0 0 $N 3
run CCU_SCHED_TIE=0 python tools/prof_one.py quad_jac 1 0 0 0 $N 3
run python tools/prof_one.py quad_jac 1 0 0 0 $N 3
run python tools/prof_one.py quad_jac 1 2500 128 2 $N 3
run python tools/prof_one.py quad_jac 1 1200 256 2 $N 3
run python tools/prof_one.py quad_jac 1 2500 128 3 $N 3
run python tools/prof_one.py quad_jac 1 4400 128 2 $N 3
run CCU_JIT_SEGWEIGHT=24000 python tools/prof_one.py quad_jac 1 4400 128 2 $N 3
run CCU_CSE=0 python tools/prof_one.py rocket_hess 1 0 0 0 $N 3
run python tools/prof_one.py rocket_hess 1 2500 128 2 $N 3
run python tools/prof_one.py quad_adj 1 2500 128 2 $N 3
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "every_specialisation or agree_bitwise or full_size or flagged or exact_class" > gpurun_out/g5_pytest.txt 2>&1
tail -3 gpurun_out/g5_pytest.txt
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/g5_bench.json 2> gpurun_out/g5_bench.err
tail -c 600 gpurun_out/g5_bench.err
