set -x
mkdir -p gpurun_out
export CASADI_CUDA_LIB=$PWD/casadi_b200/lib/libcasadi_cuda.so
( time timeout 600 python -m pytest tests -x -q -m gpu --durations=6 ) > gpurun_out/g17_pytest.txt 2>&1
tail -14 gpurun_out/g17_pytest.txt
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --no-extra > gpurun_out/g17_bench.json 2> gpurun_out/g17_bench.err
tail -6 gpurun_out/g17_bench.err
timeout 100 tests/integration/_build/bin/cuda_bench mc 2000000 2 1 pageable reduce > gpurun_out/g17_mc_e2e.json 2> gpurun_out/g17_mc_e2e.err
cut -c1-400 gpurun_out/g17_mc_e2e.json; tail -2 gpurun_out/g17_mc_e2e.err
timeout 200 python tools/table.py > gpurun_out/g17_table.jsonl 2> gpurun_out/g17_table.err
cut -c1-260 gpurun_out/g17_table.jsonl; tail -3 gpurun_out/g17_table.err
