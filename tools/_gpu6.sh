set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/g6_env.txt; nproc >> gpurun_out/g6_env.txt; free -g >> gpurun_out/g6_env.txt
timeout 900 python tools/sweep_plans.py > gpurun_out/g6_sweep.jsonl 2> gpurun_out/g6_sweep.err
tail -2 gpurun_out/g6_sweep.err
timeout 1500 python -m pytest tests -x -q -m gpu --durations=15 > gpurun_out/g6_pytest.txt 2>&1
tail -25 gpurun_out/g6_pytest.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/g6_bench.json 2> gpurun_out/g6_bench.err
tail -30 gpurun_out/g6_bench.err
