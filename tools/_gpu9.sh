set -x
mkdir -p gpurun_out
timeout 400 python tools/sweep_roll.py > gpurun_out/g9_roll.jsonl 2> gpurun_out/g9_roll.err
cut -c1-330 gpurun_out/g9_roll.jsonl; tail -3 gpurun_out/g9_roll.err
timeout 300 python -m pytest tests/test_reroll.py -x -q -m gpu > gpurun_out/g9_pytest.txt 2>&1
tail -5 gpurun_out/g9_pytest.txt
for W in 16 32; do
  CCU_JIT_REMAT=$W SWEEP_PLANS="1800:128:2" timeout 200 python tools/sweep_plans.py quad_adj kkt_qr >> gpurun_out/g9_sweep.jsonl 2>> gpurun_out/g9_sweep.err
done
SWEEP_PLANS="0:0:-1" timeout 200 python tools/sweep_plans.py quad_adj kkt_qr kkt_ldl rocket_hess >> gpurun_out/g9_sweep.jsonl 2>> gpurun_out/g9_sweep.err
cut -c1-230 gpurun_out/g9_sweep.jsonl
