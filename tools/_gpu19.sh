set -x
mkdir -p gpurun_out
export CASADI_CUDA_LIB=$PWD/casadi_b200/lib/libcasadi_cuda.so
T0=$(date +%s)
timeout 60 python -m pytest tests/test_gpu_parity.py -x -q -k page_locked > gpurun_out/g19_pytest.txt 2>&1; tail -4 gpurun_out/g19_pytest.txt
timeout 70 tests/integration/_build/bin/cuda_bench quad_ms 2000000 2 1 registered > gpurun_out/g19_e2e_registered.json 2> gpurun_out/g19_e2e_registered.err
cut -c1-600 gpurun_out/g19_e2e_registered.json; tail -2 gpurun_out/g19_e2e_registered.err
echo "elapsed $(( $(date +%s) - T0 ))"
timeout 50 python tools/sweep_env.py 40 kkt "" "CCU_JIT_SEG=8000 CCU_JIT_THREADS=128 CCU_JIT_MINBLOCKS=1" "CCU_JIT_SEG=8000 CCU_JIT_THREADS=64 CCU_JIT_MINBLOCKS=2" "CCU_JIT_CHAIN=1" "CCU_JIT_SEG=8000 CCU_JIT_THREADS=64 CCU_JIT_MINBLOCKS=3" > gpurun_out/g19_sweep_kkt.jsonl 2> gpurun_out/g19_sweep_kkt.err
cut -c1-330 gpurun_out/g19_sweep_kkt.jsonl; tail -2 gpurun_out/g19_sweep_kkt.err
echo "elapsed $(( $(date +%s) - T0 ))"
timeout 30 python tools/sweep_env.py 22 rocket_hess "" "CCU_JIT_CHAIN=1" > gpurun_out/g19_sweep_hess.jsonl 2> gpurun_out/g19_sweep_hess.err
cut -c1-330 gpurun_out/g19_sweep_hess.jsonl; tail -2 gpurun_out/g19_sweep_hess.err
timeout 25 python tools/sweep_env.py 18 mc "" "CCU_JIT_RING=64" "CCU_JIT_THREADS=128 CCU_JIT_MINBLOCKS=4" > gpurun_out/g19_sweep_mc.jsonl 2> gpurun_out/g19_sweep_mc.err
cut -c1-330 gpurun_out/g19_sweep_mc.jsonl; tail -2 gpurun_out/g19_sweep_mc.err
echo "elapsed $(( $(date +%s) - T0 ))"
