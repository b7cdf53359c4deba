set -x
mkdir -p gpurun_out
export CASADI_CUDA_LIB=$PWD/casadi_b200/lib/libcasadi_cuda.so
T0=$(date +%s)
( time timeout 170 python -m pytest tests/test_integration.py -x -q -m gpu ) > gpurun_out/g20_integration.txt 2>&1; tail -12 gpurun_out/g20_integration.txt
echo "elapsed $(( $(date +%s) - T0 ))"
( time timeout 50 python -m pytest tests/test_gpu_parity.py -x -q -k "host_path or null_argument or reduce or page_locked" ) > gpurun_out/g20_hostpath.txt 2>&1; tail -5 gpurun_out/g20_hostpath.txt
echo "elapsed $(( $(date +%s) - T0 ))"
