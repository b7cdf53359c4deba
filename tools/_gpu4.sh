set -x
mkdir -p gpurun_out
python -m pytest tests/test_dist.py tests/test_integration.py -x -q -m gpu > gpurun_out/g4_pytest.txt 2>&1
tail -15 gpurun_out/g4_pytest.txt
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pageable or chunked" >> gpurun_out/g4_pytest.txt 2>&1
tail -5 gpurun_out/g4_pytest.txt
export CASADI_CUDA_LIB=$PWD/casadi_b200/lib/libcasadi_cuda.so
B=tests/integration/_build/bin/cuda_bench
for m in pageable pinned; do
  echo "## cuda_bench quad_ms 4000000 $m" >> gpurun_out/g4_e2e.txt
  $B quad_ms 4000000 2 1 $m >> gpurun_out/g4_e2e.txt 2>&1
done
echo "## CCU_HOST_STAGING=0 pageable" >> gpurun_out/g4_e2e.txt
CCU_HOST_STAGING=0 $B quad_ms 4000000 2 1 pageable >> gpurun_out/g4_e2e.txt 2>&1
echo "## CCU_HOST_THREADS=16 pageable" >> gpurun_out/g4_e2e.txt
CCU_HOST_THREADS=16 $B quad_ms 4000000 2 1 pageable >> gpurun_out/g4_e2e.txt 2>&1
echo "## 2 devices pageable" >> gpurun_out/g4_e2e.txt
CASADI_CUDA_DEVICES=all $B quad_ms 4000000 2 1 pageable >> gpurun_out/g4_e2e.txt 2>&1
echo "## others" >> gpurun_out/g4_e2e.txt
$B cartpole 1000000 3 1 pageable >> gpurun_out/g4_e2e.txt 2>&1
$B mc 1000000 2 1 pageable reduce >> gpurun_out/g4_e2e.txt 2>&1
CASADI_CUDA_DEVICES=all $B mc 1000000 2 1 pageable reduce >> gpurun_out/g4_e2e.txt 2>&1
$B kkt_ldl 1000000 2 1 pageable >> gpurun_out/g4_e2e.txt 2>&1
$B rocket_hess 1000000 2 1 pageable >> gpurun_out/g4_e2e.txt 2>&1
nproc >> gpurun_out/g4_e2e.txt; free -g >> gpurun_out/g4_e2e.txt
