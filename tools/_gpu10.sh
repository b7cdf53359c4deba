set -x
mkdir -p gpurun_out
export CASADI_CUDA_LIB=$PWD/casadi_b200/lib/libcasadi_cuda.so
( time timeout 300 tests/integration/_build/bin/test_cuda_map ) > gpurun_out/g10_integration.txt 2>&1
tail -8 gpurun_out/g10_integration.txt
for S in 1 0; do
  CCU_HOST_STREAMING=$S timeout 200 tests/integration/_build/bin/cuda_bench quad_ms 2000000 3 1 pageable >> gpurun_out/g10_e2e.jsonl 2>> gpurun_out/g10_e2e.err
done
timeout 200 tests/integration/_build/bin/cuda_bench quad_ms 2000000 3 1 pinned >> gpurun_out/g10_e2e.jsonl 2>> gpurun_out/g10_e2e.err
cut -c1-330 gpurun_out/g10_e2e.jsonl
python -c "
from casadi_b200 import capi
print('host copy GB/s', [round(capi.selftest_host_copy(1<<30, 16, 0),1) for _ in range(3)])" 
CCU_HOST_STREAMING=0 python -c "
from casadi_b200 import capi
print('host copy (memcpy) GB/s', [round(capi.selftest_host_copy(1<<30, 16, 0),1) for _ in range(3)])" 
nproc
timeout 200 python tools/sweep_roll.py mc quad_fwd > gpurun_out/g10_roll.jsonl 2> gpurun_out/g10_roll.err
cut -c1-330 gpurun_out/g10_roll.jsonl; tail -3 gpurun_out/g10_roll.err
