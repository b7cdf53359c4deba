mkdir -p gpurun_out
timeout 68 python tools/sweep_env.py 58 mc,rocket_hess,kkt "" "CCU_JIT_IOBASE=1" > gpurun_out/g21_sweep_iobase.jsonl 2> gpurun_out/g21_sweep_iobase.err
cut -c1-260 gpurun_out/g21_sweep_iobase.jsonl; tail -2 gpurun_out/g21_sweep_iobase.err
