set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 240 python -m pytest tests/test_dist.py -x -q -m gpu 2>&1 | tail -6
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --no-cpu --no-interp > gpurun_out/g13_bench_2gpu.json 2> gpurun_out/g13_bench_2gpu.err
tail -5 gpurun_out/g13_bench_2gpu.err
python - <<'P'
import json
for l in open('gpurun_out/g13_bench_2gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['n_gpus'], d['ms_per_step'])
        for c in d['configs']: print(c['config']['name'], c.get('value'), c.get('collective'), c.get('sums_hex'), c.get('error'))
P
