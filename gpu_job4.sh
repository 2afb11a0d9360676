cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
for line in open('gpurun_out/bench_n2.json'):
    if line.startswith('{'):
        d=json.loads(line); print('N2', d['value'], d['n_gpus'], d['ms_per_step'], d['e2e'], d['kernel_ms_per_step'])
PY
timeout 600 python -m pytest tests/test_sharding.py -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-600
