cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2; do
HANA_BENCH_NOPROF=1 python bench.py --no-cpu-baseline > gpurun_out/np$i.json 2> gpurun_out/np$i.err; tail -2 gpurun_out/np$i.err
python bench.py --no-cpu-baseline > gpurun_out/p$i.json 2>gpurun_out/p$i.err
done
python - <<'PY'
import json
for n in ('np1','p1','np2','p2'):
    try:
        d=json.load(open('gpurun_out/%s.json'%n)); print(n, d['value'], d['ms_per_step'], sum(d['kernel_ms_per_step'].values()))
    except Exception as e: print(n, 'ERR', e)
PY
