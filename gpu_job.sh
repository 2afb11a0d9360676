cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 ) 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_p6.json 2> gpurun_out/bench_p6.err; tail -3 gpurun_out/bench_p6.err
python - <<'PY'
import json
for n in ('bench_p6',):
    try:
        d=json.load(open('gpurun_out/%s.json'%n))
        print(n, round(d['value']), round(d['us_per_frame'],2), 'e2e', round(d['e2e']['value']), {k: round(v,3) for k,v in d['kernel_ms_per_step'].items()}, 'ms/step', round(d['ms_per_step'],3), d['roofline']['frac'], d['frame_roofline']['frac'])
    except Exception as e: print(n, 'failed', e)
PY
