cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 ) 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err
tail -5 gpurun_out/bench_v3.err
python - <<'PY'
import json
for n in ('bench_v3',):
    d=json.load(open('gpurun_out/%s.json'%n))
    print(n, d['value'], d['us_per_frame'], 'e2e', d['e2e']['value'], d['e2e_color_depth']['value'], d['kernel_ms_per_step'], d['roofline']['frac'], d['frame_roofline']['frac'], 'ms/step', d['ms_per_step'])
PY
