cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) 2>&1
( time timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err ) 2>&1 | tail -3
tail -5 gpurun_out/bench_v3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_v3.json'))
print(d['value'], d['us_per_frame'], 'e2e', d['e2e']['value'], d['e2e_color_depth']['value'], d['kernel_ms_per_step'], d['roofline']['frac'], d['frame_roofline']['frac'], 'ms/step', d['ms_per_step'])
PY
