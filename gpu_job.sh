cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in base g8 o8; do
  cp variants/lib_$v.so hana-softwarerenderer_b200/libhana_b200.so
  if [ $v != o8 ]; then ( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2 ) 2>&1; fi
  timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; tail -3 gpurun_out/bench_$v.err
done
python - <<'PY'
import json
for n in ('bench_base','bench_g8','bench_o8'):
    try:
        d=json.load(open('gpurun_out/%s.json'%n))
        print(n, round(d['value']), round(d['us_per_frame'],2), 'e2e', round(d['e2e']['value']), {k: round(v,3) for k,v in d['kernel_ms_per_step'].items()}, 'ms/step', round(d['ms_per_step'],3))
    except Exception as e: print(n, 'failed', e)
PY
