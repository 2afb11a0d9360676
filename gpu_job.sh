cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nproc; free -g | head -2
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo bench rc=$?
tail -3 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2>&1; echo ncu1 rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 6 -c 2 -o gpurun_out/prof_raster python bench.py --steps 1 --warmup 3 --frames 16 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo ncu2 rc=$?
ls -la gpurun_out
