cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
timeout 900 ncu --set full --clock-control none --import-source on -s 27 -c 9 -o gpurun_out/prof_all -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_all.log 2>&1
tail -2 gpurun_out/ncu_all.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 27 -c 27 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -3 gpurun_out/launches.csv | cut -c1-300
