cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -s 27 -c 9 -o gpurun_out/prof_all -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_all.log 2>&1
tail -3 gpurun_out/ncu_all.log
