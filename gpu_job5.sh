cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python tools/time_configs.py 2>&1 | tail -4 | tee gpurun_out/configs_time.jsonl
