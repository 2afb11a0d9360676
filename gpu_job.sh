cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -3
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 ) 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/split_frame.py --reps 5 > gpurun_out/split_n2.json 2> gpurun_out/split_n2.err; tail -3 gpurun_out/split_n2.err; cat gpurun_out/split_n2.json
timeout 600 python tools/split_frame.py --reps 5 > gpurun_out/split_n1.json 2> gpurun_out/split_n1.err; cat gpurun_out/split_n1.json
