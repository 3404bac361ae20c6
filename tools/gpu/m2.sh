mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_profile.py 30 > gpurun_out/prof_n2.log 2>&1; echo "prof rc=$?"; grep -v "^W1\|^\[W\|Warning" gpurun_out/prof_n2.log | head -60
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2_n2_a.json 2> gpurun_out/bench_r2_n2_a.err; echo "bench rc=$?"; cut -c1-6000 gpurun_out/bench_r2_n2_a.json; tail -15 gpurun_out/bench_r2_n2_a.err
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k vs_oracle > gpurun_out/gputests_sharded_n2.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/gputests_sharded_n2.log
