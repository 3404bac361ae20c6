mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k vs_oracle > gpurun_out/gputests_sharded_n4.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/gputests_sharded_n4.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_r2_n4.json 2> gpurun_out/bench_r2_n4.err; echo "bench4 rc=$?"; python tools/gpu/show.py gpurun_out/bench_r2_n4.json; tail -4 gpurun_out/bench_r2_n4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; echo "bench2 rc=$?"; python tools/gpu/show.py gpurun_out/bench_r2_n2.json
