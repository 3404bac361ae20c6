mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "vs_oracle and 8-21" > gpurun_out/gputests_sharded_n8.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/gputests_sharded_n8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r2_n8.json 2> gpurun_out/bench_r2_n8.err; echo "bench8 rc=$?"; python tools/gpu/show.py gpurun_out/bench_r2_n8.json; tail -4 gpurun_out/bench_r2_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --qubits 33 --large-local-qubits 0 --steps 3 --warmup 1 > gpurun_out/bench_r2_n8_q36.json 2> gpurun_out/bench_r2_n8_q36.err; echo "bench8-q36 rc=$?"; python tools/gpu/show.py gpurun_out/bench_r2_n8_q36.json; tail -6 gpurun_out/bench_r2_n8_q36.err
