mkdir -p gpurun_out
Q1T_TEST_DEVICES=0,1,2,3,4,5,6,7 timeout 600 python -m pytest tests/test_gpu_sharded_cabi.py -m gpu -x -q -k "8-16 or 8-17" > gpurun_out/gputests_cabi_8gpu.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/gputests_cabi_8gpu.log
timeout 300 python tools/cabi_sharded_bench.py 30 0,1,2,3,4,5,6,7 8 2>&1 | tail -1
timeout 300 python tools/cabi_sharded_bench.py 32 0,1,2,3,4,5,6,7 4 2>&1 | tail -1
timeout 400 python tools/cabi_sharded_bench.py 33 0,1,2,3,4,5,6,7 3 2>&1 | tail -1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r2_n8_b.json 2> gpurun_out/bench_r2_n8_b.err; echo "bench8 rc=$?"; python tools/gpu/show.py gpurun_out/bench_r2_n8_b.json; tail -3 gpurun_out/bench_r2_n8_b.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --qubits 33 --large-local-qubits 0 --steps 3 --warmup 1 > gpurun_out/bench_r2_n8_q36_b.json 2> gpurun_out/bench_r2_n8_q36_b.err; echo "bench8-q36 rc=$?"; python tools/gpu/show.py gpurun_out/bench_r2_n8_q36_b.json
