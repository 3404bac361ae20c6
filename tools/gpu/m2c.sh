mkdir -p gpurun_out
Q1T_TEST_DEVICES=0,1 timeout 900 python -m pytest tests/test_gpu_sharded_cabi.py -m gpu -x -q > gpurun_out/gputests_cabi_2gpu.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/gputests_cabi_2gpu.log
timeout 300 python tools/cabi_sharded_bench.py 30 0,1 8 2>&1 | tail -2
timeout 300 python tools/cabi_sharded_bench.py 32 0,1 4 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 20 --warmup 5 --large-local-qubits 0 > gpurun_out/bench_r2_n2_c.json 2> gpurun_out/bench_r2_n2_c.err; echo "bench2 rc=$?"; python tools/gpu/show.py gpurun_out/bench_r2_n2_c.json
