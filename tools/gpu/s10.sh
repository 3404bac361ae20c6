mkdir -p gpurun_out
timeout 300 python tools/cfg2_probe.py 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_circuit.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/bench_configs.py --no-oracle > gpurun_out/bench_configs_r2c.jsonl 2>&1; cut -c1-330 gpurun_out/bench_configs_r2c.jsonl
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k one_gpu 2>&1 | tail -4
