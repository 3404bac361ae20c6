mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense_user or error" 2>&1 | tail -4
timeout 120 python tools/generic_bw.py 28 2>&1 | tail -7
timeout 400 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_sharded_cabi.py -m gpu -x -q 2>&1 | tail -4
