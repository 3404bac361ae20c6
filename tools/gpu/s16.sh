mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded_cabi.py -m gpu -x -q -k "basis_changes or errors" 2>&1 | tail -30
