mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k one_gpu 2>&1 | tail -40
