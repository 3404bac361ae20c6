mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "fused_remap" 2>&1 | tail -25
