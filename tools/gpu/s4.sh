mkdir -p gpurun_out
Q1T_HOT_TABLES=0 timeout 300 python tools/dense_ab.py 30 tma=0 2>&1 | tail -1
Q1T_HOT_TABLES=1 timeout 300 python tools/dense_ab.py 30 tma=0 tma=1 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q 2>&1 | grep -v "^$" | tail -30
