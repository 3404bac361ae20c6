mkdir -p gpurun_out
export Q1T_SWEEP_LOG=1
Q1T_HOT_TABLES=0 timeout 300 python tools/dense_ab.py 30 tma=0 2>&1 | grep -v "launch 0.0\|launch 2\." | tail -12
Q1T_HOT_TABLES=1 timeout 300 python tools/dense_ab.py 30 tma=0 tma=1 2>&1 | grep -v "launch 0.0\|launch 2\." | tail -24
unset Q1T_SWEEP_LOG
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
