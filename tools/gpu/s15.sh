mkdir -p gpurun_out
Q1T_SWEEP_LOG=1 timeout 300 python tools/dense_ab.py 30 tma=0 2>&1 | grep -v "launch 0.0\|launch 2\." | tail -7
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_circuit.py -m gpu -x -q 2>&1 | tail -3
