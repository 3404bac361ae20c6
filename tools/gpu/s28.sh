timeout 600 python -m pytest tests/test_gpu_circuit.py tests/test_gpu_zz_late_additions.py tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -4
