timeout 200 python tools/dense_ab.py 30 mid_relabel=0 mid_relabel=1 mid_relabel=0 mid_relabel=1 2>&1 | cut -c1-330
Q1T_MID_RELABEL=2 timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -4
