mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
L=q1tsim_b200/lib
echo "== new default (unified body, byte addresses, dense instantiation)"; timeout 200 python tools/dense_ab.py 30 2>&1 | tail -1
echo "== new lib, general instantiation"; Q1T_LADDER_DENSE=0 timeout 200 python tools/dense_ab.py 30 2>&1 | tail -1
echo "== per-length bodies, byte addresses, dense instantiation"; Q1T_LIB=$L/libq1tsim_u0b1.so timeout 200 python tools/dense_ab.py 30 2>&1 | tail -1
echo "== per-length bodies, element addresses, dense"; Q1T_LIB=$L/libq1tsim_u0b0.so timeout 200 python tools/dense_ab.py 30 2>&1 | tail -1
echo "== per-length bodies, element addresses, general (= previous kernel)"; Q1T_LADDER_DENSE=0 Q1T_LIB=$L/libq1tsim_u0b0.so timeout 200 python tools/dense_ab.py 30 2>&1 | tail -1
