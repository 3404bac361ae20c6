mkdir -p gpurun_out
L=q1tsim_b200/lib
echo "== default (per-length bodies, byte addresses, dense instantiation)"; timeout 200 python tools/dense_ab.py 30 2>&1 | tail -1
echo "== static rounds"; Q1T_LIB=$L/libq1tsim_sr.so timeout 200 python tools/dense_ab.py 30 2>&1 | tail -1
echo "== tile bits 11"; Q1T_TILE_BITS=11 timeout 200 python tools/dense_ab.py 30 2>&1 | tail -1
echo "== tile bits 11, static rounds"; Q1T_TILE_BITS=11 Q1T_LIB=$L/libq1tsim_sr.so timeout 200 python tools/dense_ab.py 30 2>&1 | tail -1
echo "== tile bits 10"; Q1T_TILE_BITS=10 timeout 200 python tools/dense_ab.py 30 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ladder_kernel -s 3 -c 3 -o gpurun_out/r2_dense_ladder_v2 -f python tools/dense_ab.py 30 > gpurun_out/ncu_dense_v2.log 2>&1; tail -2 gpurun_out/ncu_dense_v2.log
