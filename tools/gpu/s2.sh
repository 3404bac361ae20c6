mkdir -p gpurun_out
export Q1T_SWEEP_LOG=1
timeout 300 python tools/dense_ab.py 30 tma=0 tma=1 > gpurun_out/s2_persweep.log 2>&1
grep -v "^q1t launch 0.0" gpurun_out/s2_persweep.log | tail -40
unset Q1T_SWEEP_LOG
Q1T_LIB=$PWD/q1tsim_b200/lib/libq1tsim_pclk.so timeout 300 python tools/dense_ab.py 30 tma=0 > gpurun_out/s2_pclk_tma0.log 2>&1
Q1T_LIB=$PWD/q1tsim_b200/lib/libq1tsim_pclk.so timeout 300 python tools/dense_ab.py 30 tma=1 > gpurun_out/s2_pclk_tma1.log 2>&1
tail -14 gpurun_out/s2_pclk_tma0.log; tail -14 gpurun_out/s2_pclk_tma1.log
Q1T_LIB=$PWD/q1tsim_b200/lib/libq1tsim_r4c3.so timeout 300 python tools/dense_ab.py 30 tma=0 > gpurun_out/s2_r4c3.log 2>&1; tail -2 gpurun_out/s2_r4c3.log
Q1T_LIB=$PWD/q1tsim_b200/lib/libq1tsim_r4c2.so timeout 300 python tools/dense_ab.py 30 tma=0 > gpurun_out/s2_r4c2.log 2>&1; tail -2 gpurun_out/s2_r4c2.log
