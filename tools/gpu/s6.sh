mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_n1_a.json 2> gpurun_out/bench_r2_n1_a.err; echo "bench rc=$?"; cut -c1-3000 gpurun_out/bench_r2_n1_a.json; tail -5 gpurun_out/bench_r2_n1_a.err
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_r2_a.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/gputests_r2_a.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ladder_kernel -s 3 -c 3 -f -o gpurun_out/r2_dense_ladder python tools/dense_ab.py 30 tma=0 > gpurun_out/ncu_dense.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_dense.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out
