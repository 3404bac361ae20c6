mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_r2_b.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/gputests_r2_b.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_n1_b.json 2> gpurun_out/bench_r2_n1_b.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_r2_n1_b.json; tail -3 gpurun_out/bench_r2_n1_b.err
