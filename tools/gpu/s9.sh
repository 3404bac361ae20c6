mkdir -p gpurun_out
timeout 300 python tools/cfg2_probe.py 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2_n1_c.json 2>/dev/null; cut -c1-330 gpurun_out/bench_r2_n1_c.json
