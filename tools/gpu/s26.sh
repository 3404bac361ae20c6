mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_n1_h.json 2> gpurun_out/bench_r2_n1_h.err; python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench_r2_n1_h.json') if l.startswith('{')][-1])
print("step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "roof", d["roofline"]["frac"], d["roofline"].get("avg_launch_ms"), d["verified"]["ok"], d["clocks"].get("samples"), "cpu", d["cpu_baseline"]["value"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_v2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_v2.log 2>&1; tail -1 gpurun_out/ncu_bench_v2.log | cut -c1-200
