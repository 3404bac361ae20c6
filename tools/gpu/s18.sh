mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_r2_c.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/gputests_r2_c.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_n1_f.json 2> gpurun_out/bench_r2_n1_f.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench_r2_n1_f.json') if l.startswith('{')][-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["roofline"]["traffic"], d["verified"], d["clocks"], d["cpu_baseline"]["value"])
PY
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 --cpu-seconds 10 | cut -c1-400
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" | cut -c1-300
