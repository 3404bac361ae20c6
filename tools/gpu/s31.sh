Q1T_MID_RELABEL=2 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_circuit.py -m gpu -x -q 2>&1 | tail -3
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_n1_j.json 2> gpurun_out/bench_r2_n1_j.err; python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench_r2_n1_j.json') if l.startswith('{')][-1])
print("step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "roof", d["roofline"]["frac"], d["roofline"].get("avg_launch_ms"), d["verified"]["ok"], "dense step", d["dense_input_step"]["ms_per_step"])
PY
