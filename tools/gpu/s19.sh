mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_circuit.py tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2_n1_g.json 2>/dev/null; python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench_r2_n1_g.json') if l.startswith('{')][-1])
print("step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "roof", d["roofline"]["frac"], d["verified"]["ok"], d["clocks"]["samples"])
PY
for tb in 12 11 10; do Q1T_TILE_BITS=$tb timeout 200 python tools/bench_configs.py --no-oracle 2>&1 | grep "cfg2" | cut -c1-260; done
