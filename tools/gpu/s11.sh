mkdir -p gpurun_out
Q1T_GROUP_SINGLE_BUFFER=1 Q1T_INPLACE_RELABEL=1 timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "one_gpu" 2>&1 | tail -6
timeout 600 python -m pytest tests/test_gpu_circuit.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2_n1_d.json 2>/dev/null; cut -c1-330 gpurun_out/bench_r2_n1_d.json
Q1T_HOST_PROFILE=1 timeout 120 python - <<'PY' 2>&1 | tail -4
import sys; sys.path.insert(0, '.')
from q1tsim_b200 import circuit as QC, engine as E, workloads as W
c = QC.Circuit(30, 30); W.load_ops(c, W.qft_ops(30)); r = E.Rng(seed=1)
for _ in range(3): c.execute(8192, r)
PY
