mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/gputests_m2b.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/gputests_m2b.log
Q1T_SWEEP_LOG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_profile.py 30 > gpurun_out/prof_n2b.log 2>&1; echo "prof rc=$?"; grep "ms per step" gpurun_out/prof_n2b.log; grep "q1t launch" gpurun_out/prof_n2b.log | tail -12; grep -A22 "function calls" gpurun_out/prof_n2b.log | head -30
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2_n2_b.json 2> gpurun_out/bench_r2_n2_b.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_n2_b.json'))
print("main", d["ms_per_step"], d["breakdown_ms"], d["exchange"]["remaps_per_step"], d["verified"]["ok"])
dd=d["dense_input"]; print("dense", dd["ms_per_step"], dd["breakdown_ms"], dd["exchange"]["remaps_per_step"], dd["exchange"]["gb_per_s_per_direction"], dd["verified"]["ok"])
L=d["north_star_large"]
if "error" in L: print("large error", L)
else:
    for k in ("from_zero","dense"):
        x=L[k]; print("large", k, x["qubits"], x["ms_per_step"], x["breakdown_ms"], x["exchange"]["remaps_per_step"], x["exchange"]["gb_per_s_per_direction"], x["verified"]["ok"])
PY
tail -5 gpurun_out/bench_r2_n2_b.err
timeout 300 python tools/bench_configs.py --no-oracle > gpurun_out/bench_configs_r2a.jsonl 2>&1; cut -c1-420 gpurun_out/bench_configs_r2a.jsonl
