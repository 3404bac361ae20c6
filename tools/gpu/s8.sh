mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_circuit.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/gputests_s8.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/gputests_s8.log
timeout 300 python tools/bench_configs.py --no-oracle > gpurun_out/bench_configs_r2b.jsonl 2>&1; cut -c1-420 gpurun_out/bench_configs_r2b.jsonl
Q1T_GRAPHS=0 timeout 300 python tools/bench_configs.py --no-oracle 2>&1 | cut -c1-300 | head -3
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/sanitizer_memcheck_smoke.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck_smoke.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "every_builtin_gate or measure_peek or conditional or canonical_reductions or dense_user or support_tracking" > gpurun_out/sanitizer_memcheck_parity.log 2>&1; echo "memcheck2 rc=$?"; tail -5 gpurun_out/sanitizer_memcheck_parity.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/sanitizer_racecheck_smoke.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck_smoke.log
