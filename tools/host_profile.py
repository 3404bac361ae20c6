"""Host-side phases of circuit.execute() (Q1T_HOST_PROFILE=1, csrc/circuit.cpp): QFT-n + measure_all.
usage: Q1T_HOST_PROFILE=1 python tools/host_profile.py [n] [shots] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from q1tsim_b200 import circuit as QC  # noqa: E402
from q1tsim_b200 import engine as E  # noqa: E402
from q1tsim_b200 import workloads as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
shots = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
c = QC.Circuit(n, n)
W.load_ops(c, W.qft_ops(n, measure=True))
rng = E.Rng(seed=2)
for r in range(reps):
    t0 = time.perf_counter()
    c.execute(shots, rng)
    print("execute %d: %.3f ms" % (r, 1e3 * (time.perf_counter() - t0)), file=sys.stderr)
