mkdir -p gpurun_out
cat > /tmp/dbg.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
from oracle import oracle as O
from q1tsim_b200 import engine as E
n, P = 12, 2
st = E.ShardedProcessState(n, 8, [0] * P)
ref = O.OracleState(n, 8, mode=1, order=1)
def both(name, params, bits):
    m = O.gate_matrix(name, params); st.apply_gate(m, bits, name); ref.apply_gate(m, bits)
    print(name, bits, "queued", flush=True)
both("h", (), [0]); both("h", (), [3]); both("cx", (), [3, 5])
print("total", st.column_total(), st.counters(), flush=True)
both("h", (), [0])
print("total after remap gate", st.column_total(), st.counters(), flush=True)
a = st.amplitudes()
print("err", np.linalg.norm(a - ref.column(0)), flush=True)
PY
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 200 python /tmp/dbg.py 2>&1 | tail -12
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 compute-sanitizer --tool memcheck python /tmp/dbg.py 2>&1 | grep -v "^=========     at\|^=========         Device Frame\|Host Frame" | head -40
