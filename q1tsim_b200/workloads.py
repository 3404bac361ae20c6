"""Synthetic workloads of BASELINE.json, as backend-neutral op lists.

An op list is a sequence of tuples that `load_ops` feeds to any object exposing
the reference's `Circuit` builder methods (q1tsim_b200.circuit.Circuit does):
  ("gate", name, params, bits) | ("measure_all", cbits, basis) | ("measure", q, c, basis)
  | ("cond", control, target, name, params, bits) | ...
Generators follow SURVEY.md 8(d); the PRNG is SplitMix64 so C++/Python agree.
"""
import math

MASK64 = (1 << 64) - 1
BENCH_SEED = 0x1F67A51423CD2615      # benches/benchmarks/manybits.rs:24


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & MASK64

    def next_u64(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)

    def f64(self):
        return (self.next_u64() >> 11) * (1.0 / (1 << 53))

    def words(self, n):
        return [self.next_u64() for _ in range(n)]


def qft_ops(n, measure=True, swaps=True):
    """n-qubit generalisation of README.md:59-68 (cfg1/cfg3/cfg5): exact CS/CT
    for distance 1/2, CU1(pi/2^d) beyond."""
    ops = []
    for j in reversed(range(n)):
        ops.append(("gate", "h", (), [j]))
        for i in reversed(range(j)):
            d = j - i
            if d == 1:
                ops.append(("gate", "cs", (), [i, j]))
            elif d == 2:
                ops.append(("gate", "ct", (), [i, j]))
            else:
                ops.append(("gate", "cu1", (math.pi / (1 << d),), [i, j]))
    if swaps:
        for i in range(n // 2):
            ops.append(("gate", "swap", (), [i, n - 1 - i]))
    if measure:
        ops.append(("measure_all", list(range(n)), "Z"))
    return ops


def u3_layer_ops(n, seed=1):
    """Seeded product-state preparation (cfg3 input B)."""
    r = SplitMix64(seed)
    return [("gate", "u3", (2 * math.pi * r.f64(), 2 * math.pi * r.f64(), 2 * math.pi * r.f64()), [q]) for q in range(n)]


def random_circuit_ops(n=20, depth=100, seed=BENCH_SEED, measure=True):
    """cfg2: even layers H or U3 per qubit, odd layers CX/CS/CT on shuffled pairs."""
    r = SplitMix64(seed)
    ops = []
    for layer in range(depth):
        if layer % 2 == 0:
            for q in range(n):
                if r.next_u64() >> 63:
                    ops.append(("gate", "h", (), [q]))
                else:
                    ops.append(("gate", "u3", (2 * math.pi * r.f64(), 2 * math.pi * r.f64(), 2 * math.pi * r.f64()), [q]))
        else:
            qs = list(range(n))
            for k in range(n - 1, 0, -1):           # Fisher-Yates
                j = r.next_u64() % (k + 1)
                qs[k], qs[j] = qs[j], qs[k]
            for a, b in zip(qs[0::2], qs[1::2]):
                ops.append(("gate", ("cx", "cs", "ct")[r.next_u64() % 3], (), [a, b]))
    if measure:
        ops.append(("measure_all", list(range(n)), "Z"))
    return ops


def ghz_branching_ops(n=24):
    """cfg4: GHZ + X/Y/Z-basis mid-circuit measurements + conditional gates."""
    ops = [("gate", "h", (), [0])]
    ops += [("gate", "cx", (), [i, i + 1]) for i in range(n - 1)]
    ops += [("measure", 0, 0, "X"), ("measure", 1, 1, "Y"), ("measure", 2, 2, "Z")]
    ops.append(("cond", [0, 1], 0b01, "z", (), [n - 1]))
    ops.append(("cond", [2], 1, "x", (), [n // 2]))
    ops.append(("measure_all", list(range(n)), "Z"))
    return ops


def gate_count(ops):
    return sum(1 for o in ops if o[0] in ("gate", "cond"))


def load_ops(circuit, ops):
    """Feed an op list to any object with the reference's builder methods."""
    for op in ops:
        k = op[0]
        if k == "gate":
            circuit.add_gate(op[1], op[3], op[2])
        elif k == "cond":
            circuit.add_conditional_gate(op[1], op[2], op[3], op[5], op[4])
        elif k == "measure":
            circuit.measure_basis(op[1], op[2], op[3])
        elif k == "peek":
            circuit.peek_basis(op[1], op[2], op[3])
        elif k == "measure_all":
            circuit.measure_all_basis(op[1], op[2])
        elif k == "peek_all":
            circuit.peek_all_basis(op[1], op[2])
        elif k == "reset":
            circuit.reset(op[1])
        elif k == "reset_all":
            circuit.reset_all()
        elif k == "barrier":
            circuit.barrier(op[1])
        else:
            raise ValueError(k)
    return circuit


def product_state_coefs(n, seed=1):
    """Seeded per-qubit coefficient pairs (a_q, b_q) for `VectorState::from_qubit_coefs` (vectorstate.rs:62-83): a dense
    product state with no zero amplitude (cfg3 input B)."""
    r = SplitMix64(seed)
    coefs = []
    for _ in range(n):
        th, ph = math.pi * (0.15 + 0.7 * r.f64()), 2 * math.pi * r.f64()
        coefs += [complex(math.cos(th / 2), 0.0), complex(math.cos(ph) * math.sin(th / 2), math.sin(ph) * math.sin(th / 2))]
    return coefs


def qft_of_product_state(n, coefs, idx):
    """Closed form of qft_ops(n, swaps=True) applied to the product state of `coefs` (normalised per qubit, qubit 0 =
    most significant index bit), evaluated at the basis indices `idx` (numpy int64 array).

    The circuit maps |x> to 2^(-n/2) sum_y exp(2 pi i rev(x) rev(y) / 2^n) |y> (rev = n-bit reversal); with
    psi_x = prod_q c_q[x_q] and rev(x) = sum_q x_q 2^q the sum over x factorises:
        out[y] = 2^(-n/2) prod_q (c_q[0] + c_q[1] exp(2 pi i (2^q rev(y) mod 2^n) / 2^n)).
    A size-independent check of the dense path: O(n) per amplitude at any n."""
    import numpy as np
    idx = np.asarray(idx, dtype=np.int64)
    rev = np.zeros_like(idx)
    for b in range(n):
        rev |= ((idx >> b) & 1) << (n - 1 - b)
    out = np.full(idx.shape, 2.0 ** (-n / 2), dtype=np.complex128)
    for q in range(n):
        a, b = complex(coefs[2 * q]), complex(coefs[2 * q + 1])
        nrm = math.sqrt(abs(a) ** 2 + abs(b) ** 2)
        frac = ((rev & ((1 << (n - q)) - 1)) << q).astype(np.float64) / float(1 << n)      # (2^q rev mod 2^n) / 2^n, exact
        out *= (a + b * np.exp(2j * np.pi * frac)) / nrm
    return out
