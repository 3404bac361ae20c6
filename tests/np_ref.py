"""Independent numpy restatement (dense Kronecker algebra, small n only) used to
cross-check the C oracle.  Semantics: qubit 0 is the most significant index bit
(vectorstate.rs:62-74, gates.rs:100-101); the gate-matrix index has bits[0] as
its most significant bit (gates.rs:53-80 closed form)."""
import numpy as np


def apply_gate(state, mat, bits, n):
    """state: (2^n,) complex; returns new state."""
    k = len(bits)
    psi = state.reshape([2] * n)
    m = np.asarray(mat, dtype=np.complex128).reshape([2] * (2 * k))
    # contract gate input axes (k..2k-1) with state axes bits[0..k-1]
    out = np.tensordot(m, psi, axes=(list(range(k, 2 * k)), list(bits)))
    # result axes: gate outputs first (in bits order), then the remaining axes ascending
    rest = [a for a in range(n) if a not in bits]
    order = list(bits) + rest
    inv = np.argsort(order)
    return np.transpose(out, inv).reshape(-1)


def marginal0(state, qbit, n):
    psi = np.abs(state.reshape([2] * n)) ** 2
    return float(np.take(psi, 0, axis=qbit).sum())


def qft_closed_form(n, x):
    """amp[y] = 2^(-n/2) exp(2 pi i rev(x) rev(y) / 2^n)."""
    def rev(v):
        return int(format(v, "0%db" % n)[::-1], 2)
    N = 1 << n
    ys = np.array([rev(y) for y in range(N)], dtype=np.float64)
    return np.exp(2j * np.pi * ((rev(x) * ys) % N) / N) / np.sqrt(N)
