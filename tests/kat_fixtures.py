"""Loader for tests/golden/gate_kats.json: the known-answer vectors of the reference's gate unit tests
(src/gates/*.rs `test_matrix*` / `test_apply*`), extracted by tests/golden/make_gate_kats.py."""
import json
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gate_kats.json")
# reference type name -> built-in table name (composite.rs:287-445)
_NAMES = {n: n.lower() for n in ("H", "I", "X", "Y", "Z", "S", "Sdg", "T", "Tdg", "V", "Vdg", "RX", "RY", "RZ", "U1", "U2", "U3", "CX", "CY",
                                 "CZ", "Swap", "CH", "CRX", "CRY", "CRZ", "CS", "CSdg", "CT", "CTdg", "CU1", "CU2", "CU3", "CV", "CVdg",
                                 "CCX", "CCZ", "CCRX", "CCRY", "CCRZ")}


def cases(kind=None):
    data = json.load(open(_PATH))["cases"]
    return [c for c in data if kind is None or c["kind"] == kind]


def carray(a):
    a = np.asarray(a, dtype=np.float64)
    return a[..., 0] + 1j * a[..., 1]


def matrix_of(desc, gate_matrix):
    """matrix() of a fixture's gate description through `gate_matrix(name, params)` (oracle or engine table):
    `C<G>` = I (+) G (controlled.rs:60-69), `Kron<G0, G1>` = G0 (x) G1 (kron.rs:59-62)"""
    name, args = desc["name"], desc["args"]
    if name == "C":
        g = matrix_of(args[0], gate_matrix)
        m = np.eye(2 * g.shape[0], dtype=np.complex128)
        m[g.shape[0]:, g.shape[0]:] = g
        return m
    if name == "Kron":
        return np.kron(matrix_of(args[0], gate_matrix), matrix_of(args[1], gate_matrix))
    return gate_matrix(_NAMES[name], args)


def case_id(c):
    return "%s-%s" % (c["source"].replace("src/gates/", "").replace(" ", ":"), c["gate"]["name"])
