"""The oracle against every known-answer vector of the reference's gate unit tests (gate files'
test_matrix / test_apply / test_apply_mat, controlled.rs:574-675, kron.rs:202-288): 59 cases extracted from
src/gates/*.rs into tests/golden/gate_kats.json by tests/golden/make_gate_kats.py.  Tolerance 1e-15 absolute per
element, as `assert_complex_matrix_eq!` (cmatrix.rs:87-123).  Matrices come from the oracle's gate table, the apply
cases run through the oracle's gate application in both modes (faithful = the reference's loop structure)."""
import numpy as np
import pytest

from oracle import oracle as O
from q1tsim_b200 import engine as E
from tests import kat_fixtures as K

TOL = 1e-15


def test_fixture_inventory():
    cs = K.cases()
    assert len(cs) >= 59
    assert len({c["gate"]["name"] for c in cs}) >= 26          # every gate type of src/gates/ but Composite and Loop
    assert {c["kind"] for c in cs} == {"matrix", "apply"}


@pytest.mark.parametrize("case", K.cases("matrix"), ids=K.case_id)
def test_matrix_kats(case):
    want = K.carray(case["result"])
    assert np.abs(K.matrix_of(case["gate"], O.gate_matrix) - want).max() <= TOL
    # the product's host-side gate table (csrc/gates.cpp) is a separate restatement: same vectors
    assert np.abs(K.matrix_of(case["gate"], E.gate_matrix) - want).max() <= TOL


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("case", K.cases("apply"), ids=K.case_id)
def test_apply_kats(case, mode):
    """gate_test (gates.rs:373-381): the gate on the first k qubits of every column of `state`"""
    state, want = K.carray(case["state"]), K.carray(case["result"])
    m = K.matrix_of(case["gate"], O.gate_matrix)
    k, n = int(np.log2(m.shape[0])), int(np.log2(state.shape[0]))
    for col in range(state.shape[1]):
        o = O.OracleState(n, 1, mode=mode, order=1)
        o.set_column(0, state[:, col])
        o.apply_gate(m, list(range(k)))
        assert np.abs(o.column(0) - want[:, col]).max() <= TOL, (case["source"], col)
