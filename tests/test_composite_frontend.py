"""Host-side text front-end: `Composite::from_string` and `Expression::parse`
restated in q1tsim_b200/csrc/composite.cpp, pinned to the reference's own unit tests
(composite.rs:719-1617, expression.rs:504-531).  No GPU needed: parsing, matrices and the flattening
of a composite into the circuit's op list are host code."""
import math

import numpy as np
import pytest

from q1tsim_b200 import circuit as QC
from q1tsim_b200 import engine as E

TOL = 1e-15            # assert_complex_matrix_eq!


def polar(th):
    return complex(math.cos(th), math.sin(th))


def test_from_string_composition_and_arguments():
    # composite.rs:727-757
    inc3 = np.zeros((8, 8))
    for r, c in [(0, 7)] + [(i + 1, i) for i in range(7)]:
        inc3[r, c] = 1
    assert np.allclose(E.composite_matrix("CCX 2 1 0; CX 2 1; X 2"), inc3, atol=TOL)
    y = E.composite_matrix("U3(3.141592653589793,1.570796326794897,1.570796326794897) 0")
    assert np.allclose(y, [[0, -1j], [1j, 0]], atol=1e-15)


@pytest.mark.parametrize("desc,name,params", [
    ("CCRX(3.141592653589793) 0 1 2", "ccrx", (math.pi,)), ("CCRY(3.141592653589793) 0 1 2", "ccry", (math.pi,)),
    ("CCRZ(3.141592653589793) 0 1 2", "ccrz", (math.pi,)), ("CCX 0 1 2", "ccx", ()), ("CCZ 0 1 2", "ccz", ()),
    ("CH 0 1", "ch", ()), ("CRX(1.570796326794897) 0 1", "crx", (1.570796326794897,)), ("CRY(0.3) 0 1", "cry", (0.3,)),
    ("CRZ(0.7) 0 1", "crz", (0.7,)), ("CS 0 1", "cs", ()), ("CSdg 0 1", "csdg", ()), ("CT 0 1", "ct", ()),
    ("CTdg 0 1", "ctdg", ()), ("CU1(0.9) 0 1", "cu1", (0.9,)), ("CU2(0.1,0.2) 0 1", "cu2", (0.1, 0.2)),
    ("CU3(0.1,0.2,0.3) 0 1", "cu3", (0.1, 0.2, 0.3)), ("CV 0 1", "cv", ()), ("CVdg 0 1", "cvdg", ()), ("CX 0 1", "cx", ()),
    ("CY 0 1", "cy", ()), ("CZ 0 1", "cz", ()), ("H 0", "h", ()), ("I 0", "i", ()), ("RX(0.4) 0", "rx", (0.4,)),
    ("RY(0.4) 0", "ry", (0.4,)), ("RZ(0.4) 0", "rz", (0.4,)), ("S 0", "s", ()), ("Sdg 0", "sdg", ()), ("T 0", "t", ()),
    ("Tdg 0", "tdg", ()), ("Swap 0 1", "swap", ()), ("U1(0.4) 0", "u1", (0.4,)), ("U2(0.4,0.5) 0", "u2", (0.4, 0.5)),
    ("U3(0.4,0.5,0.6) 0", "u3", (0.4, 0.5, 0.6)), ("V 0", "v", ()), ("Vdg 0", "vdg", ()), ("X 0", "x", ()), ("Y 0", "y", ()),
    ("Z 0", "z", ())])
def test_from_string_every_gate_of_the_table(desc, name, params):
    """composite.rs:287-445 / test_from_string_gates (:761-1338): every name, case-insensitive"""
    want = E.gate_matrix(name, params)
    assert np.allclose(E.composite_matrix(desc), want, atol=TOL)
    assert np.allclose(E.composite_matrix(desc.lower()), want, atol=TOL)


def test_from_string_known_matrices():
    # composite.rs:768-800 (first two closed forms)
    m = E.composite_matrix("CCRX(3.141592653589793) 0 1 2")
    want = np.eye(8, dtype=complex)
    want[6:, 6:] = [[0, -1j], [-1j, 0]]
    assert np.allclose(m, want, atol=1e-15)
    m = E.composite_matrix("CCRY(3.141592653589793) 0 1 2")
    want[6:, 6:] = [[0, -1], [1, 0]]
    assert np.allclose(m, want, atol=1e-15)


@pytest.mark.parametrize("arg,value", [
    ("0.23", 0.23), ("0.23+0.16", 0.39), ("1.23-0.16", 1.07), ("0.8*0.6", 0.48), ("1.38/2", 0.69), ("-0.23", -0.23),
    ("1.03^5", 1.1592740743), ("sin(1.0)", 0.8414709848078965), ("cos(1.0)", 0.5403023058681398),
    ("tan(1.0)", 1.5574077246549023), ("exp(1.0)", 2.718281828459045), ("ln(0.8)", -0.2231435513142097),
    ("sqrt(0.8)", 0.8944271909999159), ("0.5*(1.0-0.26)", 0.37), ("3", 3.0), ("pi", math.pi)])
def test_from_string_arguments(arg, value):
    """composite.rs:1341-1551"""
    m = E.composite_matrix("U1(%s) 0" % arg)
    assert np.allclose(m, [[1, 0], [0, polar(value)]], atol=1e-15)


def test_from_string_argument_list_and_precedence():
    # composite.rs:1553-1563: right-associative power, nested parentheses, unary minus
    m = E.composite_matrix("U3(pi/2, 1.03^2^(1.05-0.23), -1.78) 0")
    want = [[0.7071067811865476, complex(0.1468526445611853, 0.6916894540076393)],
            [complex(0.3496446456390944, 0.6146125786020914), complex(0.5285973886766332, -0.4696645618782032)]]
    assert np.allclose(m, want, atol=1e-15)


@pytest.mark.parametrize("desc,text", [
    ("XYZ 0", 'Unknown gate "XYZ"'),                                             # UnknownGate
    ("X 1; 0", 'Failed to find gate name in " 0"'),                              # NoGateName
    ("RX(1.2, 3.4) 1", 'Expected 1 arguments to "RX" gate, got 2'),              # InvalidNrArguments(2, 1, _)
    ("H 0 1", 'Expected 1 bits for "H" gate, got 2'),                            # InvalidNrBits(2, 1, _)
    ("RX(abc) 1", 'Failed to parse argument "abc) 1"'),                          # InvalidArgument
    ("U1(12897231928172918729136192817936) 0", "Failed to parse argument"),      # InvalidArgument (u64 overflow)
    ("H 0; X", "Unable to find the bits gate X operates on"),                    # NoBits
    ("H 117356715625188271521875", "Failed to parse bit number in"),             # InvalidBit
    ("H 0 and something", 'Trailing text after gate description: "and something"'),   # TrailingText
    ("RX(1.2a) 1", "Unclosed parentheses in expression"),                        # UnclosedParentheses
    ("RX(1.2*(1+2 1", "Unclosed parentheses in expression"),
    ("RX(sin(1.2 1", "Unclosed parentheses in expression")])
def test_from_string_errors(desc, text):
    """composite.rs:1569-1617, texts error.rs:93-127"""
    with pytest.raises(E.ParseError) as ei:
        E.composite_matrix(desc)
    assert text in str(ei.value)


@pytest.mark.parametrize("text,value,rest", [("1 + 2 * 3", 7.0, ""), ("1/2 - (1+4)", -4.5, ""), ("sin(1/2)", 0.479425538604203, ""),
                                             ("2^3^2", 512.0, ""), ("--2", 2.0, ""), ("1.5e2*2 , 3", 300.0, " , 3"),
                                             (".5+1.", 1.5, ""), ("-pi/4) 0", -math.pi / 4, ") 0")])
def test_expression_parse(text, value, rest):
    """expression.rs:504-531 plus the literal forms of :93-97"""
    v, r = E.eval_expression(text)
    assert abs(v - value) <= max(abs(v), abs(value)) * 2.3e-16
    assert r == rest


def test_composite_is_flattened_into_the_circuit():
    """sub-gates become circuit ops on the mapped qubits; Loop repeats the body; a failing
    sub-gate leaves the circuit untouched; bit count is checked like any gate (gates.rs:176-186)."""
    c = QC.Circuit(5, 5)
    c.add_composite_gate("Inc3", "CCX 2 1 0; CX 2 1; X 2", [4, 0, 2])
    assert c.nr_ops() == 3
    c.add_composite_gate("Loop", "H 0; CX 0 1", [1, 3], nr_iterations=4)
    assert c.nr_ops() == 3 + 8
    with pytest.raises(Exception) as ei:
        c.add_composite_gate("Inc3", "CCX 2 1 0; CX 2 1; X 2", [4, 0])
    assert 'Expected 3 bits for "Inc3", got 2' in str(ei.value)
    with pytest.raises(Exception) as ei:
        c.add_composite_gate("bad", "H 0; X 1", [0, 7])          # qubit 7 does not exist: nothing is added
    assert "Invalid index 7 for a quantum bit" in str(ei.value)
    assert c.nr_ops() == 11
    with pytest.raises(Exception) as ei:
        c.add_composite_gate("bad", "H 0; Q 1", [0, 1])
    assert 'Unknown gate "Q"' in str(ei.value)
    c.close()
