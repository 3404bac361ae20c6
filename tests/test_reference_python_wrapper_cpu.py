"""the reference's own Python front-end (python/q1tsim.py + python/q1tsimffi.py, cffi, UNMODIFIED)
runs against q1tsim_b200/lib/libq1tsim.so.  The wrapper dlopens './libq1tsim.so' (q1tsimffi.py:11-32), so it is
imported from inside the library directory in a child process.  Only the host side is exercised here (building,
error texts, export); execution needs a device and is covered by tests/test_gpu_circuit.py through the same ABI.
Skipped where /root/reference is not mounted (the GPU box)."""
import os
import subprocess
import sys
import textwrap

import pytest

from q1tsim_b200 import engine as E

REF_PY = "/root/reference/python"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF_PY, "q1tsim.py")), reason="reference not mounted")


def _run(body):
    code = "import sys\nsys.path.insert(0, %r)\nimport q1tsim\n" % REF_PY + textwrap.dedent(body)
    r = subprocess.run([sys.executable, "-c", code], cwd=os.path.dirname(E.LIB_PATH), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_builder_and_exporters_through_the_reference_wrapper():
    out = _run("""
        with q1tsim.Circuit(3, 3) as c:
            assert c.nr_qbits() == 3 and c.nr_cbits() == 3
            c.h(2); c.add_gate('CS', [1, 2]); c.add_gate('CT', [0, 2]); c.h(1); c.add_gate('CS', [0, 1]); c.h(0)
            c.swap(0, 2)
            c.rx(1.5, 0); c.u3(1.0, 2.25, 3.5, 1); c.cx(0, 1)
            c.add_conditional_gate([0, 1, 2], 5, 'X', [2])
            c.measure_all([0, 1, 2])
            print(c.open_qasm())
            print('----')
            print(c.c_qasm())
            print('----')
            print(c.latex())
    """)
    oq, cq, ltx = out.split("----\n")
    assert ltx.startswith("\\Qcircuit @C=1em @R=.7em {\n    \\lstick{\\ket{0}} & \\qw & \\qw & \\ctrl{2} & \\qw & \\ctrl{1} & \\gate{H} & \\qswap \\qwx[2]")
    assert oq == ('OPENQASM 2.0;\ninclude "qelib1.inc";\nqreg q[3];\ncreg b[3];\nh q[2];\ncu1(pi/2) q[1], q[2];\n'
                  "cu1(pi/4) q[0], q[2];\nh q[1];\ncu1(pi/2) q[0], q[1];\nh q[0];\n"
                  "cx q[0], q[2]; cx q[2], q[0]; cx q[0], q[2];\nrx(1.5) q[0];\nu3(1, 2.25, 3.5) q[1];\ncx q[0], q[1];\n"
                  "if (b == 5) x q[2];\nmeasure q -> b;\n\n")
    assert cq.startswith("version 1.0\nqubits 3\nh q[2]\ncrk q[1], q[2], 1\ncrk q[0], q[2], 2\n")
    assert "not b[1]\nc-x b[0], b[1], b[2], q[2]\nnot b[1]\nmeasure_all\n" in cq


def test_error_texts_through_the_reference_wrapper():
    out = _run("""
        with q1tsim.Circuit(2, 2) as c:
            for call in (lambda: c.add_gate('NOPE', [0]), lambda: c.h(7), lambda: c.measure(0, 9), lambda: c.histogram(),
                         lambda: c.rx(1.0, 5)):
                try:
                    call()
                    print('no error')
                except Exception as e:
                    print(e)
    """)
    lines = out.strip().split("\n")
    assert lines[0] == 'Unknown gate "NOPE"'                        # error.rs ParseError::UnknownGate
    assert lines[1] == "Invalid index 7 for a quantum bit"           # error.rs:196-198
    assert lines[2] == "Invalid index 9 for a classical bit"
    assert lines[3] == "The circuit has not been executed yet"       # error.rs NotExecuted
    assert lines[4] == "Invalid index 5 for a quantum bit"
