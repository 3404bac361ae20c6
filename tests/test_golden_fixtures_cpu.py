"""The committed fixtures under tests/golden/ are exactly what their generator scripts extract from the reference's
unit tests: each script is re-run into a scratch directory and its output compared with the committed JSON.
Skipped where /root/reference is not mounted (the GPU box) -- the fixtures themselves travel."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/src/gates"), reason="reference not mounted")


@pytest.mark.parametrize("script,fixture,min_cases", [("make_gate_kats.py", "gate_kats.json", 59), ("make_latex_kats.py", "latex_kats.json", 41),
                                                       ("make_qasm_kats.py", "qasm_kats.json", 67)])
def test_fixture_is_current(tmp_path, script, fixture, min_cases):
    env = dict(os.environ, Q1T_GOLDEN_OUT_DIR=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(HERE, "golden", script), "/root/reference"], env=env, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    fresh = json.load(open(os.path.join(str(tmp_path), fixture)))
    committed = json.load(open(os.path.join(HERE, "golden", fixture)))
    assert len(fresh["cases"]) >= min_cases
    assert fresh == committed
