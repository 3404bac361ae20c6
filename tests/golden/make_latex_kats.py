#!/usr/bin/env python
"""Extracts the expected Qcircuit texts of the reference's per-gate `test_latex` unit tests (src/gates/*.rs) into
tests/golden/latex_kats.json.  Each case there is
    let gate = <gate>; let mut state = LatexExportState::new(nq, nc); gate.latex(&[bits], &mut state); state.code() == r#"..."#
which equals `Circuit(nq, nc)` + that one gate + `latex()` (circuit.rs:1148-1160).  Cases that switch composite
expansion off (`set_expand_composite(false)`, not reachable through the circuit API) are left out.
Reads /root/reference only here, at generation time.  usage: python tests/golden/make_latex_kats.py [/root/reference]"""
import glob
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_gate_kats import ENV0, Gate, match_bracket, statements, translate  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.environ.get("Q1T_GOLDEN_OUT_DIR") or os.path.dirname(os.path.abspath(__file__)), "latex_kats.json")


def gate_of(expr, env, strings):
    expr = expr.strip()
    m = re.match(r"Composite::from_string\((\w+),\s*(\w+)\)\.unwrap\(\)$", expr)
    if m:
        return Gate("Composite", strings[m.group(1)], strings[m.group(2)])
    m = re.match(r"Loop::new\((\w+),\s*(\d+),\s*(\w+)\)$", expr)
    if m:
        return Gate("Loop", strings[m.group(1)], int(m.group(2)), env[m.group(3)])
    return eval(translate(expr), env)


def desc(g):
    return {"name": g.name, "args": [desc(a) if isinstance(a, Gate) else a for a in g.args]}


def main():
    cases, skipped = [], []
    for path in sorted(glob.glob(os.path.join(REF, "src", "gates", "*.rs"))):
        src = open(path).read()
        rel = os.path.relpath(path, REF)
        for m in re.finditer(r"fn (test_latex)\(\)\s*\{", src):
            line0 = src.count("\n", 0, m.start()) + 1
            # the body ends at the first line that is exactly four spaces and a closing brace (raw strings hold unbalanced braces)
            end = src.index("\n    }\n", m.end())
            body = src[m.end():end]
            strings = {}

            def stash(mm):
                key = "__s%d" % len(strings)
                # raw strings verbatim; ordinary literals with their escapes resolved
                strings[key] = mm.group(1) if mm.group(1) is not None else mm.group(2).replace("\\n", "\n").replace('\\"', '"').replace("\\\\", "\\")
                return key
            body = re.sub(r'r#"(.*?)"#|"((?:[^"\\]|\\.)*)"', stash, body, flags=re.S)
            env = dict(ENV0)
            gate, nq, nc, bits, expand = None, None, None, None, True
            where = "%s:%d" % (rel, line0)
            try:
                for st in statements(body):
                    lm = re.match(r"let\s+(?:mut\s+)?(\w+)\s*=\s*(.*)$", st, re.S)
                    if lm and lm.group(2).startswith("LatexExportState::new"):
                        nq, nc = [int(v) for v in re.findall(r"\d+", lm.group(2))]
                        expand = True
                        continue
                    if lm:
                        env[lm.group(1)] = gate_of(lm.group(2), env, strings)
                        continue
                    if re.match(r"state\.set_expand_composite\(false\)$", st):
                        expand = False
                        continue
                    am = re.match(r"assert_eq!\((\w+)\.latex\(&\[([\d,\s]*)\],\s*&mut state\),\s*Ok\(\(\)\)\)$", st, re.S)
                    if am:
                        gate = env[am.group(1)]
                        bits = [int(v) for v in re.findall(r"\d+", am.group(2))]
                        continue
                    cm = re.match(r"assert_eq!\(state\.code\(\),\s*(\w+)\)$", st, re.S)
                    if cm:
                        if expand:
                            cases.append({"source": where, "gate": desc(gate), "nr_qbits": nq, "nr_cbits": nc, "bits": bits,
                                          "latex": strings[cm.group(1)]})
                        continue
                    raise ValueError("unhandled statement: " + st[:70].replace("\n", " "))
            except Exception as ex:          # noqa: BLE001
                skipped.append("%s: %s" % (where, ex))
    json.dump({"generator": "tests/golden/make_latex_kats.py", "reference": "Q1tBV/q1tsim src/gates/*.rs test_latex", "cases": cases},
              open(OUT, "w"), indent=0)
    print("%d cases written to %s" % (len(cases), OUT))
    for s in skipped:
        print("skipped", s)


if __name__ == "__main__":
    main()
