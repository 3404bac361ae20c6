#!/usr/bin/env python
"""Extracts the expected instruction texts of the reference's per-gate `test_open_qasm` / `test_c_qasm` unit tests
(src/gates/*.rs) into tests/golden/qasm_kats.json:
    let bit_names = [String::from("qb0"), ...]; let qasm = <gate>.open_qasm(&bit_names, &[bits]);
    assert_eq!(qasm, Ok(String::from("...")));
The conditional variants take a free-form condition string that the circuit API never produces; they are left out
(the circuit-level conditional forms are pinned by circuit.rs:2059-2075, :2137-2151 in tests/test_export_cpu.py).
Reads /root/reference only here, at generation time.  usage: python tests/golden/make_qasm_kats.py [/root/reference]"""
import glob
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_gate_kats import ENV0, Gate, statements  # noqa: E402
from make_latex_kats import desc, gate_of  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.environ.get("Q1T_GOLDEN_OUT_DIR") or os.path.dirname(os.path.abspath(__file__)), "qasm_kats.json")


def main():
    cases, skipped = [], []
    for path in sorted(glob.glob(os.path.join(REF, "src", "gates", "*.rs"))):
        src = open(path).read()
        rel = os.path.relpath(path, REF)
        for m in re.finditer(r"fn (test_open_qasm|test_c_qasm)\(\)\s*\{", src):
            line0 = src.count("\n", 0, m.start()) + 1
            end = src.index("\n    }\n", m.end())
            body = src[m.end():end]
            strings = {}

            def stash(mm):
                key = "__s%d" % len(strings)
                # raw strings verbatim; ordinary literals with their escapes resolved
                strings[key] = mm.group(1) if mm.group(1) is not None else mm.group(2).replace("\\n", "\n").replace('\\"', '"').replace("\\\\", "\\")
                return key
            body = re.sub(r'r#"(.*?)"#|"((?:[^"\\]|\\.)*)"', stash, body, flags=re.S)
            body = re.sub(r"//[^\n]*", "", body)
            env = dict(ENV0)
            names, results = None, {}
            where = "%s:%d" % (rel, line0)
            try:
                for st in statements(body):
                    lm = re.match(r"let\s+(?:mut\s+)?(\w+)\s*=\s*(.*)$", st, re.S)
                    bm = lm and re.match(r"Composite::new\((\w+),\s*\d+\)$", lm.group(2).strip())
                    if bm:                       # builder style: Composite::new(name, n) + add_gate calls (composite.rs:1620-1665)
                        env[lm.group(1)] = Gate("Composite", strings[bm.group(1)], "")
                        continue
                    gm = re.match(r"(\w+)\.add_gate\((.*),\s*&\[([\d,\s]*)\]\)$", st, re.S)
                    if gm and isinstance(env.get(gm.group(1)), Gate):
                        comp, sub = env[gm.group(1)], gate_of(gm.group(2), env, strings)
                        text = sub.name + ("(" + ",".join(repr(float(a)) for a in sub.args) + ")" if sub.args else "")
                        text += " " + " ".join(re.findall(r"\d+", gm.group(3)))
                        comp.args = (comp.args[0], (comp.args[1] + "; " if comp.args[1] else "") + text)
                        continue
                    if lm and lm.group(1) == "bit_names":
                        names = [strings[k] for k in re.findall(r"__s\d+", lm.group(2))]
                        continue
                    qm = lm and re.match(r"(.*)\.(open_qasm|c_qasm)\(&bit_names,\s*&\[([\d,\s]*)\]\)$", lm.group(2), re.S)
                    if qm:
                        g = env[qm.group(1)] if qm.group(1) in env and isinstance(env[qm.group(1)], Gate) else gate_of(qm.group(1), env, strings)
                        results[lm.group(1)] = (g, qm.group(2), [int(v) for v in re.findall(r"\d+", qm.group(3))], list(names))
                        continue
                    if lm and re.match(r"String::from\((\w+)\)$", lm.group(2).strip()):
                        env[lm.group(1)] = strings[re.match(r"String::from\((\w+)\)$", lm.group(2).strip()).group(1)]
                        continue
                    if lm:
                        env[lm.group(1)] = gate_of(lm.group(2), env, strings)
                        continue
                    am = re.match(r"assert_eq!\((\w+),\s*Ok\((?:String::from\(\s*(\w+)\s*\)|(\w+)|String::new\(\))\)\)$", st, re.S)
                    if am and am.group(1) in results:
                        g, kind, bits, nm = results[am.group(1)]
                        key = am.group(2) or am.group(3)
                        text = "" if key is None else strings[key] if key in strings else env[key]
                        cases.append({"source": where, "kind": kind, "gate": desc(g), "bit_names": nm, "bits": bits, "text": text})
                        continue
                    raise ValueError("unhandled statement: " + st[:70].replace("\n", " "))
            except Exception as ex:          # noqa: BLE001
                skipped.append("%s: %s" % (where, ex))
    json.dump({"generator": "tests/golden/make_qasm_kats.py", "reference": "Q1tBV/q1tsim src/gates/*.rs test_open_qasm / test_c_qasm",
               "cases": cases}, open(OUT, "w"), indent=0)
    print("%d cases written to %s" % (len(cases), OUT))
    for s in skipped:
        print("skipped", s)


if __name__ == "__main__":
    main()
