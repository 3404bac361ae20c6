#!/usr/bin/env python
"""Extracts the known-answer vectors of the reference's own gate unit tests into tests/golden/gate_kats.json.

Reads /root/reference/src/gates/*.rs (only here, at generation time; the tests read the JSON) and collects, from
every `test_matrix*` / `test_apply*` function of the files' `mod tests`:
  * `assert_complex_matrix_eq!(<gate>.matrix(), array![...])`          -> kind "matrix"
  * `gate_test(<gate>, &mut state, &result)` (gates.rs:373-381: the gate's apply_slice on every column of
    `state`)                                                             -> kind "apply"
  * `<gate>.apply_mat(&mut state)` / `.apply(&mut state)` followed by `assert_complex_matrix_eq!(&state, &result)`
                                                                          -> kind "apply"
    (a state with more than 2^k rows: the gate acts on the first k qubits, gates.rs:273-326)
The Rust expressions of these tests are plain arithmetic over a few named constants, so they are evaluated by a
small textual translation to Python.  Statements that do not fit are reported and skipped, never guessed.

usage: python tests/golden/make_gate_kats.py [/root/reference] > report; writes gate_kats.json next to itself"""
import cmath
import glob
import json
import math
import os
import re
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.environ.get("Q1T_GOLDEN_OUT_DIR") or os.path.dirname(os.path.abspath(__file__)), "gate_kats.json")

CONSTS = {"PI": math.pi, "FRAC_PI_2": math.pi / 2, "FRAC_PI_3": math.pi / 3, "FRAC_PI_4": math.pi / 4, "FRAC_PI_6": math.pi / 6,
          "FRAC_PI_8": math.pi / 8, "FRAC_1_SQRT_2": 0.70710678118654752440, "SQRT_2": math.sqrt(2.0), "LN_2": math.log(2.0),
          "E": math.e, "FRAC_1_PI": 1 / math.pi, "FRAC_2_PI": 2 / math.pi, "LN_10": math.log(10.0)}


class Gate:
    def __init__(self, name, *args):
        self.name, self.args = name, args

    def desc(self):
        return {"name": self.name, "args": [a.desc() if isinstance(a, Gate) else float(a) for a in self.args]}


def mk(name, *args):
    return Gate(name, *args)


def A(rows):
    return np.array(rows, dtype=np.complex128)


def polar(r, t):
    return cmath.rect(r, t)


ENV0 = {"A": A, "mk": mk, "polar": polar, "complex": complex, "math": math, "np": np,
        "COMPLEX_ZERO": 0j, "COMPLEX_ONE": 1 + 0j, "COMPLEX_HSQRT2": complex(0.70710678118654752440, 0.0), "COMPLEX_I": 1j}
ENV0.update(CONSTS)


def match_bracket(s, i, open_ch, close_ch):
    depth = 0
    for j in range(i, len(s)):
        if s[j] == open_ch:
            depth += 1
        elif s[j] == close_ch:
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced")


def translate(expr):
    e = expr
    e = re.sub(r"(?:crate::)?cmatrix::", "", e)
    e = re.sub(r"(?:::)?std::f64::consts::", "", e)
    e = re.sub(r"(?:num_complex::)?Complex(?:64)?::new\(", "complex(", e)
    e = re.sub(r"(?:num_complex::)?Complex(?:64)?::from_polar\(", "polar(", e)
    e = re.sub(r"(?:crate::gates::)?(\w+)::new\(", r"mk('\1', ", e)
    e = e.replace("&", "")
    e = e.replace(".conj()", ".conjugate()")
    e = re.sub(r"(\d)_f64", r"\1", e)
    e = re.sub(r"(\d)f64", r"\1", e)
    # method calls on a name or a literal: x.sqrt() -> math.sqrt(x)
    for m in ("sqrt", "cos", "sin", "exp"):
        e = re.sub(r"([A-Za-z_][A-Za-z_0-9]*|\d+\.\d*)\.%s\(\)" % m, r"math.%s(\1)" % m, e)
    while "array![" in e:
        i = e.index("array![")
        j = match_bracket(e, i + 6, "[", "]")
        e = e[:i] + "A([" + e[i + 7:j] + "])" + e[j + 1:]
    if re.search(r"[A-Za-z_]\w*::|\.\w+\(|!|\bas\b", e.replace("math.", "").replace("np.", "").replace(".conjugate()", "")):
        raise ValueError("untranslatable: " + expr.strip()[:80])
    return e


def statements(body):
    """split a function body into `;`-terminated statements at bracket depth 0"""
    out, depth, cur = [], 0, ""
    for ch in body:
        cur += ch
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        elif ch == ";" and depth == 0:
            out.append(cur.strip().rstrip(";").strip())
            cur = ""
    return out


def arr(a):
    a = np.asarray(a, dtype=np.complex128)
    return [[[float(v.real), float(v.imag)] for v in row] for row in a]


def main():
    cases, skipped = [], []
    for path in sorted(glob.glob(os.path.join(REF, "src", "gates", "*.rs"))):
        src = open(path).read()
        rel = os.path.relpath(path, REF)
        for m in re.finditer(r"fn (test_matrix\w*|test_apply\w*)\(\)\s*\{", src):
            start = m.end() - 1
            end = match_bracket(src, start, "{", "}")
            body = re.sub(r"//[^\n]*", "", src[start + 1:end])
            line0 = src.count("\n", 0, m.start()) + 1
            env = dict(ENV0)
            where = "%s:%d %s" % (rel, line0, m.group(1))
            try:
                for st in statements(body):
                    lm = re.match(r"let\s+(?:mut\s+)?(\w+)\s*(?::[^=]+)?=\s*(.*)$", st, re.S)
                    if lm:
                        env[lm.group(1)] = eval(translate(lm.group(2)), env)
                        continue
                    am = re.match(r"assert_complex_matrix_eq!\((.*)\.matrix\(\)\s*,\s*(.*)\)$", st, re.S)
                    if am:
                        g = eval(translate(am.group(1)), env)
                        cases.append({"source": where, "kind": "matrix", "gate": g.desc(), "result": arr(eval(translate(am.group(2)), env))})
                        continue
                    gm = re.match(r"gate_test\((.*),\s*&mut\s+(\w+)\s*,\s*&(\w+)\)$", st, re.S)
                    if gm:
                        g = eval(translate(gm.group(1)), env)
                        cases.append({"source": where, "kind": "apply", "gate": g.desc(), "state": arr(env[gm.group(2)]),
                                      "result": arr(env[gm.group(3)])})
                        continue
                    pm = re.match(r"(.+)\.(apply_mat|apply)\(&mut\s+(\w+)\)$", st, re.S)
                    if pm:
                        g = eval(translate(pm.group(1)), env)
                        if isinstance(g, Gate):
                            env["__pending"] = (g, pm.group(3), np.array(env[pm.group(3)]))
                            continue
                    cm = re.match(r"assert_complex_(?:matrix|vector)_eq!\(&(\w+)\s*,\s*&(.+)\)$", st, re.S)
                    if cm and "__pending" in env and env["__pending"][1] == cm.group(1):
                        g, _, before = env.pop("__pending")
                        st_arr, res = np.asarray(before), np.asarray(eval(translate(cm.group(2)), env))
                        if st_arr.ndim == 1:
                            st_arr, res = st_arr.reshape(-1, 1), res.reshape(-1, 1)
                        cases.append({"source": where, "kind": "apply", "gate": g.desc(), "state": arr(st_arr), "result": arr(res)})
                        continue
                    raise ValueError("unhandled statement: " + st[:80].replace("\n", " "))
            except Exception as ex:          # noqa: BLE001 -- reported, never guessed
                skipped.append("%s: %s" % (where, ex))
    json.dump({"generator": "tests/golden/make_gate_kats.py", "reference": "Q1tBV/q1tsim src/gates/*.rs unit tests", "cases": cases},
              open(OUT, "w"), indent=0)
    print("%d cases written to %s" % (len(cases), OUT))
    for s in skipped:
        print("skipped", s)


if __name__ == "__main__":
    main()
