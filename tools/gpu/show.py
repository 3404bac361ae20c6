import json, sys
lines = [l for l in open(sys.argv[1]) if l.startswith("{")]
if not lines:
    print("no JSON line in", sys.argv[1]); sys.exit(0)
d = json.loads(lines[-1])
def leg(name, x):
    e = x["exchange"]
    print("%-14s n=%s %.2f ms/step  value %.3e  breakdown %s  remaps %.1f  GB/s/dir %s  verified %s" % (
        name, x.get("qubits", d["config"]["workload"][:7]), x["ms_per_step"], x["value"],
        {k: round(v, 2) for k, v in x["breakdown_ms"].items() if k != "note"}, e["remaps_per_step"],
        None if e["gb_per_s_per_direction"] is None else round(e["gb_per_s_per_direction"], 1), x["verified"]["ok"]))
leg("main", d)
leg("dense", d["dense_input"])
L = d.get("north_star_large")
if L:
    if "error" in L: print("large: ERROR", L["error"][:300])
    else:
        for k in ("from_zero", "dense"): leg("large " + k, L[k])
print("clocks", d["clocks"])
