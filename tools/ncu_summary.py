#!/usr/bin/env python
"""Summary of an `ncu --set full` capture of the dense ladder sweeps (read here, on the CPU box):
    python tools/ncu_summary.py gpurun_out/r2_dense_ladder_v2.ncu-rep profiles/r2_ncu_dense_ladder_v2.json "<command that was profiled>"
bench.py reads `dram_bytes_per_launch` from the JSON for `roofline.traffic`."""
import csv, io, json, subprocess, sys

rep, out, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, k, scale_unit=None):
    v = float(r[col[k]].replace(",", ""))
    u = units[col[k]]
    if scale_unit == "GB":
        v *= {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}[u]
    if scale_unit == "ms":
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}[u]
    return v


stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
launches = []
for r in rows[2:]:
    if "ladder_kernel" not in r[col["Kernel Name"]]:
        continue
    launches.append({
        "kernel": r[col["Kernel Name"]].split("(")[0],
        "duration_ms": val(r, "gpu__time_duration.sum", "ms"),
        "dram_read_gb": val(r, "dram__bytes_read.sum", "GB"), "dram_write_gb": val(r, "dram__bytes_write.sum", "GB"),
        "instructions": val(r, "smsp__inst_executed.sum"),
        "fp64_pipe_pct": val(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "cycles_per_instruction_and_warp": val(r, "smsp__average_warp_latency_per_inst_issued.ratio"),
        "registers": val(r, "launch__registers_per_thread"),
        "local_memory_requests": val(r, "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum") + val(r, "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum"),
        "stalls_per_issue": {h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]: round(val(r, h), 3) for h in stalls},
    })
n = max(len(launches), 1)
json.dump({"command": cmd, "what": "the three sweeps of the QFT-30 gate list on a dense product state (2^30 amplitudes, 34.36 GB algorithmic per launch), B200",
           "dram_bytes_per_launch": 1e9 * sum(l["dram_read_gb"] + l["dram_write_gb"] for l in launches) / n,
           "algorithmic_bytes_per_launch": 32 << 30, "launches": launches}, open(out, "w"), indent=1)
print(out, "launches", len(launches))
