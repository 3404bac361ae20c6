#!/usr/bin/env python
"""Benchmark of the statevector hot path (BASELINE.json: QFT-30 f64 + measure_all, 8192 shots on 1 B200; gate-amplitude
updates/s, HBM fraction of the sweep kernel, CPU port of the reference timed beside it; QFT-(30 + log2 N) and the
34-36-qubit configuration sharded over N = 2/4/8 B200).

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--qubits n]

One "step" = one full execution of the circuit: |0..0> state, every gate, measure_all of all shots.  `value` times
circuit.execute(shots) on a circuit that was built once (state buffers HBM-resident through the buffer cache); `e2e`
times what a user of the reference's FFI does from scratch every step (build the circuit from host data, execute,
read the classical register back to host memory).  Everything the line claims is checked after the timed loop
(`verified`): amplitudes against closed forms on a basis input and on a dense product-state input, the norm, and the
sampled outcomes.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "qft_f64_gate_amp_updates_per_s"
UNIT = "gate_amp_updates/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--qubits", type=int, default=30, help="qubits per GPU shard + log2(gpus) = circuit size")
    ap.add_argument("--shots", type=int, default=8192)
    ap.add_argument("--tile-bits", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--large-local-qubits", type=int, default=32,
                    help="N > 1: second leg with this many qubits per shard (32 = 64 GiB shards: QFT-33/34/35 on 2/4/8 GPUs); 0 = off")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md), polled through NVML every 5 ms
    from a thread of this process (nvidia-smi -lms cannot sample an 80 ms region)."""

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.stop_flag = False
        self.th = None
        self.err = None

    def _run(self):
        N, h, mx = self.N, self.h, self.mx
        try:
            while not self.stop_flag:
                sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
                rs = N.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(N, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                try:
                    pw = N.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:      # noqa: BLE001
                    pw = 0.0
                self.rows.append((sm, mx, rs, pw))
                self.first.set()
                time.sleep(0.005)
        except Exception as e:            # noqa: BLE001
            self.err = repr(e)
            self.first.set()

    def start(self):
        # NVML is initialised here, before the timed region, and start() returns once the first sample is in
        self.first = threading.Event()
        try:
            import pynvml as N
            N.nvmlInit()
            self.N = N
            self.h = N.nvmlDeviceGetHandleByIndex(self.idx)
            self.mx = N.nvmlDeviceGetMaxClockInfo(self.h, N.NVML_CLOCK_SM)
        except Exception as e:            # noqa: BLE001
            self.err = repr(e)
            return
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        self.first.wait(timeout=2.0)
        self.rows = []                    # (the sample taken before the region starts does not count)

    def stop(self):
        self.stop_flag = True
        if self.th is not None:
            self.th.join(timeout=2)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples" + (": " + self.err if self.err else "")]}
        # NVML clocks-event-reason bits
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80}
        reasons = sorted(nm for nm, b in bits.items() if any(r[2] & b for r in self.rows))
        return {"sm_mhz": float(np.median([r[0] for r in self.rows])), "sm_max_mhz": float(max(r[1] for r in self.rows)),
                "power_w_max": float(max(r[3] for r in self.rows)), "samples": len(self.rows), "reasons": reasons,
                "source": "NVML polled every 5 ms during the timed region"}


def measured_peak():
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        return float(json.load(open(pk))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md; of fallback)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per dense ladder launch, from the round's committed ncu capture"""
    for name in ("r2_ncu_dense_ladder_v3.json", "r2_ncu_dense_ladder_v2.json", "r2_ncu_dense_ladder.json"):       # (v3: the kernel and the plan as they are now, tools/ncu_summary.py)
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            d = json.load(open(p))
            return float(d["dram_bytes_per_launch"]), "profiles/%s (ncu --set full, %s)" % (name, d.get("command", ""))
    return None, None


# ---------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference's own algorithm (the reference is Rust and cannot be built in this
# image), single thread like the reference (vectorstate.rs has no parallelism).
# ---------------------------------------------------------------------------
def cpu_oracle_run(n, shots, faithful=True, threads=1):
    from oracle import oracle as O
    from q1tsim_b200 import workloads as W
    ops = W.qft_ops(n, measure=True)
    c = O.OracleCircuit(n, n, mode=0 if faithful else 1, order=0 if faithful else 1)
    W.load_ops(c, ops)
    O.lib().orc_set_threads(threads)
    t0 = time.perf_counter()
    c.execute(shots, O.Rng(seed=2))
    dt = time.perf_counter() - t0
    O.lib().orc_set_threads(1)
    return W.gate_count(ops) * float(1 << n) / dt, dt


def cpu_baseline(budget_s, shots):
    """bounded sample: the same circuit family (QFT-n + measure_all) at the largest n whose faithful single-thread run
    fits the budget; throughput is per amplitude, so the unit (gate-amplitude updates/s) carries over."""
    n = 14
    _, dt = cpu_oracle_run(n, min(shots, 1024))
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:      # noqa: BLE001
        avail = 32 << 30
    while n < 24:
        gates_ratio = ((n + 1) * (n + 2) / 2 + (n + 1) // 2) / (n * (n + 1) / 2 + n // 2)
        est = dt * 2.0 * gates_ratio
        if est > budget_s:
            break
        # the reference's measure_all collapse allocates a dense (2^n, n_distinct) matrix (vectorstate.rs:150-158),
        # and so does the faithful port: 16 B * 2^(n+1) * shots must stay well inside the host memory
        if (16 << (n + 1)) * min(shots, 1 << (n + 1)) > avail // 2:
            break
        n += 1
        _, dt = cpu_oracle_run(n, min(shots, 1024)) if est < 1.0 else (None, est)
    val, dt = cpu_oracle_run(n, shots)
    out = {"value": val, "unit": UNIT, "cores": 1, "kind": "port",
           "sample": "QFT-%d + measure_all, %d shots, oracle faithful mode (reference loop structure: per-block temporaries, "
                     "materialised bit_permutation gather/scatter, sequential prefix sum), 1 thread, %.1f s" % (n, shots, dt)}
    # beside it: the same arithmetic without the reference's temporaries, on every host core (OpenMP) -- what a tuned CPU
    # implementation of the path reaches on this box
    try:
        cores = len(os.sched_getaffinity(0))
        nf = min(n + 3, 26)
        while nf > n and (16 << nf) * min(shots, 1 << nf) > avail // 2:       # the dense collapse matrix again
            nf -= 1
        vf, dtf = cpu_oracle_run(nf, shots, faithful=False, threads=cores)
        out["fast_all_cores"] = {"value": vf, "unit": UNIT, "cores": cores,
                                 "sample": "QFT-%d + measure_all, oracle fast mode (strided loops, OpenMP), %.1f s" % (nf, dtf)}
    except Exception as e:     # noqa: BLE001
        out["fast_all_cores"] = {"error": repr(e)}
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    base = cpu_baseline(args.cpu_seconds, args.shots)
    times = []
    n = int(base["sample"].split("QFT-")[1].split(" ")[0])
    for i in range(args.warmup + args.steps):
        v, dt = cpu_oracle_run(n, args.shots)
        if i >= args.warmup:
            times.append(dt)
        if sum(times) > 150:
            break
    from q1tsim_b200 import workloads as W
    gates = W.gate_count(W.qft_ops(n))
    val = gates * float(1 << n) * len(times) / sum(times)
    base["value"] = val
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT,
            "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "QFT-%d f64 + measure_all, %d shots (bounded CPU sample of the QFT-%d workload; the reference's "
                                   "VectorState is single-threaded, so is this port)" % (n, args.shots, args.qubits)},
            "cpu_baseline": base,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# checks of what was timed
# ---------------------------------------------------------------------------
def rev_bits(v, n):
    r = 0
    for b in range(n):
        r |= ((v >> b) & 1) << (n - 1 - b)
    return r


def qft_of_basis_state(n, x, idx):
    """QFT|x> = 2^(-n/2) exp(2 pi i rev(x) rev(y) / 2^n) at the indices idx"""
    rx = rev_bits(x, n)
    ry = np.zeros_like(idx)
    for b in range(n):
        ry |= ((idx >> b) & 1) << (n - 1 - b)
    # rev(x) * rev(y) mod 2^n, exactly: multiply in two 31-bit halves of rev(x)
    N = 1 << n
    lo, hi = rx & ((1 << 20) - 1), rx >> 20
    prod = (ry * lo + (((ry * hi) % N) << 20)) % N if n <= 42 else None
    return np.exp(2j * np.pi * (prod.astype(np.float64) / float(N))) * 2.0 ** (-n / 2)


def window_offsets(nloc, seed, count=6, width=4096):
    rs = np.random.default_rng(seed)
    N = 1 << nloc
    return [0, N - width] + [int(v) & ~(width - 1) for v in rs.integers(0, N - width, size=count)]


def run_ours(args, rank, world, local):
    import torch
    from q1tsim_b200 import engine as E
    from q1tsim_b200 import workloads as W
    if not torch.cuda.is_available() or E.lib().q1t_device_count() < 1:
        raise RuntimeError("bench.py: no CUDA device; the engine has no CPU fallback")
    dev = local % E.lib().q1t_device_count()
    torch.cuda.set_device(dev)
    n, shots = args.qubits, args.shots
    ops = W.qft_ops(n, measure=True)
    gates = [(E.gate_matrix(o[1], o[2]), o[3], o[1]) for o in ops if o[0] == "gate"]
    ngates = len(gates)
    cbits = list(range(n))

    def sync():
        torch.cuda.synchronize()

    # ---- value: the reference-shaped call on a circuit built once: circuit.execute(nr_shots)
    # (circuit.rs:562-600: fresh |0..0> state, every gate, measure_all).  The op list lives in host memory and is
    # lowered, planned (plan cache) and launched by the C++ host layer; the state buffers are HBM-resident ----
    from q1tsim_b200 import circuit as QC
    circ = QC.Circuit(n, n, dev)
    W.load_ops(circ, ops)
    rng = E.Rng(seed=2)

    def step():
        circ.execute(shots, rng)

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(dev)
    # device-side timing: CUDA events bracket the K steps.  Every step ends synchronously (execute() returns the
    # sampled classical register), so the events see the whole region, host planning included
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    sampler.start()
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    sync()
    dt_wall = time.perf_counter() - t0
    dt = ev0.elapsed_time(ev1) * 1e-3
    clocks = sampler.stop()
    stats = circ.engine_stats()                # statistics of the last execute()
    # the sampled outcomes of the last step: QFT|0..0> is the uniform superposition -- every outcome < 2^n, and with
    # 8192 shots over 2^30 outcomes a repeated value has probability ~3 %: a stuck or collapsed sampler shows at once
    cs = np.asarray(circ.cstate(), dtype=np.uint64)
    outcomes_ok = bool(cs.size == shots and (cs >> np.uint64(n)).max() == 0 and np.unique(cs).size >= shots - 8)

    peak, peak_src = measured_peak()
    traffic, traffic_src = ncu_traffic()

    # ---- roofline of the dominant kernel: CUDA events on the engine's stream around every sweep launch ----
    # (a) DENSE input (seeded product state, SURVEY 8(d) cfg3 input B): every sweep reads and writes all 2^n
    #     amplitudes, 32 B each (the SURVEY 8(d) unit) -- the figure the kernel is judged by, and the one `roofline` leads with
    coefs = W.product_state_coefs(n, seed=1)
    sd = E.VectorState.from_qubit_coefs(coefs, shots, dev)
    if args.tile_bits:
        sd.set_option("tile_bits", args.tile_bits)
    res = np.zeros(shots, dtype=np.uint64)
    for m, b, name in gates:
        sd.apply_gate(m, b, name)
    sd.flush()
    dense_err = 0.0
    for off in window_offsets(n, 3):
        want = W.qft_of_product_state(n, coefs, np.arange(off, off + 4096, dtype=np.int64))
        dense_err = max(dense_err, float(np.linalg.norm(sd.column(0, off, 4096) - want) / np.linalg.norm(want)))
    dense_norm = float(sd.column_totals()[0])
    sd.set_timing(True)
    sd.reset_stats()
    dense_reps = 2
    for _ in range(dense_reps):
        for m, b, name in gates:
            sd.apply_gate(m, b, name)
        sd.flush()                      # (the state stays dense; repeated QFTs of it are as good as any dense input)
    td = sd.stats()
    sd.set_timing(False)
    # a full dense step (gates + measure_all of all shots), the N = 1 figure the sharded dense runs compare with
    sync()
    ev0.record()
    for m, b, name in gates:
        sd.apply_gate(m, b, name)
    sd.measure_all_into(cbits, res, rng)
    ev1.record()
    sync()
    dense_step_ms = ev0.elapsed_time(ev1)
    sd.close()

    def roof(t, reps):
        ms = t["sweep_ms"] / max(t["sweeps"], 1)
        by = t["sweep_bytes"] / max(t["sweeps"], 1)
        ach = by / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        return {"bound": "hbm", "kernel": "ladder_kernel", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "peak_source": peak_src, "bytes_per_launch": by, "avg_launch_ms": ms,
                "sweeps_per_step": t["sweeps"] / float(reps), "sweep_ms_per_step": t["sweep_ms"] / float(reps)}

    roofline = roof(td, dense_reps)
    roofline["traffic"] = traffic
    roofline["traffic_source"] = traffic_src
    roofline["input"] = ("dense seeded product state (from_qubit_coefs), the QFT-%d gate list: 32 B per amplitude and sweep "
                         "(SURVEY 8(d) unit), launches timed one by one with CUDA events on the engine's stream" % n)
    roofline["plan"] = ("%d sweeps, %d of them store relabelled in the middle of the plan (contiguous tiles, the next targets "
                        "rotated into the coalescing positions: DESIGN.md 4.5c)" % (round(td["sweeps"] / float(dense_reps)),
                                                                                     max(0, round(td.get("fused_relabels", 0) / float(dense_reps)) - 1)))

    # (b) the timed workload itself (input |0..0>): the engine tracks the support of the state, two of the three launches
    #     move almost nothing and the third writes the 2^n result: a write stream, measured against a write-only probe
    st = E.VectorState(n, shots, dev)
    if args.tile_bits:
        st.set_option("tile_bits", args.tile_bits)

    def qustate_step():
        st.reset_all()
        for m, b, name in gates:
            st.apply_gate(m, b, name)
        st.measure_all_into(cbits, res, rng)

    qustate_step()
    st.set_timing(True)
    st.reset_stats()
    for _ in range(2):
        qustate_step()
    ts = st.stats()
    st.set_timing(False)
    # verification on a basis input |x>: amplitudes against the closed form
    x = (0b1011 << (n - 5)) | 0b101
    st.reset_all()
    for q in range(n):
        if (x >> (n - 1 - q)) & 1:
            st.apply_gate(E.gate_matrix("x"), [q], "X")
    for m, b, name in gates:
        st.apply_gate(m, b, name)
    st.flush()
    basis_err = 0.0
    for off in window_offsets(n, 5, count=4):
        idx = np.arange(off, off + 4096, dtype=np.int64)
        want = qft_of_basis_state(n, x, idx)
        basis_err = max(basis_err, float(np.linalg.norm(st.column(0, off, 4096) - want) / np.linalg.norm(want)))
    basis_norm = float(st.column_totals()[0])
    st.close()
    # write-only probe, measured here: cudaMemset of 2^n amplitudes through torch (a pure write stream)
    probe = torch.empty(1 << (n + 1), dtype=torch.float64, device="cuda")
    probe.zero_()
    sync()
    ev0.record()
    for _ in range(3):
        probe.zero_()
    ev1.record()
    sync()
    write_peak = 3 * probe.numel() * 8 / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
    del probe
    tracked = roof(ts, 2)
    tracked["bound"] = "hbm_write"
    tracked["write_only_peak_gbs"] = write_peak
    tracked["write_only_peak_source"] = "torch zero_() of %d GiB, best-effort, measured in this run" % ((16 << n) >> 30)
    tracked["frac_of_write_only_peak"] = tracked["achieved"] / write_peak
    tracked["read_pass_ms_per_step"] = ts["read_ms"] / 2.0
    tracked["input"] = ("|0..0> (the timed workload): support tracking, bytes_per_launch = what the launches have to move "
                        "(two launches touch a few tiles, the third writes the 2^n result)")
    roofline["timed_workload"] = tracked

    # ---- e2e: the call a user makes, host buffers in, host buffers out ----
    e2e_steps = max(3, min(args.steps, 5))
    tot = {"h2d": 0, "d2h": 0}

    def e2e_step():
        # the reference-facing call: build the circuit through the ffi.rs-compatible C ABI,
        # execute(nr_shots) (fresh state, gate lowering + planning, every H2D/D2H copy), read c_state
        c = QC.Circuit(n, n, dev)
        W.load_ops(c, ops)
        c.execute(shots, rng)
        out = c.cstate()
        s_ = c.engine_stats()
        tot["h2d"] += s_["h2d_bytes"]
        tot["d2h"] += s_["d2h_bytes"]
        c.close()
        return out

    e2e_step()
    tot["h2d"] = tot["d2h"] = 0
    sync()
    ev0.record()
    for _ in range(e2e_steps):
        e2e_step()
    ev1.record()
    sync()
    dte = ev0.elapsed_time(ev1) * 1e-3

    updates_per_step = ngates * float(1 << n)
    value = updates_per_step * args.steps / dt
    e2e_val = updates_per_step * e2e_steps / dte
    verified = {"dense_input_rel_l2_vs_closed_form": dense_err, "dense_input_norm": dense_norm,
                "basis_input_rel_l2_vs_closed_form": basis_err, "basis_input_norm": basis_norm,
                "timed_outcomes_in_range_and_distinct": outcomes_ok,
                "ok": bool(dense_err < 1e-10 and basis_err < 1e-10 and abs(dense_norm - 1) < 1e-10 and abs(basis_norm - 1) < 1e-10 and outcomes_ok)}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "QFT-%d f64 + measure_all, %d shots, input |0..0>" % (n, shots), "gates": ngates,
                   "state_bytes": 16 << n, "l2": "state (16 GiB at n=30) is far larger than the 126 MB L2; no explicit flush",
                   "parallelism": "1 GPU"},
        "circuit_ms": 1e3 * dt / args.steps, "wall_ms_per_step": 1e3 * dt_wall / args.steps,
        "timing": "CUDA events around the K steps (every step ends synchronously), max over ranks",
        "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": 1e3 * dte / e2e_steps,
                "h2d_bytes_per_step": int(tot["h2d"] / e2e_steps), "d2h_bytes_per_step": int(tot["d2h"] / e2e_steps), "steps": e2e_steps,
                "bytes_source": "counted by the engine at every cudaMemcpyAsync (q1t_stats h2d_bytes / d2h_bytes)",
                "call": "Circuit built through the ffi.rs-compatible C ABI from host data, execute(shots), c_state read back to host, every step"},
        "gpu_launches": int(stats["kernel_launches"]) * args.steps,
        "engine_stats": stats,
        "roofline": roofline,
        "dense_input_step": {"ms_per_step": dense_step_ms, "value": updates_per_step / (dense_step_ms * 1e-3), "unit": UNIT,
                             "what": "the same gate list + measure_all on the dense product state: the N = 1 figure a dense sharded run compares with"},
        "verified": verified,
        "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.cpu_seconds, shots)
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# N > 1: one process per GPU, the state sharded by its top log2(N) qubits
# ---------------------------------------------------------------------------
def sharded_leg(args, rank, world, dev, n, shots, steps, warmup, dist, torch, dense):
    """QFT-n + measure_all over `world` ranks.  dense=False: from |0..0> through reset_all (the timed workload);
    dense=True: from a dense product state (every sweep and every exchange at full size)."""
    from q1tsim_b200 import engine as E
    from q1tsim_b200 import sharded as S
    from q1tsim_b200 import workloads as W
    g = int(round(math.log2(world)))
    ops = W.qft_ops(n, measure=True)
    gate_ops = [o for o in ops if o[0] == "gate"]
    ngates = len(gate_ops)
    rng = E.Rng(seed=2)
    res = np.zeros(shots, dtype=np.uint64)
    coefs = W.product_state_coefs(n, seed=1) if dense else None
    if dense:
        st = S.ShardedState.from_qubit_coefs(coefs, shots, device=dev)
    else:
        st = S.ShardedState(n, shots, device=dev)
    lcoefs = None
    if dense:
        import ctypes as C
        lcoefs = np.ascontiguousarray(np.asarray(coefs[2 * g:], dtype=np.complex128))
        scal = 1.0 + 0.0j
        for q in range(g):
            a, b = coefs[2 * q], coefs[2 * q + 1]
            scal *= (a, b)[(rank >> (g - 1 - q)) & 1] / math.sqrt(abs(a) ** 2 + abs(b) ** 2)

    def restart():
        if dense:
            st.where = [st._canonical(q) for q in range(n)]
            st.pin = [None] * g
            st.local.set_product_state(lcoefs)
            st.local.scale(scal)
            st._start_tag = "product"
        else:
            st.reset_all()

    evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    acc = {"ms": 0.0}

    def step():
        restart()
        if dense:
            # a dense input has to be built first (product_state_kernel): that is not part of the circuit, so the dense
            # legs time the circuit alone, step by step (every step ends synchronously)
            evs[0].record()
        st.run_ops(ops, E.gate_matrix, res, rng)        # gates (taped schedule) + canonical layout + measure_all
        if dense:
            evs[1].record()
            evs[1].synchronize()
            acc["ms"] += evs[0].elapsed_time(evs[1])

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 2)):
        step()
    st.exchanges = st.remaps = st.exchanged_bytes = 0
    st.exchange_seconds = 0.0
    st.local.reset_stats()
    acc["ms"] = 0.0
    sampler = ClockSampler(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    barrier()
    dt = ev0.elapsed_time(ev1) * 1e-3       # every step ends synchronously (sampled outcomes gathered on the host)
    if dense:
        dt = acc["ms"] * 1e-3
    clocks = sampler.stop()
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    stt = st.local.stats()
    remaps, exch_bytes = st.remaps / steps, st.exchanged_bytes / steps
    swap_ms = stt["peer_swap_ms"] / steps
    launches = stt["kernel_launches"]
    outcomes = res.copy()
    # ---- breakdown: one instrumented step (per-launch CUDA events serialise the stream, so it is not the timed one) ----
    st.local.set_timing(True)
    st.local.reset_stats()
    step()
    tb = st.local.stats()
    st.local.set_timing(False)
    sweep_ms, read_ms = tb["sweep_ms"], tb["read_ms"]
    # ---- verification ----
    ver = {}
    nl = n - g
    if dense:
        restart()
        st.run_ops(gate_ops, E.gate_matrix)
        st.canonicalize()
        err = 0.0
        for off in window_offsets(nl, 7 + rank, count=3):
            idx = (np.int64(rank) << np.int64(nl)) | np.arange(off, off + 4096, dtype=np.int64)
            want = W.qft_of_product_state(n, coefs, idx)
            err = max(err, float(np.linalg.norm(st.local.st.column(0, off, 4096) - want) / np.linalg.norm(want)))
        ver["dense_input_rel_l2_vs_closed_form"] = err
    else:
        x = (0b1011 << (n - 5)) | 0b101
        st.reset_all()
        xs = [("gate", "x", (), [q]) for q in range(n) if (x >> (n - 1 - q)) & 1]
        st.run_ops(xs + gate_ops, E.gate_matrix)
        st.canonicalize()
        err = 0.0
        for off in window_offsets(nl, 9 + rank, count=3):
            idx = (np.int64(rank) << np.int64(nl)) | np.arange(off, off + 4096, dtype=np.int64)
            want = qft_of_basis_state(n, x, idx)
            err = max(err, float(np.linalg.norm(st.local.st.column(0, off, 4096) - want) / np.linalg.norm(want)))
        ver["basis_input_rel_l2_vs_closed_form"] = err
    tot = float(st.column_totals()[0])
    ver["norm"] = tot
    ver["outcomes_in_range"] = bool((outcomes >> np.uint64(n)).max() == 0)
    if not dense:
        ver["outcomes_distinct"] = bool(np.unique(outcomes).size >= shots - 8)
    e = torch.tensor([err, abs(tot - 1.0)], device="cuda", dtype=torch.float64)
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    ver["max_over_ranks_rel_l2"] = float(e[0].item())
    ver["ok"] = bool(float(e[0].item()) < 1e-10 and float(e[1].item()) < 1e-10 and ver["outcomes_in_range"] and ver.get("outcomes_distinct", True))
    dist.barrier()
    st.local.group_close()
    dist.barrier()
    st.local.st.close()
    ms = 1e3 * dt / steps
    out = {"qubits": n, "gates": ngates, "value": ngates * float(1 << n) * steps / dt, "ms_per_step": ms, "steps": steps,
           "state_bytes": 16 << n, "shard_bytes": 16 << nl,
           "input": "dense product state (from_qubit_coefs)" if dense else "|0..0>",
           "breakdown_ms": {"sweeps": sweep_ms, "read_passes": read_ms, "swap_kernel": swap_ms,
                            "other": max(ms - sweep_ms - read_ms - swap_ms, 0.0),
                            "note": "sweeps / read passes from one instrumented step (per-launch events), swap from the timed steps"},
           "exchange": {"remaps_per_step": remaps, "bytes_out_per_rank_per_step": exch_bytes,
                        "swap_kernel_ms_per_step": swap_ms,
                        "gb_per_s_per_direction": (exch_bytes / (swap_ms * 1e-3) / 1e9) if swap_ms > 0 else None,
                        "fused_remaps_per_step": stt.get("fused_remaps", 0) / steps,
                        "path": ("remap read through by the ladder sweep that follows it (its tile loads come from the peers' shards over "
                                 "NVLink while the local part of every tile is swept; engine option fused_remap, default for 2 ranks)"
                                 if stt.get("fused_remaps", 0) else
                                 "in-place multi-bit remap kernel over CUDA IPC peer memory (NVLink)") +
                                "; mailbox barriers on the device; torch.distributed carries only host-side control messages"},
           "sweep_launches_per_step": tb["sweeps"], "gpu_launches": int(launches), "verified": ver, "clocks": clocks}
    return out


def run_sharded(args, rank, world, local):
    import torch
    import torch.distributed as dist
    from q1tsim_b200 import engine as E
    dev = local % max(E.lib().q1t_device_count(), 1)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    g = int(round(math.log2(world)))
    n, shots = args.qubits + g, args.shots
    peak, peak_src = measured_peak()
    main = sharded_leg(args, rank, world, dev, n, shots, args.steps, args.warmup, dist, torch, dense=False)
    dense = sharded_leg(args, rank, world, dev, n, shots, max(2, min(args.steps, 5)), 2, dist, torch, dense=True)
    large = None
    if args.large_local_qubits and args.large_local_qubits > args.qubits:
        try:
            large = {"from_zero": sharded_leg(args, rank, world, dev, args.large_local_qubits + g, shots, 3, 2, dist, torch, dense=False),
                     "dense": sharded_leg(args, rank, world, dev, args.large_local_qubits + g, shots, 2, 1, dist, torch, dense=True)}
        except Exception as e:       # noqa: BLE001
            large = {"error": repr(e)}
    if rank != 0:
        dist.destroy_process_group()
        return
    dsw = dense["breakdown_ms"]["sweeps"] / max(dense["sweep_launches_per_step"], 1)
    ach = (32 << (n - g)) / (dsw * 1e-3) / 1e9 if dsw > 0 else 0.0
    line = {
        "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "QFT-%d f64 + measure_all, %d shots, input |0..0>, sharded by the top %d qubits" % (n, shots, g),
                   "gates": main["gates"],
                   "state_bytes": 16 << n, "shard_bytes": 16 << (n - g),
                   "l2": "shards (16 GiB) are far larger than the 126 MB L2; no explicit flush",
                   "parallelism": "%d ranks (one process per GPU), state sharded by index bits; qubit remaps by an in-place multi-bit swap kernel "
                                  "over NVLink peer memory (CUDA IPC), device-side mailbox barriers; NCCL carries control messages only" % world},
        "circuit_ms": main["ms_per_step"],
        "timing": "CUDA events around the K steps (every step ends synchronously), max over ranks",
        "e2e": {"value": main["value"], "unit": UNIT, "ms_per_step": main["ms_per_step"],
                "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                "note": "the timed loop already goes through the public ShardedState API from host buffers (reset_all, the op list, "
                        "every H2D/D2H copy, remaps and sampling inside the timed region)"},
        "gpu_launches": main["gpu_launches"],
        "breakdown_ms": main["breakdown_ms"], "exchange": main["exchange"], "verified": main["verified"],
        "dense_input": dense,
        "north_star_large": large,
        "roofline": {"bound": "hbm", "kernel": "ladder_kernel", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": None, "bytes_per_launch": 32 << (n - g), "avg_launch_ms": dsw, "peak_source": peak_src,
                     "input": "dense product state, per shard: 32 B per amplitude and sweep"},
        "clocks": main["clocks"],
    }
    print(json.dumps(line))
    dist.destroy_process_group()


def main():
    args = parse_args()
    rank, world, local = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif world > 1:
        run_sharded(args, rank, world, local)
    else:
        run_ours(args, rank, world, local)


if __name__ == "__main__":
    main()
