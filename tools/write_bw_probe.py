"""Write-only and copy bandwidth of this GPU (torch fill_ / copy_ over 16 GiB, CUDA events, best of 5):
the denominator for sweeps that only write (generated / support-tracked input)."""
import json, torch
n = 1 << 31                      # 2^31 doubles = 16 GiB
a = torch.empty(n, dtype=torch.float64, device="cuda")
def best(f, reps=5):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)
t_fill = best(lambda: a.fill_(1.5))
t_zero = best(lambda: a.zero_())
b = torch.empty(n // 2, dtype=torch.float64, device="cuda")
t_copy = best(lambda: b.copy_(a[: n // 2]))
print(json.dumps({"fill_16GiB_ms": t_fill, "fill_GBps": 8 * n / t_fill / 1e6, "zero_16GiB_ms": t_zero, "zero_GBps": 8 * n / t_zero / 1e6,
                  "copy_8GiB_ms": t_copy, "copy_GBps_read_plus_write": 2 * 8 * (n // 2) / t_copy / 1e6}))
