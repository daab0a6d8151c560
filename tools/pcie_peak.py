"""Dev helper: pinned host<->device copy bandwidth of the box (the ceiling of bench.py's e2e number)."""
import json
import torch
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n // 8, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n // 8, dtype=torch.uint8, device="cuda")
s2 = torch.cuda.Stream()
res = {}
def best(fn, reps=5):
    b = 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        b = max(b, n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return b
res["h2d_GBps"] = best(lambda: d.copy_(h, non_blocking=True))
res["d2h_GBps"] = best(lambda: h.copy_(d, non_blocking=True))
def both():
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
    d.copy_(h, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)
res["h2d_GBps_with_concurrent_d2h_of_one_eighth"] = best(both)
print(json.dumps(res))
