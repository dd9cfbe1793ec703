"""Write-only and read-only HBM bandwidth next to the read+write copy figure the rooflines use (MEASURED_PEAKS.json): CRBA is
write-dominated (11.5 GB written, 0.3 GB read per launch), so the write-only figure is its actual ceiling."""
import json

import torch

dev = torch.device("cuda:0")
n = 1 << 30  # 8 GiB of doubles
x = torch.empty(n, dtype=torch.float64, device=dev)
y = torch.empty(n, dtype=torch.float64, device=dev)
out = {}


def timed(fn, nbytes, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = max(best, nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


out["write_only_fill_gbs"] = timed(lambda: x.fill_(1.5), 8.0 * n)
out["write_only_memset_gbs"] = timed(lambda: x.zero_(), 8.0 * n)
out["read_only_sum_gbs"] = timed(lambda: x.sum(), 8.0 * n)
out["copy_read_plus_write_gbs"] = timed(lambda: y.copy_(x), 16.0 * n)
print(json.dumps(out))
