#!/usr/bin/env python
"""Streaming-bandwidth ceilings of this GPU for the roofline discussion: pure write (fill), copy
(read+write, how MEASURED_PEAKS.json's hbm_gbs was taken) and pure read (sum)."""
import json
import torch

n = 1 << 30  # 1 Gi doubles = 8 GiB
a = torch.empty(n, dtype=torch.float64, device="cuda")
b = torch.empty(n // 2, dtype=torch.float64, device="cuda")


def t(fn, it=10):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(it):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


out = {}
ms = t(lambda: a.fill_(1.0))
out["write_only_fill_GBps"] = 8 * n / ms / 1e6
ms = t(lambda: torch.cuda.current_stream().synchronize() or a.zero_())
out["write_only_memset_GBps"] = 8 * n / ms / 1e6
ms = t(lambda: b.copy_(a[: n // 2]))
out["copy_GBps_read_plus_write"] = 2 * 8 * (n // 2) / ms / 1e6
ms = t(lambda: a.sum())
out["read_only_sum_GBps"] = 8 * n / ms / 1e6
print(json.dumps(out))
