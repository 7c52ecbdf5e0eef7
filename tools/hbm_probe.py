"""HBM bandwidth of this GPU by access mix (CUDA events, best of 10): copy (read + write, the figure
MEASURED_PEAKS.json holds), write-only (fill), read-only (sum).  A pure vertex-write stream such as the
rigged-mesh replay is bounded by the WRITE-only figure.   Usage: python tools/hbm_probe.py [GiB]"""
import json
import sys

import torch

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
n = int(gib * (1 << 30) / 4)
dev = torch.device("cuda:0")
a = torch.empty(n, device=dev, dtype=torch.float32).normal_()
b = torch.empty_like(a)


def best(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t = min(t, e0.elapsed_time(e1))
    return t


nbytes = n * 4
out = {"bytes": nbytes,
       "copy_read_plus_write_gbs": 2 * nbytes / best(lambda: b.copy_(a)) / 1e6,
       "fill_write_only_gbs": nbytes / best(lambda: b.fill_(1.5)) / 1e6,
       "memset_write_only_gbs": nbytes / best(lambda: b.zero_()) / 1e6,
       "sum_read_only_gbs": nbytes / best(lambda: a.sum()) / 1e6}
print(json.dumps(out))
