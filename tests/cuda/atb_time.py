"""Per-launch time of gtb_rows_atb_f32 on 1M rows x 64 x 64 (the E-sized weight-gradient launches of a training step)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gnn_tracking_b200 import ops
n = 1_000_000
a, b = torch.randn(n, 64, device="cuda"), torch.randn(n, 64, device="cuda")
idx = torch.randperm(n, device="cuda").to(torch.int32)
out, cs = torch.zeros(64, 64, device="cuda"), torch.zeros(64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for label, kw in (("plain", {}), ("relu+colsum", dict(a_relu=True, colsum=cs)), ("gather", dict(a_index=idx))):
    ts = []
    for i in range(8):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.rows_atb(a, b, out, **kw); e.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(s.elapsed_time(e))
    t = sum(ts) / len(ts)
    print(f"rows_atb {label}: {t * 1e3:.1f} us per launch, {n * 512 / t / 1e6:.0f} GB/s")
for label, (ka, nb) in (("64x1", (64, 1)), ("4x64", (4, 64))):
    aa, bb = torch.randn(n, ka, device="cuda"), torch.randn(n, nb, device="cuda")
    oo = torch.zeros(ka, nb, device="cuda")
    ts = []
    for i in range(8):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.rows_atb(aa, bb, oo); e.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(s.elapsed_time(e))
    print(f"rows_atb {label}: {sum(ts) / len(ts) * 1e3:.1f} us per launch")
