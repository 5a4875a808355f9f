"""Per-phase timing of gtb_dbscan_f32 (run on the GPU box)."""
import ctypes as C
import torch
from gnn_tracking_b200 import ops
from gnn_tracking_b200._lib import lib
from gnn_tracking_b200.postprocessing.dbscan import dbscan

g = torch.Generator(device="cuda").manual_seed(0)
n = 100_000
for name, x, eps, mp in [
    ("tight clusters d=8", (torch.rand((n // 10, 8), device="cuda", generator=g) * 20)[torch.randint(0, n // 10, (n,), device="cuda", generator=g)] + 0.02 * torch.randn((n, 8), device="cuda", generator=g), 0.2, 1),
    ("gauss*1.5 d=8", torch.randn((n, 8), device="cuda", generator=g) * 1.5, 0.8, 3),
    ("gauss*3 d=2", torch.randn((n, 2), device="cuda", generator=g) * 3, 0.02, 2),
]:
    dbscan(x, eps, mp)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(3):
        l = dbscan(x, eps, mp)
    ev[1].record()
    torch.cuda.synchronize()
    print(f"{name}: {ev[0].elapsed_time(ev[1]) / 3:.2f} ms per clustering, {int(l.max()) + 1} clusters, {(l < 0).sum().item()} noise")
