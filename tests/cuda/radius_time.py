"""Radius graph at the latent-space size of config 5: cell list against the all-pairs walk (CUDA events, 5 calls each)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from gnn_tracking_b200.cluster import radius_graph  # noqa: E402

out = {}
for n, d, r in ((100_000, 3, 0.02), (100_000, 8, 0.35)):
    gen = torch.Generator().manual_seed(0)
    x = torch.rand(n, d, generator=gen).cuda()
    row = {}
    for method in ("grid", "brute"):
        e = radius_graph(x, r, max_num_neighbors=64, method=method)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(5):
            e = radius_graph(x, r, max_num_neighbors=64, method=method)
        t1.record()
        torch.cuda.synchronize()
        row[method] = {"ms": t0.elapsed_time(t1) / 5, "edges": int(e.size(1))}
        row.setdefault("same", True)
        if method == "brute":
            row["same"] = bool(torch.equal(e, first))
        first = e
    out[f"n{n}_d{d}_r{r}"] = row
print(json.dumps(out))
