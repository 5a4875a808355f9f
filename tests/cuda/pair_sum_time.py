"""Fused pair sums (condensation repulsion, mode 1) at 100k hits: cell list against the all-pairs walk, forward + backward."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from gnn_tracking_b200.metrics.losses.metric_learning import radius_pair_sum  # noqa: E402

n, d = 100_000, 3
gen = torch.Generator().manual_seed(0)
x = torch.rand(n, d, generator=gen).cuda()
pid = torch.randint(0, 5000, (n,), generator=gen).cuda()
beta = (torch.rand(n, generator=gen) * 0.98 + 0.01).cuda()
flag = (torch.rand(n, generator=gen) < 0.05).cuda()
out = {}
for method in ("grid", "brute"):
    os.environ["GTB_RADIUS_BRUTE"] = "1" if method == "brute" else "0"

    def step():
        xg, bg = x.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        o = radius_pair_sum(x=xg, particle_id=pid, src_flag=flag, r=0.05, mode=1, beta=bg, q_min=0.01, max_num_neighbors=256)
        o[0].backward()
        return o
    o = step()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        o = step()
    t1.record()
    torch.cuda.synchronize()
    out[method] = {"fwd_bwd_ms": t0.elapsed_time(t1) / 5, "sum": float(o[0]), "edges": float(o[1])}
print(json.dumps(out))
