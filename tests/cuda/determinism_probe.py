"""Repeats one EC forward and reports bitwise mismatches between runs (summation-order effects vs real bugs)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gnn_tracking_b200.models.edge_classifier import ECForGraphTCN
from gnn_tracking_b200.plan import clear_plan_cache

def graph(n, e, seed=0):
    gen = torch.Generator().manual_seed(seed)
    return (torch.randn(n, 14, generator=gen).cuda(), torch.randint(0, n, (2, e), generator=gen).cuda(), torch.randn(e, 4, generator=gen).cuda())

for label, kw, (n, e) in (("default dims 500/4000", dict(hidden_dim=64, L_ec=2), (500, 4000)),
                          ("wide 3000/40000", dict(hidden_dim=64, L_ec=3, interaction_node_dim=64, interaction_edge_dim=64), (3000, 40000)),
                          ("wide 100k/1M", dict(hidden_dim=64, L_ec=3, interaction_node_dim=64, interaction_edge_dim=64), (100000, 1000000))):
    torch.manual_seed(0)
    m = ECForGraphTCN(node_indim=14, edge_indim=4, **kw).cuda()
    x, ei, ea = graph(n, e)
    junk = []
    with torch.no_grad():
        ref = {k: v.clone() for k, v in m.forward_tensors(x, ei, ea).items()}
        bad = {k: 0 for k in ref}; worst = {k: 0.0 for k in ref}
        reps = 30 if n > 10000 else 200
        for i in range(reps):
            if i % 3 == 0:
                junk.append(torch.randn(1 + 977 * (i % 7), 33, device="cuda"))  # perturb the allocator
                clear_plan_cache()
            xi, eii, eai = x.clone(), ei.clone(), ea.clone()
            out = m.forward_tensors(xi, eii, eai)
            for k in ref:
                d = (out[k] - ref[k]).abs().max().item()
                if d != 0.0:
                    bad[k] += 1; worst[k] = max(worst[k], d / max(1.0, ref[k].abs().max().item()))
    print(label, "mismatching runs of", reps, bad, "worst rel diff", worst)
