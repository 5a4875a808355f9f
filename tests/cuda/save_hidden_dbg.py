import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gnn_tracking_b200 import ops
from gnn_tracking_b200.ops import Block, ACT_RELU
from gnn_tracking_b200.models.interaction_network import InteractionNetwork
from gnn_tracking_b200.plan import get_plan
torch.manual_seed(2)
n, e = 700, 9000
gen = torch.Generator().manual_seed(1)
ei = torch.randint(0, n, (2, e), generator=gen).cuda()
x = torch.randn(n, 64, generator=gen).cuda(); ea = torch.randn(e, 64, generator=gen).cuda()
m = InteractionNetwork(node_indim=64, edge_indim=64, node_outdim=64, edge_outdim=64, node_hidden_dim=64, edge_hidden_dim=64).cuda()
plan = get_plan(ei, n)
rel = m.relational_model
blocks = [Block(x, plan.dst_sorted, False, sorted_index=True), Block(x, plan.src_sorted, False), Block(ea, plan.perm, False, unique_index=True)]
from gnn_tracking_b200.models.mlp import _run_nograd
hid = []
aggr = torch.zeros(n, 64, device="cuda")
with torch.no_grad():
    out = _run_nograd(rel._cache, rel.linears, blocks, e, out_index=plan.perm, aggr=aggr, seg_id=plan.dst_sorted, rowptr=plan.rowptr, save_hidden=hid)
    print("saved:", len(hid))
    lin = rel.linears
    cat = torch.cat([x[plan.dst_sorted.long()], x[plan.src_sorted.long()], ea[plan.perm.long()]], 1).double()
    h0 = torch.relu(cat @ lin[0].weight.double().T + lin[0].bias.double())
    h1 = torch.relu(h0 @ lin[1].weight.double().T + lin[1].bias.double())
    for name, got, ref in (("h0", hid[0], h0), ("h1", hid[1], h1)):
        d = (got.double() - ref).abs()
        bad = (d > 1e-4).any(1).nonzero().flatten()
        print(name, "max err", float(d.max()), "bad rows", bad.numel(), bad[:20].tolist(), "cols of first bad", (d[bad[0]] > 1e-4).nonzero().flatten().tolist() if bad.numel() else None)
with torch.no_grad():
    aggr2 = torch.zeros(n, 64, device="cuda")
    out2 = _run_nograd(rel._cache, rel.linears, blocks, e, out_index=plan.perm, aggr=aggr2, seg_id=plan.dst_sorted, rowptr=plan.rowptr)
    z = h1 @ lin[2].weight.double().T + lin[2].bias.double()
    ref_out = torch.empty_like(z); ref_out[plan.perm.long()] = z
    ref_aggr = torch.zeros(n, 64, dtype=torch.float64, device="cuda").index_add_(0, plan.dst_sorted.long(), z)
    for name, got, ref in (("out(save)", out, ref_out), ("out(plain)", out2, ref_out), ("aggr(save)", aggr, ref_aggr), ("aggr(plain)", aggr2, ref_aggr)):
        d = (got.double() - ref).abs()
        bad = (d > 1e-4).any(1).nonzero().flatten()
        print(name, "max err", float(d.max()), "bad rows", bad.numel(), bad[:12].tolist())

def grads(env):
    if env: os.environ["GTB_NO_SAVE_HIDDEN"] = "1"
    else: os.environ.pop("GTB_NO_SAVE_HIDDEN", None)
    m.zero_grad()
    xc, ec = x.clone().requires_grad_(), ea.clone().requires_grad_()
    g = torch.Generator().manual_seed(5)
    gx, ge = torch.randn(n, 64, generator=g).cuda(), torch.randn(e, 64, generator=g).cuda()
    xt, et = m(xc, ei, ec)
    ((xt * gx).sum() + (et * ge).sum()).backward()
    return {"x": xc.grad.clone(), "e": ec.grad.clone(), **{k: p.grad.clone() for k, p in m.named_parameters()}}
a, b = grads(False), grads(True)
for k in a:
    d = (a[k] - b[k]).abs().max().item(); s = b[k].abs().max().item()
    print(f"{k:40s} diff {d:.3e} scale {s:.3e}")

import gnn_tracking_b200.autograd as AG
orig = ops.fused_mlp
log = []
def spy(blocks, n_rows, packed, **kw):
    out = orig(blocks, n_rows, packed, **kw)
    if kw.get("gate") is not None:
        log.append((blocks[0].tensor.clone(), kw["gate"].clone(), out.clone(), packed.buf.clone()))
    return out
ops.fused_mlp = spy
la = []; lb = []
log = la; grads(False)
log = []
def spy2(blocks, n_rows, packed, **kw):
    out = orig(blocks, n_rows, packed, **kw)
    if kw.get("gate") is not None:
        lb.append((blocks[0].tensor.clone(), kw["gate"].clone(), out.clone(), packed.buf.clone()))
    return out
def spy1(blocks, n_rows, packed, **kw):
    out = orig(blocks, n_rows, packed, **kw)
    if kw.get("gate") is not None:
        la.append((blocks[0].tensor.clone(), kw["gate"].clone(), out.clone(), packed.buf.clone()))
    return out
la.clear(); ops.fused_mlp = spy1; grads(False)
ops.fused_mlp = spy2; grads(True)
print("gated launches", len(la), len(lb))
for i, (p, q) in enumerate(zip(la, lb)):
    print(i, "rows", p[0].shape, "in diff", (p[0] - q[0]).abs().max().item(), "gate diff", (p[1] - q[1]).abs().max().item(),
          "mask diff rows", ((p[1] > 0) != (q[1] > 0)).any(1).sum().item(), "out diff", (p[2] - q[2]).abs().max().item(),
          "pack diff", (p[3] != q[3]).sum().item())
