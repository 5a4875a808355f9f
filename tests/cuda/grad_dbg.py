import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gnn_tracking_b200.metrics.losses.ec import EdgeWeightBCELoss
from gnn_tracking_b200.models.edge_classifier import ECForGraphTCN
from oracle import in_oracle as O, losses_oracle as LO
gen = torch.Generator().manual_seed(5)
n, e = 600, 7000
ei = torch.randint(0, n, (2, e), generator=gen); ei[1, : e // 20] = 3
x = torch.randn(n, 14, generator=gen); ea = torch.randn(e, 4, generator=gen)
y = (torch.rand(e, generator=gen) < 0.3)
kw = dict(interaction_node_dim=64, interaction_edge_dim=64, hidden_dim=64, L_ec=2)
torch.manual_seed(3)
m0 = ECForGraphTCN(node_indim=14, edge_indim=4, **kw)
sd = {k: v.detach().double().requires_grad_() for k, v in m0.state_dict().items()}
ref = O.ec_forward(x.double(), ei, ea.double(), sd)
LO.bce_mean(ref["W"], y.double()).backward()
for impl in ("ffma", "auto"):
    os.environ["GTB_IMPL"] = impl
    torch.manual_seed(3)
    m = ECForGraphTCN(node_indim=14, edge_indim=4, **kw).cuda()
    out = m.forward_tensors(x.cuda(), ei.cuda(), ea.cuda())
    EdgeWeightBCELoss()(w=out["W"], y=y.cuda()).backward()
    print(impl, "W err", float((out["W"].detach().cpu().double() - ref["W"].detach()).abs().max()))
    for k, p in m.named_parameters():
        r = sd[k].grad; g = p.grad.cpu().double()
        print(f"  {k:50s} rel {float((g - r).abs().max() / r.abs().max()):.2e}  scale {float(r.abs().max()):.2e}")
