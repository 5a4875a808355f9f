"""One launch of the W head kernel (csrc/head_ws.cu) on random tensors against float64 (used under compute-sanitizer)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gnn_tracking_b200 import ops
from gnn_tracking_b200.ops import ACT_SIGMOID_AFFINE, Block
from gnn_tracking_b200.models.mlp import MLP
torch.manual_seed(0)
n, e = 300, 2000 + 37
gen = torch.Generator().manual_seed(3)
h = torch.randn(n, 64, generator=gen).cuda()
es = [torch.randn(e, 64, generator=gen).cuda() for _ in range(4)]
src = torch.randint(0, n, (e,), generator=gen).to(torch.int32).cuda()
dst = torch.sort(torch.randint(0, n, (e,), generator=gen)).values.to(torch.int32).cuda()
perm = torch.randperm(e, generator=gen).to(torch.int32).cuda()
W = MLP(384, 1, 64, L=3).cuda()
blocks = [Block(h, src), Block(h, dst, sorted_index=True), Block(es[0]), Block(es[1]), Block(es[2]), Block(es[3], perm, unique_index=True)]
l0 = ops.launch_count()
with torch.no_grad():
    w = W.forward_blocks(blocks, e, final_act=ACT_SIGMOID_AFFINE, act_eps=0.001, out_index=perm)
torch.cuda.synchronize()
d = torch.float64
lin = W.linears
cat = torch.cat([h[src.long()], h[dst.long()], es[0], es[1], es[2], es[3][perm.long()]], 1).to(d)
z = torch.relu(cat @ lin[0].weight.to(d).T + lin[0].bias.to(d))
z = torch.relu(z @ lin[1].weight.to(d).T + lin[1].bias.to(d))
z = z @ lin[2].weight.to(d).T + lin[2].bias.to(d)
ref = torch.empty_like(z); ref[perm.long()] = 0.001 + 0.998 * torch.sigmoid(z)
print("launches", ops.launch_count() - l0, "max err", float((w.to(d) - ref).abs().max()))
assert float((w.to(d) - ref).abs().max()) < 1e-5
