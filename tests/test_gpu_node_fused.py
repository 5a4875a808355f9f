"""The one-launch node side of a 64-wide Interaction-Network layer (``gtb_in_node_fused_f32``,
csrc/node_ws.cu) against float64 arithmetic, and the stacks that use it against the CPU oracle and
against the launch-per-op path (``GTB_NO_NODE_WS=1``).  Tolerance 1e-5 * scale (fp32 parity bar)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-5


def close(a, b, tol=TOL, what=""):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(float(b.abs().max()), 1.0) if b.numel() else 1.0
    err = float((a - b).abs().max()) if a.numel() else 0.0
    assert err <= tol * scale, f"{what}: max|d|={err:.3e} scale={scale:.3e}"


def _lin(gen, o, i, bias=True):
    w = (torch.rand(o, i, generator=gen) * 2 - 1) / i ** 0.5
    b = (torch.rand(o, generator=gen) * 2 - 1) / i ** 0.5 if bias else None
    return w, b


@pytest.mark.parametrize("n", [1, 127, 300, 4096, 40000])
@pytest.mark.parametrize("mode", ["obj+proj", "obj", "proj"])
@pytest.mark.parametrize("relu_x,proj_relu,with_res", [(True, True, True), (False, False, False)])
def test_node_fused_vs_float64(n, mode, relu_x, proj_relu, with_res):
    from gnn_tracking_b200 import ops
    gen = torch.Generator().manual_seed(n + len(mode))
    x = torch.randn(n, 64, generator=gen)
    aggr = torch.randn(n, 64, generator=gen) * 3
    res = torch.randn(n, 64, generator=gen)
    (w0, b0), (w1, b1), (w2, b2) = _lin(gen, 64, 128), _lin(gen, 64, 64), _lin(gen, 64, 64)
    (wa, _), (wb, _) = _lin(gen, 64, 64, False), _lin(gen, 64, 64, False)
    obj = proj = None
    if "obj" in mode:
        obj = ops.pack_linears([w0.cuda(), w1.cuda(), w2.cuda()], [b0.cuda(), b1.cuda(), b2.cuda()], ops.IMPL_TCGEN05,
                               block_widths=[64, 64])
    if "proj" in mode:
        proj = (ops.pack_linears([wa.cuda()], [None], ops.IMPL_TCGEN05), ops.pack_linears([wb.cuda()], [None], ops.IMPL_TCGEN05))
    aggr_d = aggr.cuda()
    xo, pa, pb = ops.in_node_fused(x.cuda(), relu_x, aggr=aggr_d if obj is not None else None, zero_aggr=True, packed_obj=obj,
                                   res=res.cuda() if with_res else None, res_a=0.6, res_b=0.8, proj=proj, proj_relu=proj_relu)
    torch.cuda.synchronize()
    d = torch.float64
    xin = (torch.relu(x) if relu_x else x).to(d)
    y = None
    if obj is not None:
        h = torch.relu(torch.cat([xin, aggr.to(d)], 1) @ w0.to(d).T + b0.to(d))
        h = torch.relu(h @ w1.to(d).T + b1.to(d))
        y = 0.8 * (h @ w2.to(d).T + b2.to(d))
        if with_res:
            y = y + 0.6 * res.to(d)
        close(xo, y, what="x_out")
        assert float(aggr_d.abs().max()) == 0.0, "aggr must come back zeroed"
    else:
        assert xo is None
    if proj is not None:
        src = y if y is not None else x.to(d)
        src = torch.relu(src) if proj_relu else src
        close(pa, src @ wa.to(d).T, what="p_a")
        close(pb, src @ wb.to(d).T, what="p_b")
    else:
        assert pa is None and pb is None


def _ec(residual_type, L, **kw):
    from gnn_tracking_b200.models.edge_classifier import ECForGraphTCN
    torch.manual_seed(3)
    return ECForGraphTCN(node_indim=14, edge_indim=4, interaction_node_dim=64, interaction_edge_dim=64, hidden_dim=64,
                         L_ec=L, residual_type=residual_type, **kw)


@pytest.mark.parametrize("residual_type,L,kw", [
    ("skip1", 3, {}), ("skip1", 1, {}), ("skip2", 4, {"use_intermediate_edge_embeddings": False}), ("skip_top", 3, {"residual_kwargs": {"connect_to": 1}}),
    ("skip1", 2, {"use_node_embedding": False}), ("skip1", 2, {"use_intermediate_edge_embeddings": False}),
    ("skip1", 2, {"alpha": 0.0})])
def test_ec_fused_stack_vs_oracle_and_unfused(residual_type, L, kw, monkeypatch):
    from gnn_tracking_b200 import ops
    from oracle import in_oracle as O
    gen = torch.Generator().manual_seed(11)
    n, e = 3000, 40011
    x = torch.randn(n, 14, generator=gen)
    ei = torch.randint(0, n, (2, e), generator=gen)
    ei[1, :600] = 7  # one heavy destination
    ea = torch.randn(e, 4, generator=gen)
    m = _ec(residual_type, L, **kw)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        ref = O.ec_forward(x, ei, ea, sd, alpha=kw.get("alpha", 0.5), residual_type=residual_type,
                           use_intermediate_edge_embeddings=kw.get("use_intermediate_edge_embeddings", True),
                           use_node_embedding=kw.get("use_node_embedding", True), residual_kwargs=kw.get("residual_kwargs"))
    m = m.cuda()
    xd, eid, ead = x.cuda(), ei.cuda(), ea.cuda()
    with torch.no_grad():
        l0 = ops.launch_count()
        out = m.forward_tensors(xd, eid, ead)
        fused_launches = ops.launch_count() - l0
        l0 = ops.launch_count()
        out2 = m.forward_tensors(xd, eid, ead)  # second call: the aggregate came back zeroed, packs are cached
        fused_launches2 = ops.launch_count() - l0
    monkeypatch.setenv("GTB_NO_NODE_WS", "1")
    with torch.no_grad():
        m.forward_tensors(xd, eid, ead)  # re-packs for the other calling pattern
        l0 = ops.launch_count()
        plain = m.forward_tensors(xd, eid, ead)
        plain_launches = ops.launch_count() - l0
    for k in ("W", "node_embedding", "edge_embedding"):
        close(out[k], ref[k], what=f"fused vs oracle {k}")
        close(out2[k], ref[k], what=f"fused (2nd call) vs oracle {k}")
        close(plain[k], ref[k], what=f"unfused vs oracle {k}")
    assert fused_launches2 < plain_launches, (fused_launches2, plain_launches)


def test_fused_path_is_off_under_grad_and_small_graphs():
    m = _ec("skip1", 2).cuda()
    net = m.ec_resin.network
    h, ea = torch.randn(100, 64, device="cuda"), torch.randn(1000, 64, device="cuda")
    assert not net.fused_ok(h, ea)  # parameters require grad and grad mode is on
    with torch.no_grad():
        assert net.fused_ok(h, ea)
        assert not net.fused_ok(h, ea[:150])  # fewer than two edges per node: node blocks are not pre-projected
    x = torch.randn(100, 14, device="cuda")
    ei = torch.randint(0, 100, (2, 1000), device="cuda")
    out = m.forward_tensors(x, ei, torch.randn(1000, 4, device="cuda"))
    out["W"].sum().backward()  # the autograd path still works on the same module
    assert m.W.layers[0].weight.grad is not None


@pytest.mark.parametrize("n", [1, 127, 128, 300, 70001])
@pytest.mark.parametrize("gather,bias,final_relu", [(True, False, True), (False, True, False), (True, True, True)])
def test_edge_encoder_vs_float64(n, gather, bias, final_relu):
    """``gtb_edge_encoder_f32`` (csrc/enc_ws.cu): 4 -> 64 -> 64 over rows gathered through a permutation."""
    from gnn_tracking_b200 import ops
    gen = torch.Generator().manual_seed(5 * n + gather)
    x = torch.randn(n, 4, generator=gen) * 2
    (w0, b0), (w1, b1) = _lin(gen, 64, 4, bias), _lin(gen, 64, 64, bias)
    perm = torch.randperm(n, generator=gen).to(torch.int32) if gather else None
    packed = ops.pack_linears([w1.cuda()], [b1.cuda() if bias else None], ops.IMPL_TCGEN05)
    out = ops.edge_encoder(x.cuda(), perm.cuda() if gather else None, n, w0.cuda(), b0.cuda() if bias else None, packed, final_relu)
    torch.cuda.synchronize()
    d = torch.float64
    xs = x[perm.long()] if gather else x
    h = xs.to(d) @ w0.to(d).T
    if bias:
        h = h + b0.to(d)
    y = torch.relu(h) @ w1.to(d).T
    if bias:
        y = y + b1.to(d)
    close(out, torch.relu(y) if final_relu else y, what="encoder")


def test_ec_uses_the_one_launch_encoder(monkeypatch):
    from gnn_tracking_b200 import ops
    m = _ec("skip1", 1).cuda()
    x = torch.randn(500, 14, device="cuda")
    ei = torch.randint(0, 500, (2, 6000), device="cuda")
    ea = torch.randn(6000, 4, device="cuda")
    with torch.no_grad():
        assert m.ec_edge_encoder.k4_ok(ea)
        a = m.forward_tensors(x, ei, ea)
        monkeypatch.setenv("GTB_NO_ENC_WS", "1")
        assert not m.ec_edge_encoder.k4_ok(ea)
        b = m.forward_tensors(x, ei, ea)
    for k in ("W", "node_embedding", "edge_embedding"):
        close(a[k], b[k], tol=2e-6, what=k)
