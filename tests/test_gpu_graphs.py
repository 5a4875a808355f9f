"""``graphs.CapturedForward``: the edge-classifier forward (plan build included) as one CUDA graph launch gives
the eager forward's results on every new graph of the captured shape."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _graph(n, e, seed):
    gen = torch.Generator().manual_seed(seed)
    return (torch.randn(n, 14, generator=gen).cuda(), torch.randint(0, n, (2, e), generator=gen).cuda(),
            torch.randn(e, 4, generator=gen).cuda())


@pytest.mark.parametrize("kw", [dict(interaction_node_dim=64, interaction_edge_dim=64, hidden_dim=64, L_ec=3), dict(hidden_dim=64, L_ec=2)])
def test_captured_forward_matches_eager(kw):
    from gnn_tracking_b200.graphs import CapturedForward
    from gnn_tracking_b200.models.edge_classifier import ECForGraphTCN
    torch.manual_seed(0)
    m = ECForGraphTCN(node_indim=14, edge_indim=4, **kw).cuda()
    n, e = 3000, 40011
    fwd = CapturedForward(m, *_graph(n, e, 0))
    assert fwd.launches > 0
    for seed in (1, 2, 3):
        x, ei, ea = _graph(n, e, seed)
        with torch.no_grad():
            ref = {k: v.clone() for k, v in m.forward_tensors(x, ei, ea).items()}
        out = fwd(x, ei, ea)
        torch.cuda.synchronize()
        for k in ref:
            scale = max(1.0, float(ref[k].abs().max()))
            assert float((out[k] - ref[k]).abs().max()) <= 2e-6 * scale, (seed, k)  # atomic-add order only
    assert not fwd.fits(*_graph(n, e + 1, 0))
    with pytest.raises(ValueError):
        fwd(*_graph(n + 1, e, 0))
    # the eager path still works after a capture (no scratch shared with the graph's memory pool)
    with torch.no_grad():
        m.forward_tensors(*_graph(500, 6000, 5))


def test_captured_forward_refuses_stale_weights():
    from gnn_tracking_b200.graphs import CapturedForward
    from gnn_tracking_b200.models.edge_classifier import ECForGraphTCN
    torch.manual_seed(0)
    m = ECForGraphTCN(node_indim=14, edge_indim=4, hidden_dim=64, L_ec=1).cuda()
    g = _graph(500, 6000, 0)
    fwd = CapturedForward(m, *g)
    fwd(*g)
    with torch.no_grad():
        m.W.layers[0].weight.add_(1.0)  # bumps the version counter, as an optimizer step does
    with pytest.raises(RuntimeError, match="capture again"):
        fwd(*g)
