"""Plan cache behaviour the reference's callers rely on (round-1 advisor findings): forwards under
``torch.inference_mode()`` (Lightning's validation / test / predict loops), eviction with the graph, and the
deferred range check of ``edge_index`` (the reference raises an IndexError inside index_select)."""
import gc

import pytest
import torch

pytestmark = pytest.mark.gpu


def _graph(n=500, e=4000, seed=0):
    gen = torch.Generator().manual_seed(seed)
    return (torch.randn(n, 14, generator=gen).cuda(), torch.randint(0, n, (2, e), generator=gen).cuda(),
            torch.randn(e, 4, generator=gen).cuda())


def test_forward_under_inference_mode_matches_no_grad():
    from gnn_tracking_b200.models.edge_classifier import ECForGraphTCN
    torch.manual_seed(0)
    m = ECForGraphTCN(node_indim=14, edge_indim=4, hidden_dim=64, L_ec=2).cuda()
    x, ei, ea = _graph()
    with torch.no_grad():
        ref = m.forward_tensors(x, ei, ea)["W"].clone()
    with torch.inference_mode():
        xi, eii, eai = x.clone(), ei.clone(), ea.clone()  # inference tensors: no version counter
        assert eii.is_inference()
        w1 = m.forward_tensors(xi, eii, eai)["W"]
        w2 = m.forward_tensors(xi, eii, eai)["W"]          # second call: plan cache hit
    # same kernels, same inputs: equal up to the order of the atomic adds of the per-destination sums (one ulp;
    # tests/cuda/determinism_probe.py: 3e-8 .. 1.2e-7 relative between any two runs)
    assert float((w1 - ref).abs().max()) <= 1e-6 and float((w2 - ref).abs().max()) <= 1e-6


def test_plan_cache_entry_dies_with_the_graph():
    from gnn_tracking_b200 import plan as P
    P.clear_plan_cache()
    x, ei, ea = _graph(seed=1)
    P.get_plan(ei, x.size(0))
    ei2 = ei.clone()
    P.adopt_plan(ei2, x.size(0), P.build_plan(ei2, x.size(0)))
    assert len(P._CACHE) == 2
    del ei, ei2
    gc.collect()
    assert len(P._CACHE) == 0


def test_out_of_range_edge_index_is_reported_without_touching_foreign_memory():
    from gnn_tracking_b200 import plan as P
    from gnn_tracking_b200.models.edge_classifier import ECForGraphTCN
    P.clear_plan_cache()
    x, ei, ea = _graph(seed=2)
    bad = ei.clone()
    bad[0, 7] = x.size(0) + 5
    bad[1, 11] = -3
    torch.manual_seed(0)
    m = ECForGraphTCN(node_indim=14, edge_indim=4, hidden_dim=64, L_ec=1).cuda()
    with torch.no_grad():
        m.forward_tensors(x, bad, ea)        # ids are clamped on the device: no out-of-bounds access
    torch.cuda.synchronize()
    with pytest.raises(IndexError):          # ... and the next plan call raises what index_select would have
        P.get_plan(ei, x.size(0))
    P.get_plan(ei, x.size(0))                # raised once


@pytest.mark.gpu
@pytest.mark.parametrize("n,e,keep_frac", [(5000, 40000, 0.3), (5000, 40000, 0.0), (300, 20, 1.0), (100000, 1000000, 0.25)])
def test_prune_orphans_equals_rebuild(n, e, keep_frac):
    """``gtb_plan_prune_orphans`` (no second sort) against a plan rebuilt from the relabelled sub-graph the
    way the reference derives it (track_condensation_networks.py:251-259): every array bit-exact."""
    from gnn_tracking_b200.plan import build_plan, prune_orphans
    gen = torch.Generator().manual_seed(n + e)
    ei = torch.randint(0, n, (2, e), generator=gen)
    ei[:, : e // 4] = ei[:, : e // 4] % max(1, n // 3)  # a crowded region and many orphans elsewhere
    keep = torch.rand(e, generator=gen) < keep_frac
    eid = ei.cuda()
    parent = build_plan(eid, n)
    sub, _, kept = parent.filtered(keep.cuda())
    pruned, node_ids, new_id = prune_orphans(sub)
    torch.cuda.synchronize()
    sub_ei = ei[:, keep]
    connected = torch.unique(sub_ei)
    relabel = torch.full((n,), -1, dtype=torch.long)
    relabel[connected] = torch.arange(connected.numel())
    assert torch.equal(node_ids.cpu().long(), connected)
    assert torch.equal(new_id.cpu().long(), relabel)
    assert pruned.n_nodes == connected.numel() and pruned.n_edges == int(keep.sum())
    if pruned.n_edges:
        ref = build_plan(relabel[sub_ei].cuda(), connected.numel())
        for name in ("perm", "rowptr", "src_sorted", "dst_sorted"):
            assert torch.equal(getattr(pruned, name), getattr(ref, name)), name
    else:
        assert pruned.n_nodes == 0 and pruned.rowptr.cpu().tolist() == [0]
