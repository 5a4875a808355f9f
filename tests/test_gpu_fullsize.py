"""BASELINE.json full-size case (100k nodes / 1M edges, hidden 64) through size-independent
properties, where the CPU oracle is too slow to be the checker on every run:
plan invariants (bit-exact), checksum-of-checksums of the segmented sum, edge-order equivariance
of the edge classifier, idempotence of the plan filter."""
import pytest
import torch

pytestmark = pytest.mark.gpu
N, E = 100_000, 1_000_000


@pytest.fixture(scope="module")
def graph():
    gen = torch.Generator().manual_seed(2024)
    src = torch.randint(0, N, (E,), generator=gen)
    dst = (src + torch.randint(-300, 301, (E,), generator=gen)).clamp(0, N - 1)
    dst[:5000] = 77  # one very heavy destination
    return {"edge_index": torch.stack([src, dst]).cuda(), "x": torch.randn(N, 14, generator=gen).cuda(),
            "edge_attr": torch.randn(E, 4, generator=gen).cuda()}


def test_plan_invariants_full_size(graph):
    from gnn_tracking_b200.plan import build_plan
    ei = graph["edge_index"]
    plan = build_plan(ei, N)
    plan.validate()
    perm = plan.perm.long()
    assert torch.equal(torch.sort(perm).values, torch.arange(E, device="cuda"))       # a permutation
    d = plan.dst_sorted.long()
    assert torch.all(d[1:] >= d[:-1])                                                   # sorted
    same = d[1:] == d[:-1]
    assert torch.all(perm[1:][same] > perm[:-1][same])                                  # stable
    assert torch.equal(d, ei[1][perm]) and torch.equal(plan.src_sorted.long(), ei[0][perm])
    rp = plan.rowptr.long()
    assert rp[0] == 0 and rp[-1] == E and torch.all(rp[1:] >= rp[:-1])
    assert torch.equal(rp[1:] - rp[:-1], torch.bincount(ei[1], minlength=N))
    # filtering with an all-true mask is the identity
    sub, new_id, kept = plan.filtered(torch.ones(E, dtype=torch.bool, device="cuda"))
    for f in ("perm", "rowptr", "src_sorted", "dst_sorted"):
        assert torch.equal(getattr(sub, f), getattr(plan, f))
    assert torch.equal(new_id.long(), torch.arange(E, device="cuda"))


@pytest.mark.parametrize("impl", ["ffma", "tcgen05"])
def test_segmented_sum_checksum_full_size(graph, impl):
    """sum over nodes of the aggregate == sum over edges of the messages (per column), and the
    aggregate of the heavy node equals the direct sum of its 5000+ messages."""
    from gnn_tracking_b200 import ops
    from gnn_tracking_b200.ops import Block
    from gnn_tracking_b200.plan import build_plan
    plan = build_plan(graph["edge_index"], N)
    gen = torch.Generator().manual_seed(5)
    ws = [(torch.randn(64, 64, generator=gen) / 8).cuda() for _ in range(3)]
    bs = [(torch.randn(64, generator=gen) * 0.1).cuda() for _ in range(3)]
    ee = torch.randn(E, 64, generator=gen).cuda()
    packed = ops.pack_linears(ws, bs, ops.IMPL_FFMA if impl == "ffma" else ops.IMPL_TCGEN05)
    aggr = torch.zeros(N, 64, device="cuda")
    out = ops.fused_mlp([Block(ee, plan.perm)], E, packed, out_index=plan.perm, aggr=aggr, seg_id=plan.dst_sorted,
                        rowptr=plan.rowptr)
    torch.cuda.synchronize()
    tot_e, tot_a = out.double().sum(0), aggr.double().sum(0)
    scale = out.double().abs().sum(0)
    assert torch.all((tot_e - tot_a).abs() <= 1e-6 * scale)
    heavy = out[graph["edge_index"][1] == 77].double().sum(0)
    assert torch.all((aggr[77].double() - heavy).abs() <= 1e-5 * heavy.abs().clamp_min(1.0))
    deg0 = torch.bincount(graph["edge_index"][1], minlength=N) == 0
    assert torch.all(aggr[deg0] == 0)  # isolated nodes keep an exact zero


def test_edge_order_equivariance_full_size(graph):
    """Shuffling the caller's edge order permutes W / edge embeddings and leaves the node
    embedding unchanged (up to the summation order of the aggregation: 1e-5)."""
    from gnn_tracking_b200.models.edge_classifier import ECForGraphTCN
    torch.manual_seed(0)
    m = ECForGraphTCN(node_indim=14, edge_indim=4, L_ec=3, interaction_node_dim=64, interaction_edge_dim=64,
                      hidden_dim=64).cuda()
    shuf = torch.randperm(E, generator=torch.Generator().manual_seed(1)).cuda()
    with torch.no_grad():
        a = m.forward_tensors(graph["x"], graph["edge_index"], graph["edge_attr"])
        b = m.forward_tensors(graph["x"], graph["edge_index"][:, shuf].contiguous(), graph["edge_attr"][shuf].contiguous())
    assert float((a["W"][shuf] - b["W"]).abs().max()) <= 1e-5
    s = float(a["edge_embedding"].abs().max().clamp_min(1.0))
    assert float((a["edge_embedding"][shuf] - b["edge_embedding"]).abs().max()) <= 1e-5 * s
    s = float(a["node_embedding"].abs().max().clamp_min(1.0))
    assert float((a["node_embedding"] - b["node_embedding"]).abs().max()) <= 1e-5 * s
    assert bool(torch.isfinite(a["W"]).all()) and float(a["W"].min()) >= 0.001 - 1e-6 and float(a["W"].max()) <= 0.999 + 1e-6
