"""GPU parity of the GraphTCN forward and of the EC / condensation losses against the
golden vectors of the reference's own classes and its known-answer test values
(reference tests/test_losses.py:112-123).

Tolerances: model outputs 1e-5 (relative to the output scale).  The reference pins its
losses in float64; the CUDA losses compute per-element terms in fp32 and accumulate in
fp64, so they are compared at rel 2e-5 (fp32 rounding of the inputs and of atanh /
log / sqrt), counts bit-exact."""
import pytest
import torch

from tests.golden.common import case_inputs
from tests.test_gpu_in_parity import _Data, _cases, _dev, close

pytestmark = pytest.mark.gpu
LOSS_RTOL = 2e-5


@pytest.mark.parametrize("name", _cases("tcn"))
def test_graph_tcn_vs_reference_golden(name, golden_models, golden_graphs):
    from gnn_tracking_b200.models.track_condensation_networks import GraphTCN
    c = golden_models[name]
    gd = _dev(case_inputs(c, golden_graphs))
    m = GraphTCN(**c["kwargs"]).cuda()
    m.load_state_dict(c["state_dict"])
    data = _Data(**gd)
    with torch.no_grad():
        out = m(data)
    ref = c["outputs"]
    assert torch.equal(out["ec_edge_mask"].cpu(), ref["ec_edge_mask"])
    assert torch.equal(out["ec_hit_mask"].cpu(), ref["ec_hit_mask"])
    close(out["W"], ref["W"], what="W")
    close(out["H"], ref["H"], what="H")
    close(out["B"], ref["B"], what="B")
    # the reference attaches the EC output to the caller's data object
    assert data.edge_weights.shape == (gd["edge_index"].size(1), 1)


def _loss_close(got, ref, what):
    g, r = float(got), float(ref)
    if r != r:
        assert g != g, what  # NaN stays NaN (no noise hits)
    else:
        assert g == pytest.approx(r, rel=LOSS_RTOL, abs=1e-7), what


@pytest.mark.parametrize("wname", ["w_default", "w_wide"])
def test_ec_losses_vs_reference(wname, golden_losses, golden_models, golden_graphs):
    from gnn_tracking_b200.metrics.losses.ec import EdgeWeightBCELoss, EdgeWeightFocalLoss, HaughtyFocalLoss
    gd = golden_graphs["sector0"]
    w = golden_models["ec_default_h64_sector0" if wname == "w_default" else "ec_wide64_sector0"]["outputs"]["W"]
    args = dict(w=w.cuda(), y=gd["y"].cuda(), edge_index=gd["edge_index"].cuda(), pt=gd["pt"].cuda())
    ref = golden_losses["ec_losses"][wname]
    fns = {
        "bce": EdgeWeightBCELoss(), "bce_pt0.9": EdgeWeightBCELoss(pt_thld=0.9),
        "focal": EdgeWeightFocalLoss(), "focal_a0.4_g1.5_pt0.5": EdgeWeightFocalLoss(alpha=0.4, gamma=1.5, pt_thld=0.5),
        "focal_pw": EdgeWeightFocalLoss(pos_weight=torch.tensor([2.5])),
        "haughty": HaughtyFocalLoss(), "haughty_pt0.9": HaughtyFocalLoss(pt_thld=0.9),
    }
    for k, fn in fns.items():
        with torch.no_grad():
            _loss_close(fn(**args), ref[k], k)


def test_ec_loss_bool_labels_and_saturated_weights():
    from gnn_tracking_b200.metrics.losses.ec import EdgeWeightBCELoss
    from oracle import losses_oracle as L
    w = torch.tensor([0.0, 1.0, 0.3, 0.999, 1e-30])
    y = torch.tensor([True, False, True, False, True])
    ref = L.edge_weight_bce(w=w, y=y)
    with torch.no_grad():
        got = EdgeWeightBCELoss()(w=w.cuda(), y=y.cuda())
    _loss_close(got, ref, "bce with clamped logs")


@pytest.mark.parametrize("td", ["td1", "td2"])
def test_condensation_tiger_known_answers(td, golden_losses):
    """The reference's own known-answer values (tests/test_losses.py:112-123)."""
    from gnn_tracking_b200.metrics.losses.oc import CondensationLossTiger
    d = {k: v.cuda() for k, v in golden_losses[td]["data"].items()}
    ka = golden_losses["known_answers"][f"{td}_condensation"]
    with torch.no_grad():
        r = CondensationLossTiger()(beta=d["beta"].float(), x=d["x"].float(), particle_id=d["particle_id"],
                                    reconstructable=d["reconstructable"], pt=d["pt"], eta=d["eta"])
    for k, v in ka.items():
        _loss_close(r.loss_dct[k], v, f"{td}.{k}")
    res = golden_losses[td]["results"]
    assert int(r.extra_metrics["n_rep"]) == int(res["tiger_default"]["n_rep"])
    if "tiger_alt" in res:
        with torch.no_grad():
            r = CondensationLossTiger(q_min=0.1, pt_thld=0.3, max_eta=3.5)(
                beta=d["beta"].float(), x=d["x"].float(), particle_id=d["particle_id"],
                reconstructable=d["reconstructable"], pt=d["pt"], eta=d["eta"])
        for k in ("attractive", "repulsive", "coward", "noise"):
            _loss_close(r.loss_dct[k], res["tiger_alt"][k], f"{td}.alt.{k}")
        assert int(r.extra_metrics["n_rep"]) == int(res["tiger_alt"]["n_rep"])


@pytest.mark.parametrize("cname", ["tcn_default_sector0", "tcn_orphans_sector1"])
def test_condensation_tiger_on_tcn_output(cname, golden_losses, golden_models, golden_graphs):
    from gnn_tracking_b200.metrics.losses.oc import CondensationLossTiger
    c = golden_models[cname]
    gd = _dev(golden_graphs[c["graph"]])
    o = {k: v.cuda() for k, v in c["outputs"].items()}
    with torch.no_grad():
        r = CondensationLossTiger(pt_thld=0.5)(beta=o["B"], x=o["H"], particle_id=gd["particle_id"],
                                               reconstructable=gd["reconstructable"], pt=gd["pt"], eta=gd["eta"],
                                               ec_hit_mask=o["ec_hit_mask"])
    ref = golden_losses[f"tiger_{cname}"]
    # the fp32 reference goes through torch.cdist's matmul path for > 25 rows, whose own
    # cancellation error is ~1e-4 relative on the repulsive term: compare at 1e-3 there.
    for k in ("attractive", "coward", "noise"):
        _loss_close(r.loss_dct[k], ref[k], f"{cname}.{k}")
    assert float(r.loss_dct["repulsive"]) == pytest.approx(float(ref["repulsive"]), rel=1e-3)
    assert abs(int(r.extra_metrics["n_rep"]) - int(ref["n_rep"])) <= 2 + int(ref["n_rep"]) // 10000
    assert float(r.loss) == pytest.approx(float(r.loss_dct["attractive"] + r.loss_dct["repulsive"]), rel=1e-6)


def test_condensation_tiger_large_vs_oracle():
    """Seeded 20k hits x ~1.8k condensation points (spans several CP tiles / grid.y
    slices) against the float64 CPU oracle."""
    from gnn_tracking_b200.metrics.losses.oc import condensation_loss_tiger
    from oracle import losses_oracle as L
    gen = torch.Generator().manual_seed(9)
    n = 20000
    pid = torch.randint(0, 2000, (n,), generator=gen)
    mask = (torch.rand(n, generator=gen) < 0.8) & (pid > 0)
    beta = torch.rand(n, generator=gen).clamp(1e-3, 1 - 1e-3)
    x = torch.randn(n, 3, generator=gen) * 2
    ref, extra = L.condensation_tiger(beta=beta.double(), x=x.double(), object_id=pid, object_mask=mask)
    with torch.no_grad():
        got, gx = condensation_loss_tiger(beta=beta.cuda(), x=x.cuda(), object_id=pid.cuda(),
                                          object_mask=mask.cuda(), q_min=0.01)
    assert torch.equal(gx["unique_ids"].cpu(), extra["unique_ids"])
    assert torch.equal(gx["alphas"].cpu().long(), extra["alphas"])
    for k in ref:
        _loss_close(got[k], ref[k], k)
    assert abs(int(gx["n_rep"]) - int(extra["n_rep"])) <= 3  # pairs with dist within 1 ulp of 1


def test_condensation_losses_sampling_options():
    """``sample_pids < 1`` (oc.py:221-225, 410-414) draws the reference's own per-hit mask -- the same
    ``torch.rand_like(beta, dtype=float16)`` call on the same generator -- and ``max_n_rep`` (oc.py:320-328) returns
    the full repulsive sum, the expectation of the reference's sub-sampled estimate."""
    from gnn_tracking_b200.metrics.losses.oc import CondensationLossRG, CondensationLossTiger
    from oracle import losses_oracle as L
    gen = torch.Generator().manual_seed(21)
    n = 6000
    pid = torch.randint(0, 500, (n,), generator=gen)
    beta = torch.rand(n, generator=gen).clamp(1e-3, 1 - 1e-3)
    x = torch.randn(n, 3, generator=gen) * 2
    pt = torch.rand(n, generator=gen) * 3
    eta = torch.randn(n, generator=gen) * 2
    rec = torch.rand(n, generator=gen) < 0.9
    args = dict(beta=beta.cuda(), x=x.cuda(), particle_id=pid.cuda(), reconstructable=rec.cuda(), pt=pt.cuda(), eta=eta.cuda())
    # the mask the loss will draw
    torch.manual_seed(77)
    drawn = (torch.rand_like(args["beta"], dtype=torch.float16) < 0.5).cpu()
    base = L.good_node_mask(pt=pt, particle_id=pid, reconstructable=rec, eta=eta, pt_thld=0.9, max_eta=4.0)
    ref, _ = L.condensation_tiger(beta=beta.double(), x=x.double(), object_id=pid, object_mask=base & drawn)
    torch.manual_seed(77)
    with torch.no_grad():
        got = CondensationLossTiger(sample_pids=0.5, max_n_rep=1000)(**args)
    for k in ref:
        _loss_close(got.loss_dct[k], ref[k], k)
    # the same switch on the radius-graph variant: fewer masked hits than without it, still finite
    torch.manual_seed(77)
    with torch.no_grad():
        rg_half = CondensationLossRG(sample_pids=0.5)(**args)
        rg_full = CondensationLossRG()(**args)
    assert all(torch.isfinite(v) for v in rg_half.loss_dct.values())
    assert float(rg_half.loss_dct["attractive"]) != float(rg_full.loss_dct["attractive"])
