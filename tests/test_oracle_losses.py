"""Pins oracle/losses_oracle.py against the reference's known-answer values
(reference tests/test_losses.py:112-123, :194-203) and reference outputs."""
import pytest
import torch

from oracle import losses_oracle as L


def _f(d):
    return {k: float(v) for k, v in d.items()}


@pytest.mark.parametrize("td", ["td1", "td2"])
def test_condensation_known_answers(td, golden_losses):
    d = golden_losses[td]["data"]
    ka = golden_losses["known_answers"][f"{td}_condensation"]
    losses, _ = L.condensation_tiger_loss(beta=d["beta"], x=d["x"], particle_id=d["particle_id"],
                                          reconstructable=d["reconstructable"], pt=d["pt"], eta=d["eta"])
    assert _f(losses) == pytest.approx(ka, rel=1e-6)
    mask = L.good_node_mask(pt=d["pt"], particle_id=d["particle_id"], reconstructable=d["reconstructable"], eta=d["eta"])
    rg = L.condensation_rg(beta=d["beta"], x=d["x"], particle_id=d["particle_id"], mask=mask)
    assert _f(rg) == pytest.approx(ka, rel=1e-6)


@pytest.mark.parametrize("td", ["td1", "td2"])
def test_condensation_vs_reference_outputs(td, golden_losses):
    d = golden_losses[td]["data"]
    res = golden_losses[td]["results"]
    for name, kw in (("default", {}), ("alt", dict(q_min=0.1, pt_thld=0.3, max_eta=3.5))):
        if f"tiger_{name}" not in res:
            continue
        losses, extra = L.condensation_tiger_loss(beta=d["beta"], x=d["x"], particle_id=d["particle_id"],
                                                  reconstructable=d["reconstructable"], pt=d["pt"], eta=d["eta"], **kw)
        ref = res[f"tiger_{name}"]
        for k in ("attractive", "repulsive", "coward", "noise"):
            assert float(losses[k]) == pytest.approx(float(ref[k]), rel=1e-9)
        assert int(extra["n_rep"]) == int(ref["n_rep"])


def test_hinge_known_answers(golden_losses):
    d = golden_losses["td1"]["data"]
    args = dict(x=d["x"], particle_id=d["particle_id"], batch=d["batch"], true_edge_index=d["true_edge_index"],
                pt=d["pt"], eta=d["eta"], reconstructable=d["reconstructable"])
    l, _ = L.hinge_loss(**args)
    assert _f(l) == pytest.approx(golden_losses["known_answers"]["td1_hinge"], rel=1e-6)
    l, _ = L.hinge_loss(**args, rep_normalization="n_rep_edges")
    assert _f(l) == pytest.approx(golden_losses["known_answers"]["td1_hinge_n_rep_edges"], rel=1e-6)


@pytest.mark.parametrize("td", ["td1", "td2"])
def test_hinge_vs_reference_outputs(td, golden_losses):
    d = golden_losses[td]["data"]
    res = golden_losses[td]["results"]
    args = dict(x=d["x"], particle_id=d["particle_id"], batch=d["batch"], true_edge_index=d["true_edge_index"],
                pt=d["pt"], eta=d["eta"], reconstructable=d["reconstructable"])
    for name, kw in (("default", {}), ("n_rep_edges", dict(rep_normalization="n_rep_edges")),
                     ("n_att_edges_p2", dict(rep_normalization="n_att_edges", p_attr=2.0, p_rep=2.0, r_emb=0.5)),
                     ("all_hits", dict(rep_oi_only=False))):
        l, extra = L.hinge_loss(**args, **kw)
        ref = res[f"hinge_{name}"]
        assert float(l["attractive"]) == pytest.approx(float(ref["attractive"]), rel=1e-9)
        assert float(l["repulsive"]) == pytest.approx(float(ref["repulsive"]), rel=1e-9)
        for k in ("n_hits_oi", "n_edges_att", "n_edges_rep"):
            assert int(extra[k]) == int(ref[k])


def test_first_occurrences():
    """reference tests/test_losses.py:131-134."""
    assert L.first_occurrences(torch.tensor([0, 0, 1, 1, 2, 2])).tolist() == [0, 2, 4]


def test_focal_vs_bce_identity():
    """reference tests/test_losses.py:152-161."""
    g = torch.Generator().manual_seed(3)
    w = torch.rand(10, generator=g)
    y = (torch.rand(10, generator=g) > 0.5).float()
    assert float(L.focal_mean(w, y, alpha=0.5, gamma=0.0)) == pytest.approx(0.5 * float(L.bce_mean(w, y)), rel=1e-6)
    assert float(L.bce_mean(w, y)) == pytest.approx(float(torch.nn.functional.binary_cross_entropy(w, y)), rel=1e-6)


def test_ec_losses_vs_reference_outputs(golden_losses, golden_graphs, golden_models):
    gd = golden_graphs["sector0"]
    for wname, case in (("w_default", "ec_default_h64_sector0"), ("w_wide", "ec_wide64_sector0")):
        w = golden_models[case]["outputs"]["W"]
        ref = golden_losses["ec_losses"][wname]
        a = dict(w=w, y=gd["y"], edge_index=gd["edge_index"], pt=gd["pt"])
        got = {
            "bce": L.edge_weight_bce(**a), "bce_pt0.9": L.edge_weight_bce(**a, pt_thld=0.9),
            "focal": L.edge_weight_focal(**a),
            "focal_a0.4_g1.5_pt0.5": L.edge_weight_focal(**a, alpha=0.4, gamma=1.5, pt_thld=0.5),
            "focal_pw": L.edge_weight_focal(**a, pos_weight=torch.tensor([2.5])),
            "haughty": L.haughty_focal(**a), "haughty_pt0.9": L.haughty_focal(**a, pt_thld=0.9),
        }
        for k, v in got.items():
            assert float(v) == pytest.approx(float(ref[k]), rel=2e-6), k


def test_tiger_on_tcn_outputs(golden_losses, golden_graphs, golden_models):
    for cname in ("tcn_default_sector0", "tcn_orphans_sector1"):
        c = golden_models[cname]
        gd = golden_graphs[c["graph"]]
        o = c["outputs"]
        l, extra = L.condensation_tiger_loss(beta=o["B"], x=o["H"], particle_id=gd["particle_id"],
                                             reconstructable=gd["reconstructable"], pt=gd["pt"], eta=gd["eta"],
                                             ec_hit_mask=o["ec_hit_mask"], pt_thld=0.5)
        ref = golden_losses[f"tiger_{cname}"]
        for k in ("attractive", "repulsive", "coward", "noise"):
            a, b = float(l[k]), float(ref[k])
            assert (a != a and b != b) or a == pytest.approx(b, rel=1e-5), k
        assert int(extra["n_rep"]) == int(ref["n_rep"])
