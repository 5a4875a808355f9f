#!/usr/bin/env python
"""Generate the committed golden vectors by running the reference's OWN classes
(``/root/reference/src/gnn_tracking``, imported through ``oracle/shims.py``).

Run in the authoring container only (``/root/reference`` does not exist on the GPU
box):  ``python tests/golden/make_golden.py``.  Writes

* ``graphs.pt``     -- fixture graphs: the reference's bundled ``test_graph.pt``,
                      the two conftest "2-sector" graphs (tests/conftest.py:42-70),
                      one seeded synthetic graph with heavy / isolated nodes.
* ``models.pt``     -- per case: ctor kwargs, reference state_dict, reference outputs.
* ``losses.pt``     -- td1 / td2 of reference tests/test_losses.py:46-76 (seed 0), the
                      known-answer dicts of :112-123 / :194-203 and the reference's
                      outputs for every loss on the path.
"""
from __future__ import annotations

import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
os.environ["TORCHDYNAMO_DISABLE"] = "1"

from oracle import reference_loader as rl  # noqa: E402
from tests.golden.common import widen  # noqa: E402

rl.load()
HERE = Path(__file__).resolve().parent
REF_TESTS = Path("/root/reference/tests")

GRAPH_KEYS = ("x", "edge_index", "edge_attr", "y", "particle_id", "pt", "eta", "reconstructable", "layer")


def graph_to_dict(g):
    return {k: getattr(g, k).clone() for k in GRAPH_KEYS if hasattr(g, k)}


def build_graphs():
    from gnn_tracking.graph_construction.graph_builder import GraphBuilder
    from gnn_tracking.preprocessing.point_cloud_builder import PointCloudBuilder

    graphs = {}
    g = torch.load(REF_TESTS / "test_data/graphs/test_graph.pt", weights_only=False)
    graphs["test_graph"] = graph_to_dict(g)

    tdir = REF_TESTS / "test_data/trackml"
    pc, gr = Path(tempfile.mkdtemp()), Path(tempfile.mkdtemp())
    PointCloudBuilder(indir=tdir, outdir=str(pc), n_sectors=2, pixel_only=True, redo=False,
                      measurement_mode=True, thld=0.9, detector_config=tdir / "detectors.csv.gz",
                      add_true_edges=True).process()
    GraphBuilder(str(pc), str(gr), redo=True, measurement_mode=True).process(stop=None)
    for i, f in enumerate(sorted(os.listdir(gr))):
        graphs[f"sector{i}"] = graph_to_dict(torch.load(gr / f, weights_only=False))

    # seeded synthetic graph: heavy nodes (degree > 300), isolated nodes, duplicate edges, a self loop
    gen = torch.Generator().manual_seed(1234)
    n, e = 1000, 6000
    src = torch.randint(0, n - 100, (e,), generator=gen)
    dst = torch.randint(0, n - 100, (e,), generator=gen)
    dst[:700] = 7          # one very heavy destination (spans many 128-edge tiles)
    dst[700:1000] = 8
    src[1000:1010] = 3
    dst[1000:1010] = 4      # duplicate edges
    src[1010] = dst[1010] = 11  # self loop
    graphs["synthetic"] = {
        "x": torch.randn(n, 14, generator=gen),
        "edge_index": torch.stack([src, dst]),
        "edge_attr": torch.randn(e, 4, generator=gen),
        "y": (torch.rand(e, generator=gen) < 0.3).float(),
        "particle_id": torch.randint(0, 150, (n,), generator=gen),
        "pt": torch.rand(n, generator=gen) * 3,
        "eta": (torch.rand(n, generator=gen) - 0.5) * 9,
        "reconstructable": (torch.rand(n, generator=gen) < 0.9).long(),
        "layer": torch.randint(0, 18, (n,), generator=gen),
    }
    return graphs


def as_data(gd, **over):
    from torch_geometric.data import Data

    d = {k: v.clone() for k, v in gd.items()}
    d.update(over)
    return Data(**d)


def scale_weights(model, factor):
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("weight"):
                p.mul_(factor)


def sd_of(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def build_models(graphs):
    from gnn_tracking.models.edge_classifier import ECForGraphTCN
    from gnn_tracking.models.interaction_network import InteractionNetwork
    from gnn_tracking.models.resin import ResIN
    from gnn_tracking.models.track_condensation_networks import GraphTCN

    cases = {}

    # ---- single IN layers (BASELINE config 1 and friends)
    in_cases = {
        "in_testgraph_default": ("test_graph", None, dict(node_indim=14, edge_indim=14, node_outdim=5, edge_outdim=4, node_hidden_dim=40, edge_hidden_dim=40), 1.0),
        "in_sector0_raw": ("sector0", None, dict(node_indim=14, edge_indim=4, node_outdim=5, edge_outdim=4, node_hidden_dim=40, edge_hidden_dim=40), 1.0),
        "in_sector0_wide64": ("sector0", (64, 64), dict(node_indim=64, edge_indim=64, node_outdim=64, edge_outdim=64, node_hidden_dim=64, edge_hidden_dim=64), 1.0),
        "in_synthetic_wide64_x3": ("synthetic", (64, 64), dict(node_indim=64, edge_indim=64, node_outdim=64, edge_outdim=64, node_hidden_dim=64, edge_hidden_dim=64), 3.0),
        "in_synthetic_odd": ("synthetic", (7, 3), dict(node_indim=7, edge_indim=3, node_outdim=6, edge_outdim=9, node_hidden_dim=33, edge_hidden_dim=21), 2.0),
        "in_testgraph_wide128": ("test_graph", (128, 128), dict(node_indim=128, edge_indim=128, node_outdim=128, edge_outdim=128, node_hidden_dim=128, edge_hidden_dim=128), 1.0),
    }
    for name, (gname, wide, kw, scale) in in_cases.items():
        gd = graphs[gname] if wide is None else widen(graphs[gname], *wide, seed=7)
        torch.manual_seed(0)
        m = InteractionNetwork(**kw)
        scale_weights(m, scale)
        with torch.no_grad():
            xt, et = m(gd["x"], gd["edge_index"], gd["edge_attr"])
        cases[name] = {"kind": "in", "graph": gname, "kwargs": kw, "state_dict": sd_of(m),
                       "widen": ((*wide, 7) if wide else None), "input_checksum": float(gd["x"].double().sum()),
                       "outputs": {"x_tilde": xt, "e_tilde": et}}

    # ---- ResIN stacks
    resin_cases = {
        "resin_skip1_sector0": ("sector0", (5, 4), dict(node_dim=5, edge_dim=4, object_hidden_dim=32, relational_hidden_dim=48, alpha=0.5, n_layers=3, residual_type="skip1", residual_kwargs={"collect_hidden_edge_embeds": True}), 2.0),
        "resin_skip1_alpha0": ("sector1", (8, 8), dict(node_dim=8, edge_dim=8, object_hidden_dim=16, relational_hidden_dim=16, alpha=0.0, n_layers=2, residual_type="skip1"), 2.0),
        "resin_skip2_sector0": ("sector0", (5, 4), dict(node_dim=5, edge_dim=4, object_hidden_dim=24, relational_hidden_dim=24, alpha=0.5, n_layers=2, residual_type="skip2", residual_kwargs={"collect_hidden_edge_embeds": True}), 2.0),
        "resin_skip2_L4": ("test_graph", (6, 5), dict(node_dim=6, edge_dim=5, object_hidden_dim=24, relational_hidden_dim=24, alpha=0.3, n_layers=4, residual_type="skip2", residual_kwargs={"collect_hidden_edge_embeds": True}), 2.0),
        "resin_skiptop_sector1": ("sector1", (5, 4), dict(node_dim=5, edge_dim=4, object_hidden_dim=24, relational_hidden_dim=24, alpha=0.5, n_layers=3, residual_type="skip_top", residual_kwargs={"collect_hidden_edge_embeds": True, "connect_to": 1}), 2.0),
    }
    for name, (gname, wide, kw, scale) in resin_cases.items():
        gd = widen(graphs[gname], *wide, seed=11)
        torch.manual_seed(0)
        kw_c = {k: (dict(v) if isinstance(v, dict) else v) for k, v in kw.items()}
        m = ResIN(**kw_c)
        scale_weights(m, scale)
        with torch.no_grad():
            x, e, es = m(gd["x"], gd["edge_index"], gd["edge_attr"])
        cases[name] = {"kind": "resin", "graph": gname, "kwargs": kw, "state_dict": sd_of(m),
                       "widen": (*wide, 11), "input_checksum": float(gd["x"].double().sum()),
                       "outputs": {"x": x, "edge_attr": e, "edge_attrs": es}}

    # ---- Edge classifiers
    ec_cases = {
        "ec_yml_testgraph": ("test_graph", dict(node_indim=14, edge_indim=14, L_ec=1), 1.0),
        "ec_default_h64_sector0": ("sector0", dict(node_indim=14, edge_indim=4, hidden_dim=64, L_ec=3), 1.0),
        "ec_default_h64_sector1_x3": ("sector1", dict(node_indim=14, edge_indim=4, hidden_dim=64, L_ec=3), 3.0),
        "ec_wide64_sector0": ("sector0", dict(node_indim=14, edge_indim=4, interaction_node_dim=64, interaction_edge_dim=64, hidden_dim=64, L_ec=3), 1.0),
        "ec_wide64_synthetic_x2": ("synthetic", dict(node_indim=14, edge_indim=4, interaction_node_dim=64, interaction_edge_dim=64, hidden_dim=64, L_ec=3), 2.0),
        "ec_no_intermediate": ("sector0", dict(node_indim=14, edge_indim=4, hidden_dim=32, L_ec=2, use_intermediate_edge_embeddings=False), 2.0),
        "ec_no_node_emb": ("sector1", dict(node_indim=14, edge_indim=4, hidden_dim=32, L_ec=2, use_node_embedding=False), 2.0),
        "ec_skip2": ("sector0", dict(node_indim=14, edge_indim=4, hidden_dim=16, L_ec=2, residual_type="skip2"), 2.0),
        "ec_skiptop": ("sector1", dict(node_indim=14, edge_indim=4, hidden_dim=16, L_ec=3, residual_type="skip_top", residual_kwargs={"connect_to": 1}), 2.0),
    }
    for name, (gname, kw, scale) in ec_cases.items():
        gd = graphs[gname]
        torch.manual_seed(0)
        kw_c = {k: (dict(v) if isinstance(v, dict) else v) for k, v in kw.items()}
        m = ECForGraphTCN(**kw_c)
        scale_weights(m, scale)
        with torch.no_grad():
            out = m(as_data(gd))
        cases[name] = {"kind": "ec", "graph": gname, "kwargs": kw, "state_dict": sd_of(m),
                       "outputs": {k: v for k, v in out.items()}}

    # ---- GraphTCN (object condensation model)
    tcn_cases = {
        "tcn_default_sector0": ("sector0", dict(node_indim=14, edge_indim=4, hidden_dim=32, L_ec=2, L_hc=2), 2.0),
        "tcn_orphans_sector1": ("sector1", dict(node_indim=14, edge_indim=4, hidden_dim=32, L_ec=2, L_hc=3, mask_orphan_nodes=True, ec_threshold=0.4), 2.0),
        "tcn_feedw_ecemb_synthetic": ("synthetic", dict(node_indim=14, edge_indim=4, hidden_dim=24, L_ec=2, L_hc=2, feed_edge_weights=True, use_ec_embeddings_for_hc=True, ec_threshold=0.45), 2.0),
        "tcn_h128_L8_testgraph": ("test_graph", dict(node_indim=14, edge_indim=14, h_dim=128, e_dim=128, hidden_dim=128, L_ec=1, L_hc=2, ec_threshold=0.3), 1.0),
        "tcn_alpha_latent": ("sector0", dict(node_indim=14, edge_indim=4, hidden_dim=16, L_ec=1, L_hc=1, h_outdim=4, alpha_latent=0.5, n_embedding_coords=3), 2.0),
    }
    for name, (gname, kw, scale) in tcn_cases.items():
        gd = graphs[gname]
        torch.manual_seed(0)
        m = GraphTCN(**kw)
        scale_weights(m, scale)
        with torch.no_grad():
            out = m(as_data(gd))
        cases[name] = {"kind": "tcn", "graph": gname, "kwargs": kw, "state_dict": sd_of(m),
                       "outputs": {k: v.clone() for k, v in out.items() if v is not None}}
    return cases


def build_losses(graphs, models):
    import importlib.util

    # the reference's own test module (td1, td2 and the known-answer dicts)
    spec = importlib.util.spec_from_file_location("_ref_test_losses", REF_TESTS / "test_losses.py")
    tl = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tl)
    from gnn_tracking.metrics.losses.ec import EdgeWeightBCELoss, EdgeWeightFocalLoss, HaughtyFocalLoss
    from gnn_tracking.metrics.losses.metric_learning import GraphConstructionHingeEmbeddingLoss
    from gnn_tracking.metrics.losses.oc import CondensationLossRG, CondensationLossTiger

    out = {"known_answers": {
        "td1_condensation": dict(tl._td1_c_losses), "td2_condensation": dict(tl._td2_c_losses),
        "td1_hinge": {"attractive": 0.7307405975481213, "repulsive": 11.076146539572338},
        "td1_hinge_n_rep_edges": {"attractive": 0.7307405975481213, "repulsive": 0.34612957938781874},
    }}
    for name, td in (("td1", tl.td1), ("td2", tl.td2)):
        d = {k: getattr(td, k).clone() for k in ("beta", "x", "particle_id", "pt", "eta", "reconstructable", "batch", "true_edge_index")}
        res = {}
        for strat, cls in (("tiger", CondensationLossTiger), ("rg", CondensationLossRG)):
            for kw_name, kw in (("default", {}), ("alt", dict(q_min=0.1, pt_thld=0.3, max_eta=3.5))):
                try:
                    r = cls(**kw)(beta=td.beta, x=td.x, particle_id=td.particle_id, reconstructable=td.reconstructable, pt=td.pt, eta=td.eta)
                except AssertionError:  # "No hits left after masking" for this cut on td1
                    continue
                res[f"{strat}_{kw_name}"] = {k: v.detach().clone() for k, v in r.loss_dct.items()}
                if strat == "tiger":
                    res[f"{strat}_{kw_name}"]["n_rep"] = torch.as_tensor(r.extra_metrics["n_rep"])
        for kw_name, kw in (("default", {}), ("n_rep_edges", dict(rep_normalization="n_rep_edges")),
                            ("n_att_edges_p2", dict(rep_normalization="n_att_edges", p_attr=2.0, p_rep=2.0, r_emb=0.5)),
                            ("all_hits", dict(rep_oi_only=False))):
            r = GraphConstructionHingeEmbeddingLoss(**kw)(x=td.x, particle_id=td.particle_id, reconstructable=td.reconstructable, pt=td.pt, eta=td.eta, batch=td.batch, true_edge_index=td.true_edge_index)
            res[f"hinge_{kw_name}"] = {**{k: v.detach().clone() for k, v in r.loss_dct.items()},
                                       **{k: torch.as_tensor(v) for k, v in r.extra_metrics.items()}}
        out[name] = {"data": d, "results": res}

    # EC losses on the reference EC's own W for sector0 (fp32)
    gd = graphs["sector0"]
    w = models["ec_default_h64_sector0"]["outputs"]["W"]
    w3 = models["ec_wide64_sector0"]["outputs"]["W"]
    ecl = {}
    for wname, ww in (("w_default", w), ("w_wide", w3)):
        args = dict(w=ww, y=gd["y"], edge_index=gd["edge_index"], pt=gd["pt"])
        ecl[wname] = {
            "bce": EdgeWeightBCELoss()(**args), "bce_pt0.9": EdgeWeightBCELoss(pt_thld=0.9)(**args),
            "focal": EdgeWeightFocalLoss()(**args), "focal_a0.4_g1.5_pt0.5": EdgeWeightFocalLoss(alpha=0.4, gamma=1.5, pt_thld=0.5)(**args),
            "focal_pw": EdgeWeightFocalLoss(pos_weight=torch.tensor([2.5]))(**args),
            "haughty": HaughtyFocalLoss()(**args), "haughty_pt0.9": HaughtyFocalLoss(pt_thld=0.9)(**args),
        }
    out["ec_losses"] = {k: {kk: vv.detach().clone() for kk, vv in v.items()} for k, v in ecl.items()}

    # condensation loss on a reference GraphTCN output (fp32, with ec_hit_mask)
    for cname in ("tcn_default_sector0", "tcn_orphans_sector1"):
        c = models[cname]
        gd = graphs[c["graph"]]
        o = c["outputs"]
        r = CondensationLossTiger(pt_thld=0.5)(beta=o["B"], x=o["H"], particle_id=gd["particle_id"], reconstructable=gd["reconstructable"], pt=gd["pt"], eta=gd["eta"], ec_hit_mask=o["ec_hit_mask"])
        out[f"tiger_{cname}"] = {**{k: v.detach().clone() for k, v in r.loss_dct.items()}, "n_rep": torch.as_tensor(r.extra_metrics["n_rep"])}
    return out


def main():
    torch.set_num_threads(1)  # deterministic summation order for the pins
    graphs = build_graphs()
    torch.save(graphs, HERE / "graphs.pt")
    models = build_models(graphs)
    torch.save(models, HERE / "models.pt")
    losses = build_losses(graphs, models)
    torch.save(losses, HERE / "losses.pt")
    for f in ("graphs.pt", "models.pt", "losses.pt"):
        print(f, os.path.getsize(HERE / f) // 1024, "KiB")
    np.set_printoptions(precision=9)
    print("self-check IN(test_graph) x_tilde.sum =", float(models["in_testgraph_default"]["outputs"]["x_tilde"].sum()))
    print("self-check EC(ec.yml) W.sum =", float(models["ec_yml_testgraph"]["outputs"]["W"].sum()))


if __name__ == "__main__":
    main()
