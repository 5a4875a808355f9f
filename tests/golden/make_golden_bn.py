#!/usr/bin/env python
"""Golden outputs of the reference's OWN ``ResIN(residual_type="skip2", residual_kwargs={"add_bn": True})``
(models/resin.py:117-175) in training mode (batch statistics) and in eval mode (running statistics), generated
in the authoring container only:

    python tests/golden/make_golden_bn.py        # writes tests/golden/resin_bn.pt
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
os.environ["TORCHDYNAMO_DISABLE"] = "1"

from oracle import reference_loader as rl  # noqa: E402
from tests.golden.common import load, widen  # noqa: E402

CASES = {
    "skip2_bn_narrow": ("sector0", (6, 5), dict(node_dim=6, edge_dim=5, object_hidden_dim=24, relational_hidden_dim=24, alpha=0.4,
                                               n_layers=2, residual_type="skip2",
                                               residual_kwargs={"add_bn": True, "collect_hidden_edge_embeds": True})),
    "skip2_bn_wide64_L4": ("sector1", (64, 64), dict(node_dim=64, edge_dim=64, object_hidden_dim=64, relational_hidden_dim=64,
                                                      alpha=0.5, n_layers=4, residual_type="skip2",
                                                      residual_kwargs={"add_bn": True, "collect_hidden_edge_embeds": True})),
}


def main():
    rl.load()
    from gnn_tracking.models.resin import ResIN
    graphs = load("graphs")
    out = {}
    for name, (gname, wide, kw) in CASES.items():
        gd = widen(graphs[gname], *wide, seed=13)
        torch.manual_seed(0)
        m = ResIN(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in kw.items()})
        gen = torch.Generator().manual_seed(5)
        with torch.no_grad():  # non-trivial affine parameters and running statistics
            for mod in m.modules():
                if isinstance(mod, torch.nn.BatchNorm1d):
                    mod.weight.copy_(1 + 0.3 * torch.randn(mod.weight.shape, generator=gen))
                    mod.bias.copy_(0.2 * torch.randn(mod.bias.shape, generator=gen))
                    mod.running_mean.copy_(0.1 * torch.randn(mod.running_mean.shape, generator=gen))
                    mod.running_var.copy_(0.5 + torch.rand(mod.running_var.shape, generator=gen))
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        res = {}
        for mode in ("train", "eval"):
            m.load_state_dict(sd)
            m.train(mode == "train")
            with torch.no_grad():
                x, e, es = m(gd["x"], gd["edge_index"], gd["edge_attr"])
            res[mode] = {"x": x.clone(), "edge_attr": e.clone()}
            if e.size(1) <= 8:  # the narrow case also keeps the per-block edge embeddings (the wide ones are megabytes)
                res[mode]["edge_attrs"] = [t.clone() for t in es]
        out[name] = {"graph": gname, "widen": (*wide, 13), "kwargs": kw, "state_dict": sd, "outputs": res,
                     "input_checksum": float(gd["x"].double().sum())}
        print(name, {k: float(v["x"].abs().max()) for k, v in res.items()})
    torch.save(out, Path(__file__).resolve().parent / "resin_bn.pt")


if __name__ == "__main__":
    main()
