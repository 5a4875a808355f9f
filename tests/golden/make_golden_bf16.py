#!/usr/bin/env python
"""Golden outputs of the reference's OWN ``GraphTCN`` under ``torch.autocast(bfloat16)`` (BASELINE config 3:
hidden = node = edge width 128, 8 condenser layers), generated in the authoring container only:

    python tests/golden/make_golden_bf16.py        # writes tests/golden/bf16_tcn.pt

The weights are not stored (1.9 M parameters): the model is built under ``torch.manual_seed(0)``, which gives
the reference's initialisation with either package (tests/test_boundary_cpu.py pins that); a checksum of
the reference parameters is stored so that a drift of the RNG stream fails loudly instead of silently
comparing different networks.  The threshold between the edge classifier and the condenser is 0 (every
edge kept): a bf16 ulp on W must not decide which edges exist.
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
from tests.golden.common import load  # noqa: E402

KW = dict(node_indim=14, edge_indim=4, h_dim=128, e_dim=128, hidden_dim=128, L_ec=3, L_hc=8, h_outdim=2, ec_threshold=0.0)


def checksum(sd) -> float:
    return float(sum(v.double().abs().sum() for v in sd.values()))


def main():
    rl.load()
    from gnn_tracking.models.track_condensation_networks import GraphTCN
    from torch_geometric.data import Data
    graphs = load("graphs")
    out = {"kwargs": KW, "seed": 0, "cases": {}}
    for gname in ("sector0", "synthetic"):
        gd = graphs[gname]
        torch.manual_seed(0)
        m = GraphTCN(**KW)
        out["param_checksum"] = checksum(m.state_dict())
        data = Data(**{k: v.clone() for k, v in gd.items()})
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            o = m(data)
        with torch.no_grad():  # the same network in fp32: how far bf16 moves the reference itself
            o32 = m(Data(**{k: v.clone() for k, v in gd.items()}))
        res = {k: o[k].float().clone() for k in ("W", "H", "B")}
        res32 = {k: o32[k].float().clone() for k in ("W", "H", "B")}
        out["cases"][gname] = {"outputs": res, "outputs_fp32": res32,
                               "dtypes": {k: str(o[k].dtype) for k in ("W", "H", "B")}}
        for k in res:
            print(gname, k, o[k].dtype, "max|bf16 - fp32| =", float((res[k] - res32[k]).abs().max()),
                  "max|ref| =", float(res32[k].abs().max()))
    torch.save(out, Path(__file__).resolve().parent / "bf16_tcn.pt")


if __name__ == "__main__":
    main()
