"""Helpers shared by ``make_golden.py`` (generator) and the parity tests (readers)."""
from __future__ import annotations

from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent


def widen(gd: dict, dn: int, de: int, seed: int) -> dict:
    """Replace x / edge_attr of a fixture graph by seeded N(0,1) features of width
    (dn, de) (CPU generator: reproducible on every box)."""
    gen = torch.Generator().manual_seed(seed)
    return {**gd, "x": torch.randn(gd["x"].size(0), dn, generator=gen),
            "edge_attr": torch.randn(gd["edge_index"].size(1), de, generator=gen)}


def load(name: str):
    return torch.load(HERE / f"{name}.pt", weights_only=True)


def case_inputs(case: dict, graphs: dict) -> dict:
    """Graph dict (x, edge_index, edge_attr, ...) a golden model case was run on."""
    gd = graphs[case["graph"]]
    if case.get("widen"):
        dn, de, seed = case["widen"]
        gd = widen(gd, dn, de, seed)
        assert abs(float(gd["x"].double().sum()) - case["input_checksum"]) < 1e-6, "RNG drift: regenerate golden"
    return gd
