"""Generates tests/golden/dbscan.pt by running the REFERENCE DBSCANFastRescan
(/root/reference/src/gnn_tracking/postprocessing/fastrescanner.py) on seeded latent-space-like
points.  Runs only in the build container (the reference is not on the GPU box).

    python tests/golden/make_golden_dbscan.py
"""
import importlib.util
from pathlib import Path

import numpy as np
import torch

spec = importlib.util.spec_from_file_location(
    "fastrescanner", "/root/reference/src/gnn_tracking/postprocessing/fastrescanner.py")
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)


def points(seed: int, n_clusters: int, per: int, noise: int, d: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    centres = rng.uniform(-3, 3, size=(n_clusters, d))
    sizes = rng.integers(1, per + 1, size=n_clusters)
    x = np.concatenate([c + 0.08 * rng.standard_normal((s, d)) for c, s in zip(centres, sizes)]
                       + [rng.uniform(-3.5, 3.5, size=(noise, d))])
    return x[rng.permutation(len(x))].astype(np.float32)


datasets, cases = [], []
for seed, (nc, per, noise, d) in enumerate([(120, 12, 300, 2), (200, 10, 400, 3), (60, 30, 100, 8)]):
    x = points(seed, nc, per, noise, d)
    datasets.append(torch.from_numpy(x))
    scanner = mod.DBSCANFastRescan(x, max_eps=1.0)
    for eps, min_pts in [(0.05, 1), (0.12, 2), (0.2, 3), (0.35, 5), (1.0, 4)]:
        labels = scanner.cluster(eps=eps, min_pts=min_pts)
        cases.append({"dataset": seed, "eps": eps, "min_pts": min_pts,
                      "labels": torch.from_numpy(labels.astype(np.int16))})
        print(d, len(x), eps, min_pts, "clusters", labels.max() + 1, "noise", (labels < 0).sum())
torch.save({"datasets": datasets, "cases": cases}, Path(__file__).parent / "dbscan.pt")
