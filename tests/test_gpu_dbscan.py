"""GPU DBSCAN (gtb_dbscan_f32) against the reference's DBSCANFastRescan labels (committed goldens,
tests/golden/make_golden_dbscan.py) and against sklearn run live: labels must be IDENTICAL, numbering
included (integer work: bit-exact)."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).parent / "golden" / "dbscan.pt"


def test_dbscan_matches_reference_goldens():
    from gnn_tracking_b200.postprocessing.dbscan import DBSCANFastRescan

    g = torch.load(GOLDEN)
    scanners = [DBSCANFastRescan(x.cuda(), max_eps=1.0) for x in g["datasets"]]
    for case in g["cases"]:
        got = scanners[case["dataset"]].cluster(eps=case["eps"], min_pts=case["min_pts"]).cpu()
        assert torch.equal(got, case["labels"].long()), (case["dataset"], case["eps"], case["min_pts"])


@pytest.mark.parametrize("n,d,eps,min_samples", [(1, 2, 0.5, 1), (1, 2, 0.5, 2), (257, 3, 0.3, 2), (6000, 2, 0.04, 3),
                                                 (20000, 8, 0.9, 4), (3000, 16, 2.0, 2)])
def test_dbscan_matches_sklearn(n, d, eps, min_samples):
    from sklearn.cluster import DBSCAN

    from gnn_tracking_b200.postprocessing.dbscan import dbscan

    rng = np.random.default_rng(n + d)
    centres = rng.uniform(-2, 2, size=(max(n // 12, 1), d))
    x = (centres[rng.integers(0, len(centres), n)] + 0.1 * rng.standard_normal((n, d))).astype(np.float32)
    want = DBSCAN(eps=eps, min_samples=min_samples).fit_predict(x)
    got = dbscan(torch.from_numpy(x).cuda(), eps, min_samples).cpu().numpy()
    assert np.array_equal(got, want)


def test_dbscan_duplicates_and_empty():
    from sklearn.cluster import DBSCAN

    from gnn_tracking_b200.postprocessing.dbscan import dbscan

    assert dbscan(torch.zeros((0, 3), device="cuda"), 0.1, 1).numel() == 0
    x = torch.tensor([[0.0, 0.0]] * 5 + [[1.0, 1.0]] * 2 + [[5.0, 5.0]], device="cuda")
    want = DBSCAN(eps=0.1, min_samples=2).fit_predict(x.cpu().numpy())
    assert np.array_equal(dbscan(x, 0.1, 2).cpu().numpy(), want)
    with pytest.raises(RuntimeError):
        dbscan(torch.zeros((4, 17), device="cuda"), 0.1, 1)


def test_dbscan_full_size_properties():
    """100k points (BASELINE config 5 scale): properties that do not need the CPU run."""
    from gnn_tracking_b200.postprocessing.dbscan import dbscan

    g = torch.Generator(device="cuda").manual_seed(0)
    n = 100_000
    centres = torch.rand((n // 10, 8), device="cuda", generator=g) * 20
    x = centres[torch.randint(0, n // 10, (n,), device="cuda", generator=g)] + 0.02 * torch.randn((n, 8), device="cuda", generator=g)
    a = dbscan(x, 0.2, 1)
    assert int(a.min()) == 0  # min_samples=1: no noise
    # cluster ids are dense and ordered by first appearance
    first = torch.full((int(a.max()) + 1,), n, dtype=torch.int64, device="cuda").scatter_reduce(0, a, torch.arange(n, device="cuda"), "amin")
    assert bool((first[1:] > first[:-1]).all())
    # permutation invariance of the partition (labels up to renumbering)
    perm = torch.randperm(n, device="cuda", generator=g)
    b = dbscan(x[perm], 0.2, 1)
    pairs = torch.unique(torch.stack([a[perm], b]), dim=1)
    assert pairs.size(1) == int(a.max()) + 1 == int(b.max()) + 1


def test_fastrescan_equal_slowrescan():
    """The reference's own pin (tests/test_fastrescanner.py:7-14) with the GPU class in place of the
    sklearn-backed one."""
    from sklearn.cluster import DBSCAN

    from gnn_tracking_b200.postprocessing.dbscan import DBSCANFastRescan

    x = np.random.default_rng(0).uniform(size=(100, 2)).astype(np.float32)
    fr = DBSCANFastRescan(x, max_eps=0.15)  # numpy in, numpy out: the way the reference's scanner calls it
    for eps in [0.1, 0.05]:
        for min_pts in [1, 2]:
            labels = fr.cluster(eps=eps, min_pts=min_pts)
            assert isinstance(labels, np.ndarray)
            labels2 = DBSCAN(eps=eps, min_samples=min_pts).fit_predict(x)
            assert (labels == labels2).all()


def test_graph_tcn_latent_space_clustering():
    """BASELINE config 5 in small: GraphTCN forward on the CUDA path, then the DBSCAN trials of the
    hyper-parameter scan on the device-resident latent coordinates H; labels equal sklearn's on the
    same H (dbscanscanner.py:160-177 copies H to the host first)."""
    from sklearn.cluster import DBSCAN

    from gnn_tracking_b200.models.track_condensation_networks import GraphTCN
    from gnn_tracking_b200.postprocessing.dbscan import DBSCANFastRescan

    gen = torch.Generator().manual_seed(4)
    n, e = 4000, 50000
    data = type("D", (), {})()
    data.x = torch.randn(n, 14, generator=gen).cuda()
    data.edge_index = torch.randint(0, n, (2, e), generator=gen).cuda()
    data.edge_attr = torch.randn(e, 4, generator=gen).cuda()
    torch.manual_seed(2)
    m = GraphTCN(14, 4, hidden_dim=32, L_ec=2, L_hc=2, h_outdim=3, ec_threshold=0.0).cuda()
    with torch.no_grad():
        h = m(data)["H"]
    assert h.shape == (n, 3)
    spread = float(h.std())
    scanner = DBSCANFastRescan(h, max_eps=1.0)
    hc = h.cpu().numpy()
    for eps, min_pts in [(0.05 * spread, 1), (0.1 * spread, 2), (0.3 * spread, 4)]:
        got = scanner.cluster(eps=eps, min_pts=min_pts).cpu().numpy()
        want = DBSCAN(eps=eps, min_samples=min_pts).fit_predict(hc)
        assert np.array_equal(got, want), (eps, min_pts)


@pytest.mark.parametrize("n,d,eps,min_samples", [(5000, 1, 0.01, 2), (8000, 2, 0.03, 3), (8000, 3, 0.1, 2), (6000, 5, 0.4, 3),
                                                 (4000, 8, 0.9, 4), (3000, 3, 50.0, 2), (3000, 3, 1e-4, 1), (100000, 3, 0.05, 2)])
def test_dbscan_grid_equals_brute_force(n, d, eps, min_samples):
    """The cell-list search (gtb_dbscan_grid_f32) against the all-pairs walk (gtb_dbscan_f32): identical labels,
    for every grid shape (1 to 3 binned coordinates, one giant cell, cells far smaller than the spacing)."""
    from gnn_tracking_b200.postprocessing.dbscan import dbscan

    rng = np.random.default_rng(n + 7 * d)
    centres = rng.uniform(-2, 2, size=(max(n // 12, 1), d))
    x = (centres[rng.integers(0, len(centres), n)] + 0.1 * rng.standard_normal((n, d))).astype(np.float32)
    x[: n // 50] = x[n // 50: 2 * (n // 50)]  # duplicates
    xt = torch.from_numpy(x).cuda()
    a = dbscan(xt, eps, min_samples, method="grid")
    b = dbscan(xt, eps, min_samples, method="brute")
    assert torch.equal(a, b)
