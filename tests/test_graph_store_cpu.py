"""On-disk graph format (gnn_tracking_b200/graph_store.py): round trip on the host and the stored
plan against the integer oracle of the plan (bit-exact)."""
import torch

from gnn_tracking_b200 import graph_store as gs
from oracle import in_oracle as O


def _graph(n=500, e=4000, seed=0):
    gen = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, e), generator=gen)
    return dict(x=torch.randn(n, 14, generator=gen), edge_index=ei, edge_attr=torch.randn(e, 4, generator=gen),
                extras={"y": torch.rand(e, generator=gen) < 0.3, "particle_id": torch.randint(0, 50, (n,), generator=gen),
                        "pt": torch.rand(n, generator=gen).double()})


def test_round_trip_and_plan(tmp_path):
    g = _graph()
    path = tmp_path / "g.gtb"
    gs.write_graph(path, **g)
    d = gs.read_graph(path, device="cpu")
    assert torch.equal(d.x, g["x"]) and torch.equal(d.edge_index, g["edge_index"]) and torch.equal(d.edge_attr, g["edge_attr"])
    for k, v in g["extras"].items():
        got = getattr(d, k)
        assert got.dtype == v.dtype and torch.equal(got, v), k
    assert d.num_nodes == 500 and d.num_edges == 4000
    perm, rowptr, src, dst = O.plan(g["edge_index"], 500)
    pa = d._plan_arrays
    assert torch.equal(pa["plan.perm"].long(), perm.long()) and torch.equal(pa["plan.rowptr"].long(), rowptr.long())
    assert torch.equal(pa["plan.src_sorted"].long(), src.long()) and torch.equal(pa["plan.dst_sorted"].long(), dst.long())


def test_empty_graph_and_bad_files(tmp_path):
    import pytest
    path = tmp_path / "e.gtb"
    gs.write_graph(path, x=torch.zeros(3, 2), edge_index=torch.zeros((2, 0), dtype=torch.int64), edge_attr=torch.zeros(0, 1))
    d = gs.read_graph(path, device="cpu")
    assert d.num_edges == 0 and d._plan_arrays["plan.rowptr"].tolist() == [0, 0, 0, 0]
    with pytest.raises(IndexError):
        gs.write_graph(path, x=torch.zeros(3, 2), edge_index=torch.tensor([[0], [7]]), edge_attr=torch.zeros(1, 1))
    bad = tmp_path / "bad.gtb"
    bad.write_bytes(b"not a graph file at all")
    with pytest.raises(ValueError):
        gs.read_graph(bad, device="cpu")


def test_graph_loader_prefetch(tmp_path):
    import pytest
    paths = []
    for i in range(5):
        g = _graph(n=100 + i, e=700, seed=i)
        paths.append(tmp_path / f"g{i}.gtb")
        gs.write_graph(paths[-1], **g)
    loader = gs.GraphLoader(paths, device="cpu", prefetch=2)
    assert len(loader) == 5
    assert [d.num_nodes for d in loader] == [100, 101, 102, 103, 104]
    for i, d in enumerate(loader):   # early exit must not leave the reader thread blocked
        if i == 1:
            break
    with pytest.raises(FileNotFoundError):
        list(gs.GraphLoader([paths[0], tmp_path / "missing.gtb"], device="cpu"))


def test_shard_paths_follows_distributed_sampler():
    """One process per GPU: rank r walks entries r, r + W, ... of the list padded by wrap-around, the order of
    torch's ``DistributedSampler(shuffle=False)`` (what Lightning's DDP strategy gives the reference's loaders)."""
    import pytest
    from torch.utils.data import DistributedSampler
    for n_files in (1, 5, 8, 9):
        files = [f"g{i}" for i in range(n_files)]
        for world in (1, 2, 3, 8):
            shares = [gs.shard_paths(files, r, world) for r in range(world)]
            assert len({len(s) for s in shares}) == 1                      # every rank takes the same number of steps
            assert set().union(*map(set, shares)) == set(files)            # every file is visited
            for r in range(world):
                want = list(DistributedSampler(files, num_replicas=world, rank=r, shuffle=False))
                assert shares[r] == [files[i] for i in want]
    assert gs.shard_paths([], 1, 2) == []
    with pytest.raises(ValueError):
        gs.shard_paths(["a"], 2, 2)


def test_graph_loader_rank_share(tmp_path):
    paths = []
    for i in range(5):
        paths.append(tmp_path / f"g{i}.gtb")
        gs.write_graph(paths[-1], **_graph(n=100 + i, e=300, seed=i))
    seen = [[d.num_nodes for d in gs.GraphLoader(paths, device="cpu", rank=r, world_size=2)] for r in range(2)]
    assert seen == [[100, 102, 104], [101, 103, 100]]
    assert len(gs.GraphLoader(paths, device="cpu", rank=1, world_size=2)) == 3
    # without an initialised process group: one process, every file
    assert len(gs.GraphLoader(paths, device="cpu")) == 5
    # no GPU here: the topology is unknown, nothing is bound
    assert gs.gpu_local_cpus(0) == set() or isinstance(gs.gpu_local_cpus(0), set)
