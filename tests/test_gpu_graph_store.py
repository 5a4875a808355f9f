"""graph_store.read_graph on the device: one pinned staging buffer, one copy, the stored plan adopted
(bit-identical to gtb_plan_build) and used by the models without sorting."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_read_graph_adopts_plan(tmp_path):
    from gnn_tracking_b200 import graph_store as gs
    from gnn_tracking_b200 import ops
    from gnn_tracking_b200.models.edge_classifier import ECForGraphTCN
    from gnn_tracking_b200.plan import build_plan, clear_plan_cache, get_plan

    gen = torch.Generator().manual_seed(3)
    n, e = 3000, 40000
    ei = torch.randint(0, n, (2, e), generator=gen)
    x, ea = torch.randn(n, 14, generator=gen), torch.randn(e, 4, generator=gen)
    path = tmp_path / "g.gtb"
    gs.write_graph(path, x=x, edge_index=ei, edge_attr=ea, extras={"y": torch.rand(e, generator=gen) < 0.5})
    clear_plan_cache()
    d = gs.read_graph(path)
    torch.cuda.synchronize()
    assert d.x.is_cuda and torch.equal(d.x.cpu(), x) and torch.equal(d.edge_index.cpu(), ei) and d.y.dtype == torch.bool
    before = ops.launch_count()
    plan = get_plan(d.edge_index, n)          # adopted: no kernels
    assert ops.launch_count() == before
    ref = build_plan(ei.cuda(), n)
    for k in ("perm", "rowptr", "src_sorted", "dst_sorted"):
        assert torch.equal(getattr(plan, k), getattr(ref, k)), k
    torch.manual_seed(0)
    m = ECForGraphTCN(node_indim=14, edge_indim=4, L_ec=2, hidden_dim=32).cuda()
    with torch.no_grad():
        w_stored = m(d)["W"]
        clear_plan_cache()
        w_built = m(gs.GraphData(x=x.cuda(), edge_index=ei.cuda(), edge_attr=ea.cuda()))["W"]
    # same kernels on the same plan; the per-destination sums use floating-point atomics, whose order varies
    assert torch.allclose(w_stored, w_built, rtol=0, atol=1e-6)


def test_prefetched_graphs_give_the_same_results(tmp_path):
    """DevicePrefetcher (pinned host graphs) and GraphLoader (files): the copy of the next graph runs on
    a side stream under the current forward; outputs equal those of plain resident inputs."""
    from gnn_tracking_b200 import graph_store as gs
    from gnn_tracking_b200.models.edge_classifier import ECForGraphTCN
    from gnn_tracking_b200.plan import clear_plan_cache

    torch.manual_seed(1)
    m = ECForGraphTCN(node_indim=14, edge_indim=4, L_ec=2, hidden_dim=32).cuda()
    graphs, paths, want = [], [], []
    for i in range(5):
        gen = torch.Generator().manual_seed(10 + i)
        n, e = 2000 + 100 * i, 30000 + 1000 * i
        g = dict(x=torch.randn(n, 14, generator=gen), edge_index=torch.randint(0, n, (2, e), generator=gen),
                 edge_attr=torch.randn(e, 4, generator=gen))
        graphs.append(gs.GraphData(**{k: v.pin_memory() for k, v in g.items()}))
        paths.append(tmp_path / f"g{i}.gtb")
        gs.write_graph(paths[-1], **g)
        with torch.no_grad():
            want.append(m(gs.GraphData(**{k: v.cuda() for k, v in g.items()}))["W"].clone())
    for source in (gs.DevicePrefetcher(graphs), gs.GraphLoader(paths, prefetch=2),
                   gs.GraphLoader(paths, prefetch=2, numa_local=True)):  # reader thread on the GPU's NUMA node
        clear_plan_cache()
        got = []
        with torch.no_grad():
            for data in source:
                got.append(m(data)["W"].clone())
        torch.cuda.synchronize()
        assert len(got) == 5
        for a, b in zip(got, want):
            assert torch.allclose(a, b, rtol=0, atol=1e-6)


def test_result_reader_returns_every_step_result():
    """``ResultReader``: the read-back of step k runs on its own stream under step k + 1; every result arrives
    intact even though the device tensor is dropped right after ``read``."""
    from gnn_tracking_b200.graph_store import ResultReader
    host = torch.empty(1 << 20, dtype=torch.float32).pin_memory()
    reader = ResultReader(host, "cuda")
    big = torch.randn(4096, 4096, device="cuda")
    for k in range(6):
        w = torch.full((1 << 20,), float(k), device="cuda") + 0.5
        (big @ big).sum()  # work behind which the copy may hide
        reader.read(w)
        del w
        reader.wait(host=True)
        assert float(host[0]) == k + 0.5 and float(host[-1]) == k + 0.5 and float(host.sum()) == (k + 0.5) * (1 << 20)


def test_reader_thread_binds_next_to_the_gpu():
    """``bind_thread_near_gpu`` narrows the calling thread to NVML's ideal CPUs for the device (or reports that
    the topology is unknown and leaves the thread alone); the main thread is untouched."""
    import os
    import threading

    from gnn_tracking_b200 import graph_store as gs
    before = os.sched_getaffinity(0)
    seen = {}

    def work():
        seen["bound"] = gs.bind_thread_near_gpu("cuda:0")
        seen["cpus"] = os.sched_getaffinity(0)

    th = threading.Thread(target=work)
    th.start()
    th.join()
    assert os.sched_getaffinity(0) == before
    local = gs.gpu_local_cpus("cuda:0")
    if seen["bound"]:
        assert seen["cpus"] == local and local <= before and local
    else:
        assert not local and seen["cpus"] == before
