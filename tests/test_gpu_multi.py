"""Node-partitioned edge classifier on 2 GPUs (NCCL halo exchange per layer) against the
single-GPU forward of the same graph: 1e-5 on the owned edges / nodes of every rank.
Needs >= 2 CUDA devices (``gpurun --gpus 2``); skipped otherwise."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from gnn_tracking_b200.models.edge_classifier import ECForGraphTCN
    from gnn_tracking_b200.partition import HaloExchange, partition_graph

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        gen = torch.Generator().manual_seed(11)
        n, e = 6000, 50000
        src = torch.randint(0, n, (e,), generator=gen)
        dst = (src + torch.randint(-40, 41, (e,), generator=gen)).clamp(0, n - 1)
        far = torch.rand(e, generator=gen) < 0.03
        dst = torch.where(far, torch.randint(0, n, (e,), generator=gen), dst)
        ei = torch.stack([src, dst])
        x = torch.randn(n, 14, generator=gen)
        ea = torch.randn(e, 4, generator=gen)
        errs = {}
        for name, kw in (("wide", dict(interaction_node_dim=64, interaction_edge_dim=64, hidden_dim=64)),
                         ("default", dict(hidden_dim=64))):
            torch.manual_seed(0)
            m = ECForGraphTCN(node_indim=14, edge_indim=4, L_ec=3, **kw).to(dev)
            sh = partition_graph(ei, n, world, rank)
            halo = HaloExchange(sh.to(dev))
            with torch.no_grad():
                full = m.forward_tensors(x.to(dev), ei.to(dev), ea.to(dev))
                part = m.forward_tensors(x[sh.node_lo:sh.node_hi].to(dev), sh.edge_index.to(dev).contiguous(),
                                         ea[sh.edge_ids].to(dev), halo=halo)
            torch.cuda.synchronize()
            ids = sh.edge_ids.to(dev)
            errs[name] = (float((part["W"] - full["W"][ids]).abs().max()),
                          float((part["edge_embedding"] - full["edge_embedding"][ids]).abs().max()
                                / full["edge_embedding"].abs().max().clamp_min(1.0)),
                          float((part["node_embedding"] - full["node_embedding"][sh.node_lo:sh.node_hi]).abs().max()
                                / full["node_embedding"].abs().max().clamp_min(1.0)),
                          sh.n_halo, halo.bytes_sent)
        q.put((rank, errs))
    finally:
        dist.destroy_process_group()


def test_partitioned_ec_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errs in res:
        for name, (ew, ee, en, n_halo, sent) in errs.items():
            assert n_halo > 0 and sent > 0, (rank, name)
            assert ew <= TOL and ee <= TOL and en <= TOL, (rank, name, ew, ee, en)


def _grad_worker(rank, world, port, q):
    import torch.distributed as dist

    from gnn_tracking_b200.models.edge_classifier import ECForGraphTCN
    from gnn_tracking_b200.partition import HaloExchange, allreduce_gradients, partition_graph

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        gen = torch.Generator().manual_seed(13)
        n, e = 3000, 24000
        src = torch.randint(0, n, (e,), generator=gen)
        dst = (src + torch.randint(-40, 41, (e,), generator=gen)).clamp(0, n - 1)
        far = torch.rand(e, generator=gen) < 0.05
        dst = torch.where(far, torch.randint(0, n, (e,), generator=gen), dst)
        ei = torch.stack([src, dst])
        x = torch.randn(n, 14, generator=gen)
        ea = torch.randn(e, 4, generator=gen)
        y = (torch.rand(e, generator=gen) < 0.3).float()
        res = {}
        for name, kw in (("wide", dict(interaction_node_dim=64, interaction_edge_dim=64, hidden_dim=64)),
                         ("default", dict(hidden_dim=64))):
            torch.manual_seed(0)
            m = ECForGraphTCN(node_indim=14, edge_indim=4, L_ec=2, **kw).to(dev)
            # single GPU: BCE summed over all edges / E
            m.zero_grad()
            w = m.forward_tensors(x.to(dev), ei.to(dev), ea.to(dev))["W"]
            torch.nn.functional.binary_cross_entropy(w, y.to(dev), reduction="sum").div(e).backward()
            ref = {k: p.grad.clone() for k, p in m.named_parameters()}
            # partitioned: every rank sums over ITS edges, same global normaliser, gradients all-reduced
            m.zero_grad()
            sh = partition_graph(ei, n, world, rank)
            halo = HaloExchange(sh.to(dev))
            wp = m.forward_tensors(x[sh.node_lo:sh.node_hi].to(dev), sh.edge_index.to(dev).contiguous(),
                                   ea[sh.edge_ids].to(dev), halo=halo)["W"]
            torch.nn.functional.binary_cross_entropy(wp, y[sh.edge_ids].to(dev), reduction="sum").div(e).backward()
            allreduce_gradients(m)
            torch.cuda.synchronize()
            worst = 0.0
            for k, p in m.named_parameters():
                scale = float(ref[k].abs().max().clamp_min(1e-12))
                worst = max(worst, float((p.grad - ref[k]).abs().max()) / scale)
            res[name] = (worst, sh.n_halo)
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_partitioned_ec_gradients_match_single_gpu():
    """Backward through the halo exchange (reverse all-to-all-v of the halo rows' gradients + all-reduce of the
    weight gradients): every parameter gradient of the node-partitioned classifier equals the
    single-GPU one at 1e-3 of the largest entry of that tensor (the bar of tests/test_gpu_backward.py: fp32 /
    3xTF32 sums over different edge orders)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, r in res:
        for name, (worst, n_halo) in r.items():
            assert n_halo > 0
            assert worst <= 1e-3, (rank, name, worst)
