"""Node partition + halo exchange on CPU: the host-side partition logic, and the
per-layer all-to-all-v of halo rows over a world_size-2 (and 3) ``gloo`` group.  The kernels
themselves are CUDA-only; these tests cover the N > 1 plumbing they are fed by."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnn_tracking_b200.partition import GraphShard, HaloExchange, node_ranges, owner_of, partition_graph


def _graph(n, e, seed=0):
    gen = torch.Generator().manual_seed(seed)
    # mostly local edges (phi-sorted nodes) plus a few long-range ones
    src = torch.randint(0, n, (e,), generator=gen)
    dst = (src + torch.randint(-20, 21, (e,), generator=gen)).clamp(0, n - 1)
    far = torch.rand(e, generator=gen) < 0.05
    dst = torch.where(far, torch.randint(0, n, (e,), generator=gen), dst)
    return torch.stack([src, dst])


@pytest.mark.parametrize("n,world", [(10, 3), (1000, 2), (1001, 4), (7, 8)])
def test_node_ranges_and_owner(n, world):
    r = node_ranges(n, world)
    assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    ids = torch.arange(n)
    own = owner_of(ids, n, world)
    for p, (lo, hi) in enumerate(r):
        assert torch.all(own[lo:hi] == p)


@pytest.mark.parametrize("world", [1, 2, 4])
def test_partition_covers_every_edge_once(world):
    n, e = 500, 4000
    ei = _graph(n, e)
    shards = [partition_graph(ei, n, world, p) for p in range(world)]
    all_ids = torch.cat([s.edge_ids for s in shards])
    assert torch.equal(torch.sort(all_ids).values, torch.arange(e))
    for s in shards:
        assert torch.all(s.edge_ids[1:] > s.edge_ids[:-1])  # original edge order kept inside a shard
        src_l, dst_l = s.edge_index
        assert dst_l.numel() == 0 or (int(dst_l.max()) < s.n_owned and int(src_l.max()) < s.n_local)
        g_src = torch.where(src_l < s.n_owned, src_l + s.node_lo,
                            s.halo_ids[(src_l - s.n_owned).clamp_min(0)] if s.n_halo else src_l)
        assert torch.equal(g_src, ei[0][s.edge_ids])
        assert torch.equal(dst_l + s.node_lo, ei[1][s.edge_ids])
        assert sum(s.recv_counts) == s.n_halo and sum(s.send_counts) == s.send_idx.numel()
    # what p sends to q is exactly what q expects from p, in ascending global id
    for p in range(world):
        for q in range(world):
            assert shards[p].send_counts[q] == shards[q].recv_counts[p]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, e, width, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ei = _graph(n, e, seed=3)
        table = torch.randn(n, width, generator=torch.Generator().manual_seed(9))
        sh = partition_graph(ei, n, world, rank)
        halo = HaloExchange(sh)
        ok = True
        for _ in range(2):  # twice: the exchange is re-entrant (one call per layer)
            ext = halo.extend(table[sh.node_lo:sh.node_hi].clone())
            exp = torch.cat([table[sh.node_lo:sh.node_hi], table[sh.halo_ids]])
            ok = ok and torch.equal(ext, exp)
        # a per-edge gather through the local numbering equals the global gather
        ok = ok and torch.equal(ext[sh.edge_index[0]], table[ei[0][sh.edge_ids]])
        # backward through the halo: reverse() is the adjoint of extend(), i.e. summed over the ranks
        # <extend(t), g> == <t, reverse(g)>; and autograd goes through halo_extend
        from gnn_tracking_b200.partition import halo_extend
        g = torch.randn(sh.n_local, width, generator=torch.Generator().manual_seed(100 + rank), dtype=torch.float64)
        t = table[sh.node_lo:sh.node_hi].double().clone().requires_grad_(True)
        lhs = (halo_extend(t, halo) * g).sum()
        lhs.backward()
        rhs = (t.detach() * halo.reverse(g)).sum()
        both = torch.stack([lhs.detach(), rhs])
        dist.all_reduce(both)
        ok = ok and bool(torch.allclose(both[0], both[1], rtol=1e-12)) and bool(torch.allclose(t.grad, halo.reverse(g)))
        q.put((rank, ok, sh.n_halo, halo.bytes_sent))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 300, 2500, 8, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res), res
    assert any(nh > 0 for _, _, nh, _ in res)


def test_shard_single_rank_is_identity():
    ei = _graph(50, 300)
    s = partition_graph(ei, 50, 1, 0)
    assert s.n_halo == 0 and torch.equal(s.edge_index, ei) and torch.equal(s.edge_ids, torch.arange(300))
    t = torch.randn(50, 4)
    assert torch.equal(HaloExchange(s).extend(t), t)
