"""Radius-graph losses on the CUDA path (``gtb_radius_pair_sum_f32``): the hinge embedding loss
(reference metrics/losses/metric_learning.py:57-178) and the radius-graph condensation loss
(metrics/losses/oc.py:87-248) against the reference's known-answer values
(tests/test_losses.py:112-123,194-203), the outputs of its own classes on td1 / td2 and the CPU
oracle on larger seeded inputs, including a neighbour cap that truncates.  Tolerance rel 2e-5
(fp32 terms, fp64 sums), edge counts exact."""
import pytest
import torch

pytestmark = pytest.mark.gpu
RTOL = 2e-5


@pytest.fixture(autouse=True, params=["grid", "brute"])
def neighbour_search(request, monkeypatch):
    """Every test of this file runs over the cell list (default) and over the all-pairs walk."""
    if request.param == "brute":
        monkeypatch.setenv("GTB_RADIUS_BRUTE", "1")
    else:
        monkeypatch.delenv("GTB_RADIUS_BRUTE", raising=False)
    return request.param


def _close(got, ref, what):
    g, r = float(got), float(ref)
    if r != r:
        assert g != g, what
    else:
        assert g == pytest.approx(r, rel=RTOL, abs=1e-7), what


def _hinge_args(d):
    return dict(x=d["x"].float(), particle_id=d["particle_id"], batch=d["batch"], true_edge_index=d["true_edge_index"],
                pt=d["pt"], eta=d["eta"], reconstructable=d["reconstructable"])


def test_hinge_known_answers(golden_losses):
    from gnn_tracking_b200.metrics.losses.metric_learning import GraphConstructionHingeEmbeddingLoss as Hinge
    d = {k: v.cuda() for k, v in golden_losses["td1"]["data"].items()}
    ka = golden_losses["known_answers"]
    with torch.no_grad():
        r = Hinge()(**_hinge_args(d))
        for k, v in ka["td1_hinge"].items():
            _close(r.loss_dct[k], v, k)
        r = Hinge(rep_normalization="n_rep_edges")(**_hinge_args(d))
        for k, v in ka["td1_hinge_n_rep_edges"].items():
            _close(r.loss_dct[k], v, k)


@pytest.mark.parametrize("td", ["td1", "td2"])
def test_hinge_vs_reference_outputs(td, golden_losses):
    from gnn_tracking_b200.metrics.losses.metric_learning import GraphConstructionHingeEmbeddingLoss as Hinge
    d = {k: v.cuda() for k, v in golden_losses[td]["data"].items()}
    res = golden_losses[td]["results"]
    variants = {"default": {}, "n_rep_edges": dict(rep_normalization="n_rep_edges"),
                "n_att_edges_p2": dict(rep_normalization="n_att_edges", p_attr=2.0, p_rep=2.0, r_emb=0.5),
                "all_hits": dict(rep_oi_only=False)}
    for name, kw in variants.items():
        ref = res[f"hinge_{name}"]
        with torch.no_grad():
            r = Hinge(**kw)(**_hinge_args(d))
        for k in ("attractive", "repulsive"):
            _close(r.loss_dct[k], ref[k], f"{td}.{name}.{k}")
        for k in ("n_hits_oi", "n_edges_att", "n_edges_rep"):
            assert int(r.extra_metrics[k]) == int(ref[k]), (td, name, k)


@pytest.mark.parametrize("td", ["td1", "td2"])
def test_condensation_rg_known_answers(td, golden_losses):
    """tiger == RG == the reference's known answers (tests/test_losses.py:126-139)."""
    from gnn_tracking_b200.metrics.losses.oc import CondensationLossRG
    d = {k: v.cuda() for k, v in golden_losses[td]["data"].items()}
    args = dict(beta=d["beta"].float(), x=d["x"].float(), particle_id=d["particle_id"], reconstructable=d["reconstructable"],
                pt=d["pt"], eta=d["eta"])
    with torch.no_grad():
        r = CondensationLossRG()(**args)
    for k, v in golden_losses["known_answers"][f"{td}_condensation"].items():
        _close(r.loss_dct[k], v, f"{td}.{k}")
    res = golden_losses[td]["results"]
    if "rg_alt" in res:
        with torch.no_grad():
            r = CondensationLossRG(q_min=0.1, pt_thld=0.3, max_eta=3.5)(**args)
        for k in ("attractive", "repulsive", "coward", "noise"):
            _close(r.loss_dct[k], res["rg_alt"][k], f"{td}.alt.{k}")


@pytest.mark.parametrize("cap", [256, 5])
def test_radius_losses_vs_oracle_seeded(cap):
    """3000 hits in 3 dimensions, two batch entries; cap = 5 truncates most neighbourhoods (the
    oracle keeps the lowest indices, as the kernel does)."""
    from gnn_tracking_b200.metrics.losses.metric_learning import GraphConstructionHingeEmbeddingLoss as Hinge
    from gnn_tracking_b200.metrics.losses.oc import condensation_loss_rg
    from oracle import losses_oracle as L
    gen = torch.Generator().manual_seed(17)
    n = 3000
    x = torch.randn(n, 3, generator=gen) * 1.5
    pid = torch.randint(0, 300, (n,), generator=gen)
    # truth is per particle (as in real events): the tiger loss picks condensation points among ALL hits
    # of a particle of interest, the RG loss among the masked ones -- they only agree, as the reference's
    # own test asserts on td1 / td2, when the mask is constant per particle
    pt = (torch.rand(300, generator=gen) * 2)[pid]
    eta = ((torch.rand(300, generator=gen) - 0.5) * 9)[pid]
    reco = (torch.rand(300, generator=gen) < 0.9).long()[pid]
    batch = (torch.arange(n) >= n // 2).long()
    tei = torch.randint(0, n, (2, 4000), generator=gen)
    beta = torch.rand(n, generator=gen).clamp(1e-3, 1 - 1e-3)
    ref, extra = L.hinge_loss(x=x.double(), particle_id=pid, batch=batch, true_edge_index=tei, pt=pt, eta=eta,
                              reconstructable=reco, max_num_neighbors=cap, r_emb=0.8)
    with torch.no_grad():
        r = Hinge(max_num_neighbors=cap, r_emb=0.8)(x=x.cuda(), particle_id=pid.cuda(), batch=batch.cuda(),
                                                      true_edge_index=tei.cuda(), pt=pt.cuda(), eta=eta.cuda(),
                                                      reconstructable=reco.cuda())
    for k in ref:
        _close(r.loss_dct[k], ref[k], f"hinge.{k}")
    for k in extra:
        assert int(r.extra_metrics[k]) == int(extra[k]), k
    mask = L.good_node_mask(pt=pt, particle_id=pid, reconstructable=reco, eta=eta)
    ref = L.condensation_rg(beta=beta.double(), x=x.double(), particle_id=pid, mask=mask, max_num_neighbors=cap)
    with torch.no_grad():
        got, _ = condensation_loss_rg(beta=beta.cuda(), x=x.cuda(), particle_id=pid.cuda(), mask=mask.cuda(), q_min=0.01,
                                      max_num_neighbors=cap)
    for k in ref:
        _close(got[k], ref[k], f"rg.{k}")


def test_rg_condensation_points_per_hit_mask():
    """``eta`` is a per-HIT quantity in the reference's graphs (graph_builder.py:428,451 take it from the
    point cloud), so a particle can straddle |eta| = max_eta: the RG loss then picks its condensation
    points among the masked hits only (oc.py:33-43), attracts only those and normalises with
    ``mask.sum()``.  Every particle here has hits on both sides of the cut, and the most-charged hit of
    most particles is OUTSIDE the mask."""
    from gnn_tracking_b200.metrics.losses.oc import CondensationLossRG
    from oracle import losses_oracle as L
    gen = torch.Generator().manual_seed(31)
    n, n_p = 2400, 240
    x = torch.randn(n, 3, generator=gen) * 1.4
    pid = torch.randint(0, n_p, (n,), generator=gen)
    pt = (0.5 + torch.rand(n_p, generator=gen) * 2)[pid]
    reco = (torch.rand(n_p, generator=gen) < 0.95).long()[pid]
    eta = (torch.rand(n, generator=gen) - 0.5) * 10          # per hit: about 20 % of every particle's hits fail the cut
    beta = torch.rand(n, generator=gen).clamp(1e-3, 1 - 1e-3)
    beta = torch.where(eta.abs() > 4.0, 0.97 + 0.029 * beta, 0.9 * beta)  # the best hits are the cut ones; no ties
    mask = L.good_node_mask(pt=pt, particle_id=pid, reconstructable=reco, eta=eta)
    per_particle = torch.zeros(n_p).index_add_(0, pid, mask.float()) / torch.bincount(pid, minlength=n_p).clamp_min(1)
    assert ((per_particle > 0) & (per_particle < 1)).sum() > n_p // 2  # the mask really splits particles
    ref = L.condensation_rg(beta=beta.double(), x=x.double(), particle_id=pid, mask=mask)
    with torch.no_grad():
        r = CondensationLossRG()(beta=beta.cuda(), x=x.cuda(), particle_id=pid.cuda(), reconstructable=reco.cuda(),
                                 pt=pt.cuda(), eta=eta.cuda())
    for k in ref:
        _close(r.loss_dct[k], ref[k], f"rg_per_hit.{k}")


@pytest.mark.parametrize("cap,p_attr,p_rep", [(256, 1.0, 1.0), (5, 2.0, 2.0)])
def test_radius_loss_gradients(cap, p_attr, p_rep):
    """Gradients of the hinge loss w.r.t. x and of the radius-graph condensation loss w.r.t. x and
    beta (gtb_radius_pair_sum_grad_f32, gtb_edge_dist_pow_grad_f32, gtb_oc_potentials_grad) against
    float64 autograd through the oracle; 2e-5 of the largest entry per tensor."""
    from gnn_tracking_b200.metrics.losses.metric_learning import GraphConstructionHingeEmbeddingLoss as Hinge
    from gnn_tracking_b200.metrics.losses.oc import condensation_loss_rg
    from oracle import losses_oracle as L
    gen = torch.Generator().manual_seed(23)
    n = 2000
    x = torch.randn(n, 3, generator=gen) * 1.2
    pid = torch.randint(0, 200, (n,), generator=gen)
    pt = (torch.rand(200, generator=gen) * 2)[pid]
    eta = ((torch.rand(200, generator=gen) - 0.5) * 9)[pid]
    reco = (torch.rand(200, generator=gen) < 0.9).long()[pid]
    batch = (torch.arange(n) >= n // 2).long()
    tei = torch.randint(0, n, (2, 3000), generator=gen)
    beta = torch.rand(n, generator=gen).clamp(1e-2, 1 - 1e-2)

    def check(name, got, ref):
        got, ref = got.detach().cpu().double(), ref.detach().double()
        scale = float(ref.abs().max())
        err = float((got - ref).abs().max())
        assert scale > 0 and err <= 2e-5 * scale + 1e-9, f"{name}: max|d|={err:.3e} scale={scale:.3e}"

    xr = x.double().requires_grad_()
    ref, _ = L.hinge_loss(x=xr, particle_id=pid, batch=batch, true_edge_index=tei, pt=pt, eta=eta, reconstructable=reco,
                          max_num_neighbors=cap, r_emb=0.8, p_attr=p_attr, p_rep=p_rep)
    (ref["attractive"] + 0.7 * ref["repulsive"]).backward()
    xc = x.cuda().requires_grad_()
    out = Hinge(max_num_neighbors=cap, r_emb=0.8, p_attr=p_attr, p_rep=p_rep, lw_repulsive=0.7)(
        x=xc, particle_id=pid.cuda(), batch=batch.cuda(), true_edge_index=tei.cuda(), pt=pt.cuda(), eta=eta.cuda(),
        reconstructable=reco.cuda())
    out.loss.backward()
    check("hinge x", xc.grad, xr.grad)

    mask = L.good_node_mask(pt=pt, particle_id=pid, reconstructable=reco, eta=eta)
    xr, br = x.double().requires_grad_(), beta.double().requires_grad_()
    ref = L.condensation_rg(beta=br, x=xr, particle_id=pid, mask=mask, max_num_neighbors=cap)
    (ref["attractive"] + 0.5 * ref["repulsive"] + 0.2 * ref["coward"] + 0.3 * ref["noise"]).backward()
    xc, bc = x.cuda().requires_grad_(), beta.cuda().requires_grad_()
    got, _ = condensation_loss_rg(beta=bc, x=xc, particle_id=pid.cuda(), mask=mask.cuda(), q_min=0.01, max_num_neighbors=cap)
    (got["attractive"] + 0.5 * got["repulsive"] + 0.2 * got["coward"] + 0.3 * got["noise"]).backward()
    check("rg x", xc.grad, xr.grad)
    check("rg beta", bc.grad, br.grad)


@pytest.mark.parametrize("method", ["grid", "brute"])
@pytest.mark.parametrize("cap,loop,with_batch", [(32, False, True), (4, False, False), (1000, True, True)])
def test_radius_graph_edge_list(cap, loop, with_batch, method):
    """The materialised radius graph against the oracle's restatement of torch_cluster.radius_graph:
    identical edge list (integer output: bit-exact), including the neighbour cap and batch segments --
    over the cell list (default) and by the all-pairs walk."""
    from functools import partial

    from gnn_tracking_b200 import cluster
    radius_graph = partial(cluster.radius_graph, method=method)
    from oracle import losses_oracle as L
    gen = torch.Generator().manual_seed(5)
    n = 1500
    x = torch.randn(n, 3, generator=gen)
    batch = (torch.arange(n) * 3 // n) if with_batch else None
    ref = L.radius_graph(x, 0.5, batch=batch, max_num_neighbors=cap)
    if loop:
        # the oracle restates loop=False: add the self loops and restore the (centre, neighbour) order
        ref = torch.cat([ref, torch.arange(n).repeat(2, 1)], 1)
        order = torch.argsort(ref[1] * n + ref[0])
        ref = ref[:, order]
    got = radius_graph(x.cuda(), 0.5, batch=None if batch is None else batch.cuda(), loop=loop, max_num_neighbors=cap)
    assert got.dtype == torch.int64 and torch.equal(got.cpu(), ref)
    assert radius_graph(torch.zeros((0, 3), device="cuda"), 1.0).shape == (2, 0)
    flipped = radius_graph(x.cuda(), 0.5, max_num_neighbors=cap, flow="target_to_source")
    assert torch.equal(flipped.flip(0), radius_graph(x.cuda(), 0.5, max_num_neighbors=cap))


@pytest.mark.parametrize("n,d,r,cap", [(20000, 3, 0.08, 64), (20000, 8, 0.9, 16), (5000, 1, 0.001, 8), (5000, 2, 0.05, 3),
                                       (3000, 12, 2.5, 256), (700, 3, 50.0, 100), (4000, 4, -0.2, 32), (2000, 3, 0.0, 8)])
def test_radius_graph_grid_equals_brute_force(n, d, r, cap):
    """Cell list against the all-pairs walk at sizes the python oracle does not reach: the same fp32 distances
    decide, so the edge lists are bit-identical -- clustered points (dense cells, the cap bites), duplicates,
    a radius above the box (one cell), r <= 0 (r*r decides, as in the walk), batch segments."""
    from gnn_tracking_b200.cluster import radius_graph
    gen = torch.Generator().manual_seed(n + d)
    centres = torch.randn(40, d, generator=gen) * 2
    x = centres[torch.randint(0, 40, (n,), generator=gen)] + 0.1 * torch.randn(n, d, generator=gen)
    x[n // 2: n // 2 + 50] = x[:50]                       # exact duplicates
    x[-1] = 40.0                                          # an outlier stretches the box
    batch = torch.sort(torch.randint(0, 3, (n,), generator=gen)).values
    xc, bc = x.cuda(), batch.cuda()
    for b in (None, bc):
        for loop in (False, True):
            grid = radius_graph(xc, r, batch=b, loop=loop, max_num_neighbors=cap, method="grid")
            brute = radius_graph(xc, r, batch=b, loop=loop, max_num_neighbors=cap, method="brute")
            assert grid.shape == brute.shape and torch.equal(grid, brute), (n, d, r, cap, loop)
    if r > 0:
        assert brute.size(1) > 0


def test_radius_graph_grid_non_finite_coordinates():
    """NaN / Inf coordinates match nothing in either search (the distance test fails), wherever they are binned."""
    from gnn_tracking_b200.cluster import radius_graph
    gen = torch.Generator().manual_seed(9)
    x = torch.rand(3000, 3, generator=gen)
    x[5, 0] = float("nan")
    x[17, 2] = float("inf")
    x[99, 1] = float("-inf")
    xc = x.cuda()
    grid, brute = radius_graph(xc, 0.1, method="grid"), radius_graph(xc, 0.1, method="brute")
    assert torch.equal(grid, brute)
    assert not torch.isin(grid, torch.tensor([5, 17, 99], device="cuda")).any()


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n,d,r,cap", [(20000, 3, 0.1, 256), (20000, 3, 0.1, 5), (6000, 8, 0.9, 16), (3000, 2, 0.0, 4)])
def test_pair_sum_grid_equals_brute_force(n, d, r, cap, mode, monkeypatch):
    """gtb_radius_pair_sum[_grad]_grid_f32 against the all-pairs walk on clustered points: the same edges (count
    exact, also where the neighbour cap truncates -- the threshold bisection), the same fp32 terms (float64 sums
    to 1e-12), gradients to float-atomics accuracy."""
    from gnn_tracking_b200.metrics.losses.metric_learning import radius_pair_sum
    gen = torch.Generator().manual_seed(7 * n + d + mode)
    centres = torch.randn(30, d, generator=gen) * 2
    x = (centres[torch.randint(0, 30, (n,), generator=gen)] + 0.1 * torch.randn(n, d, generator=gen)).cuda()
    pid = torch.randint(0, 300, (n,), generator=gen).cuda()
    flag = (torch.rand(n, generator=gen) < (0.3 if mode == 1 else 0.8)).cuda()
    batch = torch.sort(torch.randint(0, 2, (n,), generator=gen)).values.cuda()
    beta = (torch.rand(n, generator=gen) * 0.98 + 0.01).cuda() if mode == 1 else None
    res = {}
    for method in ("grid", "brute"):
        if method == "brute":
            monkeypatch.setenv("GTB_RADIUS_BRUTE", "1")
        else:
            monkeypatch.delenv("GTB_RADIUS_BRUTE", raising=False)
        xg = x.clone().requires_grad_(True)
        bg = None if beta is None else beta.clone().requires_grad_(True)
        out = radius_pair_sum(x=xg, particle_id=pid, src_flag=flag, r=r, mode=mode, batch=batch, beta=bg, q_min=0.01, p=1.0,
                              max_num_neighbors=cap)
        (out[0] * 0.5 + (out[2] if mode == 1 else 0.0)).backward()
        res[method] = (out.detach(), xg.grad, None if bg is None else bg.grad)
    (og, xg_g, bg_g), (ob, xg_b, bg_b) = res["grid"], res["brute"]
    assert float(og[1]) == float(ob[1]) and float(og[3]) == float(ob[3])
    if r > 0:
        assert float(ob[1]) > 0
    assert float(og[0]) == pytest.approx(float(ob[0]), rel=1e-12, abs=1e-12)
    assert float(og[2]) == pytest.approx(float(ob[2]), rel=1e-12, abs=1e-12)
    scale = float(xg_b.abs().max()) + 1e-20
    assert float((xg_g - xg_b).abs().max()) <= 2e-5 * scale
    if bg_b is not None:
        scale = float(bg_b.abs().max()) + 1e-20
        assert float((bg_g - bg_b).abs().max()) <= 2e-5 * scale
