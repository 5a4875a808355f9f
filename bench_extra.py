"""The other BASELINE.json configurations and modes of ``bench.py`` (``--config tcn_bf16 | pipeline``,
``--mode train``).  Same contract as the headline line: one JSON line, CUDA-event timing with the L2
flushed between timed iterations, the CPU restatement of the reference timed beside it.  Single GPU."""
from __future__ import annotations

import math
import os
import statistics
import time

import torch

import bench


def _events(fn, steps, warmup, flush):
    from gnn_tracking_b200 import ops
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = []
    l0 = ops.launch_count()
    for _ in range(steps):
        flush.zero_()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        t.record()
        evs.append((s, t))
    torch.cuda.synchronize()
    return sum(s.elapsed_time(t) for s, t in evs), ops.launch_count() - l0


def _truth(g: dict, seed: int = 1) -> dict:
    """Per-hit truth for the condensation loss / tracking metrics: ~N/10 particles,
    pt ~ Exp(1) per particle, eta from the hit position, everything reconstructable."""
    gen = torch.Generator().manual_seed(seed)
    n = g["n_nodes"]
    pid = torch.randint(0, n // 10, (n,), generator=gen)
    pt = torch.empty(n // 10).exponential_(1.0, generator=gen)[pid]
    return {"particle_id": pid, "pt": pt, "eta": g["x"][:, 3].clone(), "reconstructable": torch.ones(n, dtype=torch.long)}


def _peak():
    import json
    f = bench.ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        return float(json.loads(f.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


# ------------------------------------------------------------------------------------ config 3
TCN_KW = dict(node_indim=bench.NODE_IN, edge_indim=bench.EDGE_IN, h_dim=128, e_dim=128, hidden_dim=128, L_ec=3, L_hc=8)


def run_tcn_bf16(args) -> None:
    """BASELINE config 3: GraphTCN (node = edge = hidden width 128, 3 + 8 IN layers) under
    ``torch.autocast(bfloat16)`` + the condensation ("potential") loss, 100k nodes / 1M edges."""
    from gnn_tracking_b200 import ops
    from gnn_tracking_b200.graph_store import GraphData
    from gnn_tracking_b200.metrics.losses.oc import CondensationLossTiger
    from gnn_tracking_b200.models.track_condensation_networks import GraphTCN
    from gnn_tracking_b200.plan import build_plan, clear_plan_cache

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    g = bench.make_graph(bench.N_NODES, bench.N_EDGES, seed=0)
    truth = _truth(g)
    n, e = g["n_nodes"], g["n_edges"]
    torch.manual_seed(0)
    model = GraphTCN(**TCN_KW).to(dev)
    loss_fn = CondensationLossTiger()
    x, ei, ea = g["x"].to(dev), g["edge_index"].to(dev), g["edge_attr"].to(dev)
    # A random-init classifier scores every edge ~0.5: the EC threshold is set to its median score so that half
    # of the edges reach the condenser (what a trained classifier's cut does to a TrackML graph), not 0 or all.
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        w0 = model._gtcn.ec(GraphData(x=x, edge_index=ei, edge_attr=ea))["W"].float()
    model._gtcn.hparams["ec_threshold"] = float(w0.median())
    tr = {k: v.to(dev) for k, v in truth.items()}
    hx, hei, hea = g["x"].pin_memory(), g["edge_index"].pin_memory(), g["edge_attr"].pin_memory()
    hloss = torch.empty(4, dtype=torch.float32).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    kept = {}

    def forward(dx, dei, dea):
        clear_plan_cache()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            out = model(GraphData(x=dx, edge_index=dei, edge_attr=dea))
        with torch.no_grad():
            r = loss_fn(beta=out["B"].float(), x=out["H"].float(), ec_hit_mask=out["ec_hit_mask"], **tr)
        kept["edges_hc"] = out["ec_edge_mask"]
        return torch.stack([r.loss_dct[k].float() for k in ("attractive", "repulsive", "coward", "noise")])

    def step_resident():
        return forward(x, ei, ea)

    def step_e2e():
        l = forward(hx.to(dev, non_blocking=True), hei.to(dev, non_blocking=True), hea.to(dev, non_blocking=True))
        hloss.copy_(l, non_blocking=True)

    with bench.ClockSampler(0) as clocks:
        ms, launches = _events(step_resident, args.steps, args.warmup, flush)
        ms_e2e, _ = _events(step_e2e, args.steps, 1, flush)
    n_hc = int(kept["edges_hc"].sum())

    # dominant kernel: the bf16 IN edge kernel of a condenser layer, alone, edge features in sorted order
    plan = build_plan(ei, n)
    gen = torch.Generator().manual_seed(1)
    bf = torch.bfloat16
    ee = torch.randn(e, 128, generator=gen).to(bf).to(dev)
    pi = torch.randn(n, 128, generator=gen).to(bf).to(dev)
    pj = torch.randn(n, 128, generator=gen).to(bf).to(dev)
    rel = model._gtcn.hc_in.network.layers[1].relational_model.linears
    packed = ops.pack_in_edge_bf16([rel[0].weight[:, 256:].contiguous(), rel[1].weight, rel[2].weight], [l.bias for l in rel])
    e_out = None
    ts = []
    for i in range(3 + args.steps):
        flush.zero_()
        del e_out
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        e_out, _ = ops.in_edge_bf16(ee, pi, pj, plan.src_sorted, plan.dst_sorted, packed, n, relu_e=True)
        t.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(s.elapsed_time(t))
    k_ms = statistics.mean(ts)
    alg = e * (16 + 2 * 128 + 2 * 128) + n * 2 * (128 + 128)  # SURVEY 8(d) with s = 2 bytes: 579.2 B/edge at N = E / 10
    peak, peak_src = _peak()
    achieved = alg / (k_ms * 1e-3) / 1e9

    line = {
        "metric": "edges/sec", "value": e * args.steps / (ms * 1e-3), "unit": "edges/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"GraphTCN forward under torch.autocast(bfloat16) + condensation (potential) loss, node = edge = hidden width 128, "
                               f"L_ec=3, L_hc=8, TrackML-shaped synthetic graph {n} nodes / {e} edges (BASELINE config 3); "
                               f"{n_hc} edges pass the EC threshold into the condenser; plan build per step",
                   "l2": "flushed between timed iterations (256 MB write)",
                   "native": "the 11 Interaction-Network edge launches (bf16 tcgen05), plan / filter kernels and the loss kernels; "
                             "encoders, object models and heads are library GEMMs on bf16 operands in this mode"},
        "e2e": {"value": e * args.steps / (ms_e2e * 1e-3), "unit": "edges/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": hx.numel() * 4 + hei.numel() * 8 + hea.numel() * 4, "d2h_bytes_per_step": 16,
                "how": "every step copies x / edge_index / edge_attr from pinned host memory and reads the four loss terms back"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "fused IN edge kernel, bf16 (tcgen05 kind::f16), width 128, one launch incl. the zero-fill of "
                                               "its fp32 aggregate, edge features in destination-sorted order",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "algorithmic_bytes_per_launch": alg, "kernel_ms": k_ms, "peak_source": peak_src},
        "clocks": clocks.summary(),
    }
    if not args.no_cpu:
        t_cpu, threads, n_s, e_s = tcn_cpu_time(reps=1)
        line["cpu_baseline"] = {"value": e_s / t_cpu, "unit": "edges/s", "cores": threads, "kind": "port",
                                "sample": f"oracle GraphTCN forward under torch.autocast('cpu', bfloat16) on a quarter-size graph "
                                          f"({n_s} nodes / {e_s} edges), 1 forward after 1 warm-up, loss not included",
                                "ms_per_step": t_cpu * 1e3}
    bench.emit(line)


def tcn_cpu_time(reps: int):
    from gnn_tracking_b200.models.track_condensation_networks import GraphTCN
    from oracle import in_oracle as O
    g = bench.make_graph(bench.N_NODES // 4, bench.N_EDGES // 4, seed=0)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    sd = {k: v.clone() for k, v in GraphTCN(**TCN_KW).state_dict().items()}
    ts = []
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        for i in range(1 + reps):
            t0 = time.perf_counter()
            O.graph_tcn_forward(g["x"], g["edge_index"], g["edge_attr"], sd)
            if i >= 1:
                ts.append(time.perf_counter() - t0)
    return statistics.median(ts), threads, g["n_nodes"], g["n_edges"]


def run_tcn_bf16_reference(args) -> None:
    t_cpu, threads, n_s, e_s = tcn_cpu_time(reps=max(1, min(args.steps, 3)))
    val = e_s / t_cpu
    bench.emit({"impl": "reference", "metric": "edges/sec", "value": val, "unit": "edges/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t_cpu * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"GraphTCN forward under CPU autocast(bfloat16), width 128, L_ec=3, L_hc=8, quarter-size graph {n_s} nodes / {e_s} edges"},
                "cpu_baseline": {"value": val, "unit": "edges/s", "cores": threads, "kind": "port", "sample": "quarter-size graph per step"},
                "e2e": {"value": val, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})


# ------------------------------------------------------------------------------------ config 5
def run_pipeline(args) -> None:
    """BASELINE config 5: GraphTCN forward (reference-default widths, fp32) + the DBSCAN hyper-parameter scan
    on the latent space of the 100k hits (postprocessing/dbscanscanner.py:146-187): ``n_trials`` clusterings
    with the labels left on the device.  sklearn's DBSCAN on the same coordinates is timed beside it and the
    labels of the first trial are compared."""
    from gnn_tracking_b200 import ops
    from gnn_tracking_b200.graph_store import GraphData
    from gnn_tracking_b200.models.track_condensation_networks import GraphTCN
    from gnn_tracking_b200.plan import clear_plan_cache
    from gnn_tracking_b200.postprocessing.dbscan import DBSCANFastRescan

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    g = bench.make_graph(bench.N_NODES, bench.N_EDGES, seed=0)
    n, e = g["n_nodes"], g["n_edges"]
    torch.manual_seed(0)
    model = GraphTCN(bench.NODE_IN, bench.EDGE_IN, hidden_dim=64, L_ec=3, L_hc=3, h_outdim=3, ec_threshold=0.0).to(dev)
    x, ei, ea = g["x"].to(dev), g["edge_index"].to(dev), g["edge_attr"].to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    n_trials = args.trials
    # The scan explores radii around the cluster size it is looking for (tracks of ~10 hits): eps runs over the
    # 3rd .. 30th nearest-neighbour distance of the latent space (median over 2000 sampled hits), whatever the
    # random-init network makes of it.
    with torch.no_grad():
        h0 = model(GraphData(x=x, edge_index=ei, edge_attr=ea))["H"].float()
        sample = h0[torch.randperm(n, device=dev, generator=torch.Generator(device=dev).manual_seed(0))[:2000]]
        knn = torch.cdist(sample, h0).topk(31, dim=1, largest=False).values.median(dim=0).values  # [31], entry 0 = self
    ks = [3 + round(27 * i / max(1, n_trials - 1)) for i in range(n_trials)]
    trials = [(float(knn[k]) * 1.0001, 1 + i % 3) for i, k in enumerate(ks)]
    hold = {}

    def step():
        clear_plan_cache()
        with torch.no_grad():
            h = model(GraphData(x=x, edge_index=ei, edge_attr=ea))["H"]
        scanner = DBSCANFastRescan(h, max_eps=max(t[0] for t in trials))
        hold["labels"] = [scanner.cluster(eps=eps, min_pts=mp) for eps, mp in trials]
        hold["h"] = h

    with bench.ClockSampler(0) as clocks:
        ms, launches = _events(step, args.steps, args.warmup, flush)
    # forward alone, to split the step
    def fwd():
        clear_plan_cache()
        with torch.no_grad():
            model(GraphData(x=x, edge_index=ei, edge_attr=ea))
    ms_fwd, _ = _events(fwd, args.steps, 1, flush)

    line = {
        "metric": "edges/sec", "value": e * args.steps / (ms * 1e-3), "unit": "edges/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"GraphTCN forward (reference-default widths, hidden 64, L_ec=3, L_hc=3, 3 latent dimensions) + {n_trials} DBSCAN "
                               f"trials of the hyper-parameter scan on the latent space of {n} hits / {e} edges (BASELINE config 5)",
                   "eps": "3rd .. 30th nearest-neighbour distance of the latent space (median over 2000 hits), min_samples 1 .. 3",
                   "split": {"forward_ms": ms_fwd / args.steps, "scan_ms": (ms - ms_fwd) / args.steps,
                             "ms_per_trial": (ms - ms_fwd) / args.steps / n_trials},
                   "l2": "flushed between timed iterations (256 MB write)"},
        "gpu_launches": launches, "clocks": clocks.summary(),
    }
    if not args.no_cpu:
        import numpy as np
        from sklearn.cluster import DBSCAN
        hc = hold["h"].float().cpu().numpy()
        eps, mp = trials[0]
        t0 = time.perf_counter()
        want = DBSCAN(eps=eps, min_samples=mp, n_jobs=-1).fit_predict(hc)
        t_sk = time.perf_counter() - t0
        same = bool(np.array_equal(hold["labels"][0].cpu().numpy(), want))
        if not same:
            raise SystemExit("PARITY FAILURE (pipeline): DBSCAN labels differ from sklearn's on the same latent coordinates")
        line["cpu_baseline"] = {"value": 1.0 / t_sk, "unit": "trials/s", "cores": os.cpu_count() or 1, "kind": "reference",
                                "sample": "one sklearn.cluster.DBSCAN(n_jobs=-1) trial on the same latent coordinates (the library the "
                                          "reference's scanner calls); labels identical to the device trial",
                                "ms_per_trial": t_sk * 1e3}
    bench.emit(line)


# ------------------------------------------------------------------------------------ training step
def run_train(args) -> None:
    """fwd + bwd: one EC training step (forward, BCE, backward, Adam) on the headline graph,
    next to forward + backward of the CPU restatement through torch autograd."""
    from gnn_tracking_b200 import ops
    from gnn_tracking_b200.metrics.losses.ec import EdgeWeightBCELoss
    from gnn_tracking_b200.models.edge_classifier import ECForGraphTCN
    from gnn_tracking_b200.plan import clear_plan_cache

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    g = bench.make_graph(bench.N_NODES, bench.N_EDGES, seed=0)
    n, e = g["n_nodes"], g["n_edges"]
    torch.manual_seed(0)
    model = ECForGraphTCN(**bench.model_kwargs(args.dims)).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    loss_fn = EdgeWeightBCELoss()
    x, ei, ea, y = g["x"].to(dev), g["edge_index"].to(dev), g["edge_attr"].to(dev), g["y"].to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step():
        clear_plan_cache()
        opt.zero_grad(set_to_none=True)
        out = model.forward_tensors(x, ei, ea)
        loss = loss_fn(w=out["W"], y=y)
        loss.backward()
        opt.step()

    torch.cuda.reset_peak_memory_stats()
    with bench.ClockSampler(0) as clocks:
        ms, launches = _events(step, args.steps, args.warmup, flush)
    line = {
        "metric": "edges/sec", "value": e * args.steps / (ms * 1e-3), "unit": "edges/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "training step: " + bench.workload_name(args.dims, n, e) + " + BCE + backward + Adam",
                   "l2": "flushed between timed iterations (256 MB write)",
                   "peak_memory_gib": torch.cuda.max_memory_allocated() / 2 ** 30},
        "gpu_launches": launches, "clocks": clocks.summary(),
    }
    if not args.no_cpu:
        from oracle import in_oracle as O
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        torch.manual_seed(0)
        sd = {k: v.clone().requires_grad_(True) for k, v in ECForGraphTCN(**bench.model_kwargs(args.dims)).state_dict().items()}
        ts = []
        for i in range(3):
            t0 = time.perf_counter()
            w = O.ec_forward(g["x"], g["edge_index"], g["edge_attr"], sd)["W"]
            loss = torch.nn.functional.binary_cross_entropy(w, g["y"].float())
            torch.autograd.grad(loss, list(sd.values()), allow_unused=True)
            if i >= 1:
                ts.append(time.perf_counter() - t0)
        t_cpu = statistics.median(ts)
        line["cpu_baseline"] = {"value": e / t_cpu, "unit": "edges/s", "cores": threads, "kind": "port",
                                "sample": "forward + BCE + backward of the oracle through torch autograd, same full graph, "
                                          "1 warm-up + 2 steps, median (no optimiser step)", "ms_per_step": t_cpu * 1e3}
    bench.emit(line)
