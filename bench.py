#!/usr/bin/env python
"""Benchmark of the Interaction-Network hot path (BASELINE.json metric:
"edges/sec on 1M-edge TrackML-shape graph ...; % HBM roofline").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dims wide|default]

A *step* is one full ``ECForGraphTCN`` forward (encoders -> 3 x IN -> W head, hidden 64,
fp32) over one synthetic TrackML-shaped graph of 100k nodes / 1M directed edges
(BASELINE.json configs[1]), including the per-graph plan build (destination sort).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

N_NODES, N_EDGES = 100_000, 1_000_000
GRAPH_KIND = "trackml"  # --graph
NODE_IN, EDGE_IN, HIDDEN, L_EC = 14, 4, 64, 3


# ------------------------------------------------------------------ synthetic graph
def make_graph(n_nodes: int, n_edges: int, seed: int = 0) -> dict:
    """Seeded TrackML-shaped graph: hits on 10 concentric layers, candidate edges
    between adjacent layers inside a phi window, doubled in both directions with
    sign-flipped (dr, dphi, dz) and unchanged dR (reference graph_builder.py:371-378,
    431-438); layer-pair block order, NOT destination sorted."""
    gen = torch.Generator().manual_seed(seed)
    n_layers = 10
    per = n_nodes // n_layers
    n_nodes = per * n_layers
    layer = torch.arange(n_layers).repeat_interleave(per)
    r = (layer.float() + 1.0) * 0.1 + torch.randn(n_nodes, generator=gen) * 1e-3
    phi = (torch.rand(n_nodes, generator=gen) * 2 - 1) * math.pi
    z = torch.randn(n_nodes, generator=gen) * 0.3
    half = n_edges // 2
    want_per_pair = half / (n_layers - 1) * 1.15
    delta = want_per_pair / per * math.pi / per  # expected neighbours per hit = per * delta / pi
    srcs, dsts = [], []
    for l in range(n_layers - 1):
        a = torch.arange(l * per, (l + 1) * per)
        b = torch.arange((l + 1) * per, (l + 2) * per)
        order = torch.argsort(phi[b])
        pb = phi[b][order]
        lo = torch.searchsorted(pb, phi[a] - delta)
        hi = torch.searchsorted(pb, phi[a] + delta)
        cnt = hi - lo
        src = a.repeat_interleave(cnt)
        start = torch.cumsum(cnt, 0) - cnt
        off = torch.arange(int(cnt.sum())) - start.repeat_interleave(cnt)
        dst = b[order][lo.repeat_interleave(cnt) + off]
        srcs.append(src)
        dsts.append(dst)
    src, dst = torch.cat(srcs), torch.cat(dsts)
    assert src.numel() >= half, (src.numel(), half)
    keep = torch.randperm(src.numel(), generator=gen)[:half].sort().values
    src, dst = src[keep], dst[keep]
    dr, dphi, dz = r[dst] - r[src], phi[dst] - phi[src], z[dst] - z[src]
    dR = torch.sqrt(dphi ** 2 + (dz * 0.5) ** 2)
    fwd = torch.stack([dr, dphi, dz, dR], 1)
    bwd = torch.stack([-dr, -dphi, -dz, dR], 1)
    edge_index = torch.cat([torch.stack([src, dst]), torch.stack([dst, src])], 1).contiguous()
    edge_attr = torch.cat([fwd, bwd], 0).contiguous()
    if GRAPH_KIND == "uniform":  # the worst-locality control: same sizes and features, random endpoints
        edge_index = torch.randint(0, n_nodes, (2, edge_index.size(1)), generator=gen)
    eta = torch.asinh(z / r)
    x = torch.cat([torch.stack([r, phi / math.pi, z, eta, r * torch.cos(phi), r * torch.sin(phi)], 1),
                   torch.randn(n_nodes, 8, generator=gen)], 1).contiguous()
    pid = torch.randint(0, n_nodes // 10, (n_nodes,), generator=gen)
    y = (pid[edge_index[0]] == pid[edge_index[1]]) & (pid[edge_index[0]] > 0)
    return {"x": x, "edge_index": edge_index, "edge_attr": edge_attr, "y": y, "n_nodes": n_nodes,
            "n_edges": edge_index.size(1)}


def model_kwargs(dims: str) -> dict:
    kw = dict(node_indim=NODE_IN, edge_indim=EDGE_IN, hidden_dim=HIDDEN, L_ec=L_EC)
    if dims == "wide":
        kw.update(interaction_node_dim=HIDDEN, interaction_edge_dim=HIDDEN)
    return kw


def workload_name(dims: str, n: int, e: int) -> str:
    d = "Dn=De=H=64 (wide)" if dims == "wide" else "Dn=5,De=4,H=64 (reference-default widths)"
    kind = "TrackML-shaped synthetic graph" if GRAPH_KIND == "trackml" else "UNIFORM-RANDOM-edge control graph"
    return f"ECForGraphTCN forward, L_ec=3, {d}, fp32, {kind} {n} nodes / {e} edges, plan build per step"


def layer_algorithmic_bytes(n: int, e: int, dn: int, de: int) -> float:
    """B_layer = E*(2*idx + s*De_in + s*De_out) + N*s*(Dn_in + Dn_out), s = 4, idx = 8."""
    return e * (16 + 4 * de + 4 * de) + n * 4 * (dn + dn)


# ------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        # NVML in-process when the bindings are there (microseconds per query); the nvidia-smi
        # subprocess (~100 ms per query, and it holds up kernel launches while it runs) otherwise
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = [0x8, 0x40, 0x20, 0x4]  # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self._stop.is_set():
                r = int(get_reasons(h))
                self.samples.append([str(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), str(mx)]
                                    + ["Active" if r & b else "Not Active" for b in bits])
                self._stop.wait(0.005)
            return
        except Exception:  # noqa: BLE001 - no NVML bindings / no permission: fall back to the CLI
            pass
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [s.strip() for s in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self) -> dict:
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# ---------------------------------------------------------------------- CPU legs
def cpu_forward_time(g: dict, dims: str, reps: int, warm: int) -> tuple[float, int, dict]:
    """The CPU restatement of the reference forward (oracle/in_oracle.py, pinned to the
    reference's own classes) on all host threads; returns (median seconds, threads, outputs)."""
    from oracle import in_oracle as O
    from gnn_tracking_b200.models.edge_classifier import ECForGraphTCN
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    sd = {k: v.clone() for k, v in ECForGraphTCN(**model_kwargs(dims)).state_dict().items()}
    ts = []
    with torch.no_grad():
        for i in range(warm + reps):
            t0 = time.perf_counter()
            ref = O.ec_forward(g["x"], g["edge_index"], g["edge_attr"], sd)
            if i >= warm:
                ts.append(time.perf_counter() - t0)
    return statistics.median(ts), threads, ref


def parity_report(out: dict, ref: dict, what: str, tol: float = 1e-5) -> dict:
    """max |GPU - oracle| per output of the edge classifier (edge_classifier.py:89-121) at
    tol * max(1, max|oracle|); a mismatch fails the run: a fast kernel with different results is not done."""
    rep = {}
    for k in ("W", "node_embedding", "edge_embedding"):
        r = ref[k].float()
        o = out[k].detach().float().cpu().reshape(r.shape)
        err = float((o - r).abs().max()) if r.numel() else 0.0
        scale = max(1.0, float(r.abs().max())) if r.numel() else 1.0
        rep[k] = err / scale
        if not err <= tol * scale:
            raise SystemExit(f"PARITY FAILURE ({what}): {k} max|err| {err:.3e} > {tol:g} * {scale:.3g}")
    return rep


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the same full graph as the GPU arm (about 1 s per forward on 16 host threads: 25 steps stay well
    # inside "a few minutes"); only an explicitly long run is cut down
    n, e = N_NODES, N_EDGES
    sample = "same full graph per step"
    if args.steps + args.warmup > 120:
        n, e = N_NODES // 4, N_EDGES // 4
        sample = "quarter-size graph (25k nodes / 250k edges) per step"
    g = make_graph(n, e)
    from oracle import in_oracle as O
    from gnn_tracking_b200.models.edge_classifier import ECForGraphTCN
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    sd = {k: v.clone() for k, v in ECForGraphTCN(**model_kwargs(args.dims)).state_dict().items()}
    with torch.no_grad():
        for _ in range(args.warmup):
            O.ec_forward(g["x"], g["edge_index"], g["edge_attr"], sd)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            O.ec_forward(g["x"], g["edge_index"], g["edge_attr"], sd)
        dt = time.perf_counter() - t0
    val = g["n_edges"] * args.steps / dt
    line = {
        "impl": "reference", "metric": "edges/sec", "value": val, "unit": "edges/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.dims, g["n_nodes"], g["n_edges"]),
                   "note": "CPU restatement of the reference forward (oracle/in_oracle.py, op-for-op the PyG/ATen chain; "
                           "the Python reference cannot travel to the GPU box)"},
        "cpu_baseline": {"value": val, "unit": "edges/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# -------------------------------------------------------------------------- ours
def relabel_by_phi(g: dict) -> dict:
    """Node ids in ascending phi: contiguous id ranges are then phi wedges, and the graph
    builder's tight phi cut keeps most edges inside one range."""
    order = torch.argsort(g["x"][:, 1], stable=True)
    new_id = torch.empty_like(order)
    new_id[order] = torch.arange(order.numel())
    out = dict(g)
    out["x"] = g["x"][order].contiguous()
    out["edge_index"] = new_id[g["edge_index"]].contiguous()
    return out


def edge_kernel_time(model, x_dim, plan, n, e, dev, flush, reps, sorted_edges=True):
    """The dominant kernel alone: one launch of the fused IN edge kernel of a middle layer (ReLU on
    load; gathered pre-projected node tables, relational MLP, store, segmented sum), timed with CUDA
    events on the launching stream, L2 flushed before every launch.  ``sorted_edges``: the edge features
    come and go in the plan's destination-sorted order, as they do between the layers of the stack
    (contiguous tiles); False: in the caller's order through ``perm`` (first / last layer of a stack)."""
    from gnn_tracking_b200 import ops
    from gnn_tracking_b200.ops import Block
    layer = model.ec_resin.network.layers[1]
    rel = layer.relational_model
    dn, de = x_dim
    gen = torch.Generator(device="cpu").manual_seed(1)
    xx = torch.randn(n, dn, generator=gen).to(dev)
    ee = torch.randn(e, de, generator=gen).to(dev)
    blocks = [Block(xx, plan.dst_sorted, True, sorted_index=True), Block(xx, plan.src_sorted, True),
              Block(ee, None, True) if sorted_edges else Block(ee, plan.perm, True)]
    widths = [dn, dn, de]
    n0 = rel.linears[0].out_features
    projected = [len(rel.linears) >= 2 and n0 % 4 == 0 and 2 * n <= e] * 2 + [False]
    packed, proj = rel._cache.get(rel.linears, widths, projected)
    cur = []
    for i, b in enumerate(blocks):
        if projected[i]:
            table = ops.fused_mlp([Block(b.tensor, None, b.relu)], n, proj[i])
            cur.append(Block(table, b.index, False, projected=True, sorted_index=b.sorted_index))
        else:
            cur.append(b)
    out = torch.empty((e, de), device=dev)
    aggr = torch.zeros((n, de), device=dev)
    ts = []
    for i in range(3 + reps):
        flush.zero_()
        aggr.zero_()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ops.fused_mlp(cur, e, packed[0], out=out, out_index=None if sorted_edges else plan.perm, aggr=aggr,
                      seg_id=plan.dst_sorted, rowptr=plan.rowptr)
        t.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(s.elapsed_time(t))
    return statistics.mean(ts), ("tcgen05" if packed[0].impl == ops.IMPL_TCGEN05 else "ffma")


def layer_time(model, plan, n, e, dev, flush, reps):
    """One whole middle IN layer of the stack as the forward runs it (B_layer / t_layer): the fused
    edge launch + the one-launch node side (object model, residual, the next layer's two pre-projections, the
    aggregate handed back zeroed), timed together with CUDA events, L2 flushed before every repetition.
    None when the layer is not on the two-launch path (reference-default widths)."""
    from gnn_tracking_b200 import ops
    net = model.ec_resin.network
    gen = torch.Generator(device="cpu").manual_seed(2)
    xx = torch.randn(n, HIDDEN, generator=gen).to(dev)
    ee = torch.randn(e, HIDDEN, generator=gen).to(dev)
    with torch.no_grad():
        if len(net.layers) < 3 or not net.fused_ok(xx, ee):
            return None
        layer, nxt = net.layers[1], net.layers[2].input_projection()
        aggr = torch.zeros((n, HIDDEN), device=dev)
        _, pa, pb = ops.in_node_fused(xx, False, proj=layer.input_projection(), proj_relu=True)
        ts = []
        for i in range(3 + reps):
            flush.zero_()
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            layer.forward_fused(xx, plan, ee, relu_x=True, relu_e=True, res=xx, res_a=0.7071067811865476, res_b=0.7071067811865476,
                                e_sorted=True, out_sorted=True, tables=(pa, pb), aggr=aggr, nxt=nxt, nxt_relu=True)
            t.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(s.elapsed_time(t))
    return statistics.mean(ts)


def run_ours(args) -> None:
    import torch.distributed as dist
    from gnn_tracking_b200 import ops
    from gnn_tracking_b200.models.edge_classifier import ECForGraphTCN
    from gnn_tracking_b200.partition import HaloExchange, partition_graph
    from gnn_tracking_b200.plan import build_plan, clear_plan_cache

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    partitioned = world > 1 and args.multi == "partitioned"
    halo = None
    if partitioned:
        # weak scaling on ONE graph: world x (100k nodes / 1M edges), nodes relabelled by phi and
        # partitioned into contiguous ranges; every rank owns the edges that END in its range and
        # receives its halo rows by one all-to-all-v per layer (+ one for the W head)
        scale = args.total_scale if args.total_scale > 0 else world
        gg = relabel_by_phi(make_graph(int(N_NODES * scale), int(N_EDGES * scale), seed=0))
        shard = partition_graph(gg["edge_index"], gg["n_nodes"], world, rank)
        g = {"x": gg["x"][shard.node_lo:shard.node_hi].contiguous(), "edge_index": shard.edge_index.contiguous(),
             "edge_attr": gg["edge_attr"][shard.edge_ids].contiguous(), "n_nodes": shard.n_local,
             "n_edges": int(shard.edge_ids.numel())}
        halo = HaloExchange(shard.to(dev))
        n_total_edges = gg["n_edges"]
        halo_frac = shard.n_halo / max(1, shard.n_owned)
        gg_full = gg if rank == 0 else None  # rank 0 checks its shard against the single-GPU forward below
        del gg
    else:
        # independent graphs: every rank owns one full graph per step (as the reference trains with
        # batch_size=1 graph per step); no data-path collective
        g = make_graph(N_NODES, N_EDGES, seed=rank)
        n_total_edges = g["n_edges"] * world
        halo_frac = 0.0
    n, e = g["n_nodes"], g["n_edges"]
    torch.manual_seed(0)
    model = ECForGraphTCN(**model_kwargs(args.dims)).to(dev)
    x, ei, ea = g["x"].to(dev), g["edge_index"].to(dev), g["edge_attr"].to(dev)
    hx, hei, hea = g["x"].pin_memory(), g["edge_index"].pin_memory(), g["edge_attr"].pin_memory()
    hw = torch.empty(e, dtype=torch.float32).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_eager():
        clear_plan_cache()
        with torch.no_grad():
            return model.forward_tensors(x, ei, ea, halo=halo)

    # The step as ONE CUDA graph launch (graphs.CapturedForward: plan build + forward recorded over static input
    # buffers, replayed per step; same kernels, same work, no per-launch host time).  GTB_BENCH_NO_GRAPH=1: eager.
    # N > 1: the NCCL halo exchanges are recorded inside the graph as well (GTB_BENCH_GRAPH_MULTI=0: eager there).
    cap = None
    if not os.environ.get("GTB_BENCH_NO_GRAPH") and (world == 1 or os.environ.get("GTB_BENCH_GRAPH_MULTI", "1") != "0"):
        from gnn_tracking_b200.graphs import CapturedForward
        try:
            cap = CapturedForward(model, x, ei, ea, halo=halo)
        except Exception as exc:  # noqa: BLE001 - a capture that fails must not take the measurement down: eager launches
            print(f"bench: CUDA graph capture failed ({type(exc).__name__}: {exc}); running eagerly", file=sys.stderr)
            cap = None

    def step_resident():
        return cap.replay() if cap is not None else step_eager()

    def step_e2e():
        if cap is not None:  # pinned host -> the graph's static input buffers -> replay -> read back
            out = cap(hx, hei, hea)
        else:
            clear_plan_cache()
            dx, dei, dea = hx.to(dev, non_blocking=True), hei.to(dev, non_blocking=True), hea.to(dev, non_blocking=True)
            with torch.no_grad():
                out = model.forward_tensors(dx, dei, dea, halo=halo)
        hw.copy_(out["W"], non_blocking=True)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    multi_parity = None
    if partitioned:
        # multi-GPU parity inside the driver-run command: the partitioned forward of rank 0's shard against
        # the single-GPU forward of the WHOLE graph on rank 0 (same kernels, no halo), 1e-5 * scale
        out_p = step_resident()
        barrier()
        if rank == 0:
            with torch.no_grad():
                full = model.forward_tensors(gg_full["x"].to(dev), gg_full["edge_index"].to(dev), gg_full["edge_attr"].to(dev))
            ids = shard.edge_ids.to(dev)
            ref = {"W": full["W"][ids].cpu(), "edge_embedding": full["edge_embedding"][ids].cpu(),
                   "node_embedding": full["node_embedding"][shard.node_lo:shard.node_hi].cpu()}
            multi_parity = parity_report(out_p, ref, f"partitioned x{world} vs single GPU")
            del full, ref, gg_full
            torch.cuda.empty_cache()
        barrier()

    def timed_e2e_pipelined(steps, warmup):
        """The loop a user of the loader writes: ``for data in DevicePrefetcher(host graphs): model(data)``.
        Every step copies its own inputs from pinned host memory and reads its result back; the copy
        of step k + 1 is in flight on the loader's side stream while step k computes.  One event pair
        around the ``steps`` timed steps of one running loop (per-step L2 flush and result copies included)."""
        from gnn_tracking_b200.graph_store import DevicePrefetcher, GraphData, ResultReader
        host_graph = GraphData(x=hx, edge_index=hei, edge_attr=hea)
        reader = ResultReader(hw, dev)

        # ONE loader loop over warm-up + timed steps: the timed region starts (barrier + synchronize, then the start
        # event) in front of step `warmup`, whose copy was prefetched under the step before -- steady state, as in a
        # training / inference loop that has been running for a while
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i = 0
        for data in DevicePrefetcher((host_graph for _ in range(warmup + steps)), dev):
            if i == warmup:
                barrier()
                s.record()
            i += 1
            flush.zero_()
            if cap is not None:
                # the prefetched graph into the static inputs (device copies, 37.6 MB), one graph launch; W leaves
                # the static output buffer before the next replay overwrites it
                w = cap(data.x, data.edge_index, data.edge_attr)["W"].clone()
            else:
                clear_plan_cache()
                with torch.no_grad():
                    w = model.forward_tensors(data.x, data.edge_index, data.edge_attr, halo=halo)["W"]
            reader.read(w)  # read-back of step k on its own stream, under step k + 1
        reader.wait()       # the timed region ends behind the last read-back
        t.record()
        barrier()
        ms = s.elapsed_time(t)
        if world > 1:
            tt = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        evs = []
        l0 = ops.launch_count()
        for _ in range(steps):
            flush.zero_()  # L2 flush between timed iterations (outside the events)
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            t.record()
            evs.append((s, t))
        barrier()
        launches = ops.launch_count() - l0
        ms = sum(s.elapsed_time(t) for s, t in evs)
        if world > 1:
            tt = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, launches

    with ClockSampler(local) as clocks:
        ms, launches = timed(step_resident, args.steps, args.warmup)
        ms_e2e_serial, _ = timed(step_e2e, args.steps, max(1, args.warmup))
    # outside the sampler: an nvidia-smi query takes ~100 ms and holds up kernel launches while it runs; the
    # legs above ride through that on their launch backlog, the throttled loader loop (two steps ahead) cannot
    # (at N > 1 as well: every rank runs its own loader loop, the halo exchanges keep the ranks in step)
    ms_e2e = timed_e2e_pipelined(args.steps, max(1, args.warmup)) if not os.environ.get("GTB_BENCH_NO_PIPELINED") else None

    # ---- dominant kernel alone (rank 0's graph)
    dn, de = (HIDDEN, HIDDEN) if args.dims == "wide" else (5, 4)
    plan = build_plan(ei, n)
    k_ms, k_impl = edge_kernel_time(model, (dn, de), plan, n, e, dev, flush, args.steps, sorted_edges=True)
    k_ms_perm, _ = edge_kernel_time(model, (dn, de), plan, n, e, dev, flush, args.steps, sorted_edges=False)
    l_ms = layer_time(model, plan, n, e, dev, flush, args.steps) if args.dims == "wide" and halo is None else None
    if world > 1:
        hf = torch.tensor([halo_frac], device=dev, dtype=torch.float64)
        dist.all_reduce(hf, op=dist.ReduceOp.MAX)
        halo_frac = float(hf.item())

    def leave():
        """Tear-down at N > 1.  With NCCL exchanges recorded inside a CUDA graph (GTB_BENCH_GRAPH_MULTI=1)
        ``destroy_process_group`` was seen to hang behind the finished run: the graph is dropped first and the process
        then leaves without the collective tear-down (everything measured has been synchronised and printed)."""
        nonlocal cap
        if world == 1:
            return
        if cap is not None:
            cap = None
            torch.cuda.synchronize()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()

    if rank != 0:
        leave()
        return

    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
    n_owned = x.size(0)
    alg = layer_algorithmic_bytes(n_owned, e, dn, de)
    achieved = alg / (k_ms * 1e-3) / 1e9
    traffic = None
    tf = ROOT / "profiles" / "roofline_traffic.json"
    if tf.exists():
        traffic = json.loads(tf.read_text()).get(f"{args.dims}_{k_impl}")

    value = n_total_edges * args.steps / (ms * 1e-3)
    # two end-to-end loops are timed (serial copy -> compute -> read back, and the prefetching loader at
    # N = 1); the line carries both and `value` is the faster of the two
    e2e_serial_val = n_total_edges * args.steps / (ms_e2e_serial * 1e-3)
    e2e_pipelined_val = n_total_edges * args.steps / (ms_e2e * 1e-3) if ms_e2e is not None else None
    e2e_val = max(e2e_pipelined_val or 0.0, e2e_serial_val)
    e2e_ms = min(ms_e2e if ms_e2e is not None else ms_e2e_serial, ms_e2e_serial) / args.steps
    multi = ("one graph of %g x (100k nodes / 1M edges), node-partitioned by phi wedge, edges owned by their destination's rank, "
             "one NCCL all-to-all-v of halo rows per IN layer + one for the W head; max halo/owned = %.3f"
             % (args.total_scale if args.total_scale > 0 else world, halo_frac)
             if partitioned else "one independent graph per rank per step, no data-path collective")
    line = {
        "metric": "edges/sec", "value": value, "unit": "edges/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak" if not (partitioned and args.total_scale > 0) else "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": (workload_name(args.dims, N_NODES, N_EDGES) + (f" x {world} ranks" if world > 1 else "")
                                if not (partitioned and args.total_scale > 0) else
                                workload_name(args.dims, int(N_NODES * args.total_scale), int(N_EDGES * args.total_scale))
                                + f" partitioned over {world} ranks"),
                   "l2": "flushed between timed iterations (256 MB write)", "multi_gpu": multi,
                   "launch": ("one CUDA graph launch per step (graphs.CapturedForward: plan build + forward over static "
                              "input buffers, %d library kernels per replay)" % cap.launches) if cap is not None else
                             "eager: one host launch per kernel",
                   "impl": os.environ.get("GTB_IMPL", "auto")},
        "e2e": {"value": e2e_val, "unit": "edges/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": hx.numel() * 4 + hei.numel() * 8 + hea.numel() * 4, "d2h_bytes_per_step": e * 4,
                "how": "every step copies its inputs from pinned host memory and reads W back; value = the faster of "
                       "(a) graph_store.DevicePrefetcher / ResultReader loop (copy of step k+1 and read-back of step k-1 on "
                       "side streams under step k; warm-up steps and timed steps in ONE running loop, one event pair "
                       "around the timed steps, L2 flush inside, the last read-back inside) and (b) serial copy -> "
                       "forward -> read back with per-step events",
                "pipelined_value": e2e_pipelined_val,
                "pipelined_ms_per_step": ms_e2e / args.steps if ms_e2e is not None else None,
                "serial_value": e2e_serial_val, "serial_ms_per_step": ms_e2e_serial / args.steps},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm",
                     "kernel": f"fused IN edge kernel ({k_impl}): gathered pre-projected node rows + relational MLP + store + "
                               "segmented sum, one layer, one launch, edge features in destination-sorted order as between "
                               "the layers of the stack (kernel_ms_perm: the same launch reading / writing the caller's "
                               "edge order through perm, as the last layer does)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "algorithmic_bytes_per_launch": alg, "kernel_ms": k_ms, "kernel_ms_perm": k_ms_perm,
                     "frac_perm": alg / (k_ms_perm * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                     # the whole layer (edge launch + one-launch node side) against the same bytes
                     "layer_ms": l_ms, "layer_frac": (alg / (l_ms * 1e-3) / 1e9 / peak) if l_ms else None},
        "clocks": clocks.summary(),
    }
    if multi_parity is not None:
        line["parity"] = {"vs": "single-GPU forward of the whole graph on rank 0 (rank 0's shard)", "tol": 1e-5,
                          "max_err_over_scale": multi_parity}
    if world == 1 and not args.no_cpu:
        t_cpu, threads, ref = cpu_forward_time(g, args.dims, reps=3, warm=1)
        line["cpu_baseline"] = {"value": e / t_cpu, "unit": "edges/s", "cores": threads, "kind": "port",
                                "sample": "same full graph, 1 warm-up + 3 forwards, median",
                                "ms_per_step": t_cpu * 1e3}
        # full-size parity: the GPU forward of the bench graph against the oracle's, same weights
        line["parity"] = {"vs": "oracle/in_oracle.py on the same full graph and weights", "tol": 1e-5,
                          "max_err_over_scale": parity_report(step_resident(), ref, "full-size bench graph")}
    emit(line)
    leave()


_RESULT_FD = None


def emit(line: dict) -> None:
    """The ONE line of stdout (libraries' banners, e.g. NCCL's version line, went to stderr)."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main() -> None:
    global _RESULT_FD
    # keep file descriptor 1 for the result line only: native libraries that print to stdout (NCCL with
    # NCCL_DEBUG=VERSION prints "NCCL version ...") are pointed at stderr
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dims", default="wide", choices=["wide", "default"])
    ap.add_argument("--config", default="ec", choices=["ec", "tcn_bf16", "pipeline"],
                    help="ec: BASELINE config 2 (headline); tcn_bf16: config 3; pipeline: config 5 (see bench_extra.py)")
    ap.add_argument("--mode", default="forward", choices=["forward", "train"], help="train: one EC training step (fwd + bwd + Adam)")
    ap.add_argument("--graph", default="trackml", choices=["trackml", "uniform"],
                    help="uniform: edge_index = randint(0, N, (2, E)), the worst-locality control of SURVEY 8d / 8e")
    ap.add_argument("--trials", type=int, default=20, help="DBSCAN trials per step of --config pipeline")
    ap.add_argument("--total-scale", type=float, default=0.0,
                    help="N > 1, partitioned: the ONE graph has total-scale x (100k nodes / 1M edges); default = N (weak scaling). "
                         "BASELINE config 4 (500k nodes / 5M edges on 8 GPUs) is --gpus 8 --total-scale 5")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--multi", default="partitioned", choices=["partitioned", "independent"],
                    help="N > 1: one node-partitioned graph with halo exchange (default) or one graph per rank")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    global GRAPH_KIND
    GRAPH_KIND = args.graph
    if args.config != "ec" or args.mode != "forward":
        import bench_extra
        # this file runs as __main__: bench_extra holds a second copy of the module
        bench_extra.bench._RESULT_FD, bench_extra.bench.GRAPH_KIND = _RESULT_FD, GRAPH_KIND
        if int(os.environ.get("RANK", "0")) != 0:
            return  # the other configurations are single-GPU lines
        if args.config == "tcn_bf16":
            (bench_extra.run_tcn_bf16_reference if args.impl == "reference" else bench_extra.run_tcn_bf16)(args)
        elif args.config == "pipeline":
            bench_extra.run_pipeline(args)
        else:
            bench_extra.run_train(args)
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
