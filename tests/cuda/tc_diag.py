#!/usr/bin/env python
"""Diagnostic sweep of the tcgen05 fused row MLP against float64 torch on the CPU (run on the
B200 box).  Every case runs in its own process so that a trapped kernel (sticky CUDA error)
does not hide the cases behind it:

    python tests/cuda/tc_diag.py            # all cases, one line each
    python tests/cuda/tc_diag.py 3          # one case in this process
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

# name, n_rows, stream block widths, projected table rows (0 = none), Linear widths, extras
CASES = [
    ("1 layer 64->64, 128 rows", 128, [64], 0, [64], {}),
    ("1 layer 64->64, 1000 rows", 1000, [64], 0, [64], {}),
    ("1 layer 32->16", 300, [32], 0, [16], {}),
    ("2 layers 64->64->64", 1000, [64], 0, [64, 64], {}),
    ("3 layers 64->64->64->64", 5000, [64], 0, [64, 64, 64], {}),
    ("3 layers, 2 blocks 64+64", 5000, [64, 64], 0, [64, 64, 64], {}),
    ("3 layers, gathered block", 5000, [64], 0, [64, 64, 64], {"gather": True}),
    ("3 layers + 2 projected tables", 20000, [64], 700, [64, 64, 64], {}),
    ("3 layers + projected + out_index + aggr", 20000, [64], 700, [64, 64, 64], {"aggr": True, "scatter": True}),
    ("narrow blocks 5+4 -> 64 -> 64 -> 5", 3000, [5, 4], 0, [64, 64, 5], {}),
    ("narrow 14 -> 64 -> 64 (encoder, relu)", 3000, [14], 0, [64, 64], {"final_relu": True}),
    ("hidden 40: 8+8 -> 40 -> 40 -> 4", 3000, [8, 8], 0, [40, 40, 4], {}),
    ("W head 4x64 + 2 projected -> 64 -> 64 -> 1 sigmoid", 20000, [64, 64, 64, 64], 900, [64, 64, 1], {"sigmoid": True, "scatter": True}),
    ("residual + relu on load", 3000, [64, 64], 0, [64, 64, 64], {"res": True, "relu_in": True}),
    ("big: 1M rows, projected, aggr", 1000000, [64], 100000, [64, 64, 64], {"aggr": True, "scatter": True}),
    ("big W head: 1M rows, 4x64 -> 64 -> 64 -> 1 sigmoid", 1000000, [64, 64, 64, 64], 0, [64, 64, 1], {"sigmoid": True}),
    ("big encoder: 1M rows, 4 -> 64 -> 64 relu", 1000000, [4], 0, [64, 64], {"final_relu": True}),
    ("edge ws: tile in / tile out", 20000, [64], 700, [64, 64, 64], {"aggr": True}),
    ("edge ws: gather in / scatter out, relu", 20077, [64], 700, [64, 64, 64], {"aggr": True, "scatter": True, "gather": True, "relu_in": True}),
    ("edge ws: 100 rows", 100, [64], 30, [64, 64, 64], {"aggr": True, "scatter": True}),
    ("big edge ws: 1M rows, tile in / tile out", 1000000, [64], 100000, [64, 64, 64], {"aggr": True}),
]


def run_case(i: int) -> None:
    import torch

    from gnn_tracking_b200 import _lib, ops
    from gnn_tracking_b200.ops import ACT_NONE, ACT_RELU, ACT_SIGMOID_AFFINE, IMPL_FFMA, IMPL_TCGEN05, Block

    name, n, widths, n_tab, outs, ex = CASES[i]
    gen = torch.Generator().manual_seed(100 + i)
    dev = torch.device("cuda")
    k0 = sum(widths)
    dims = [k0] + outs
    ws = [torch.randn(dims[j + 1], dims[j], generator=gen) / dims[j] ** 0.5 for j in range(len(outs))]
    bs = [torch.randn(dims[j + 1], generator=gen) * 0.1 for j in range(len(outs))]
    srcs = [torch.randn(n, w, generator=gen) for w in widths]
    idx = None
    if ex.get("gather"):
        idx = torch.randint(0, n, (n,), generator=gen)
    relu_in = bool(ex.get("relu_in"))
    # float64 reference
    xin = [s.double() for s in srcs]
    if idx is not None:
        xin[0] = xin[0][idx]
    if relu_in:
        xin = [x.clamp_min(0) for x in xin]
    h = torch.cat(xin, 1) @ ws[0].double().t() + bs[0].double()
    tabs, tidx = [], []
    if n_tab:
        seg = torch.sort(torch.randint(0, n_tab, (n,), generator=gen)).values
        other = torch.randint(0, n_tab, (n,), generator=gen)
        for ind in (seg, other):
            tab = torch.randn(n_tab, outs[0], generator=gen)
            tabs.append(tab)
            tidx.append(ind)
            h = h + tab.double()[ind]
    for j in range(1, len(outs)):
        h = h.clamp_min(0) @ ws[j].double().t() + bs[j].double()
    if ex.get("final_relu"):
        h = h.clamp_min(0)
    if ex.get("sigmoid"):
        h = 0.001 + 0.998 * torch.sigmoid(h)
    res = None
    if ex.get("res"):
        res = torch.randn(n, outs[-1], generator=gen)
        h = 0.6 ** 0.5 * res.double() + 0.4 ** 0.5 * h
    perm = torch.randperm(n, generator=gen) if ex.get("scatter") else None
    ref_out = h
    ref_aggr = None
    if ex.get("aggr"):
        ref_aggr = torch.zeros(n_tab, outs[-1], dtype=torch.float64).index_add_(0, tidx[0], h)

    line = [f"case {i:2d} {name:52s}"]
    for impl, iname in ((IMPL_TCGEN05, "tc"), (IMPL_FFMA, "ffma")):
        try:
            packed = ops.pack_linears([w.to(dev) for w in ws], [b.to(dev) for b in bs], impl, block_widths=widths)
            blocks = [Block(s.to(dev), idx.int().to(dev) if (idx is not None and j == 0) else None, relu_in)
                      for j, s in enumerate(srcs)]
            for tab, ind, srt in zip(tabs, tidx, (True, False)):
                blocks.append(Block(tab.to(dev), ind.int().to(dev), False, projected=True, sorted_index=srt))
            kw = {}
            if ex.get("final_relu"):
                kw["final_act"] = ACT_RELU
            if ex.get("sigmoid"):
                kw.update(final_act=ACT_SIGMOID_AFFINE, act_eps=0.001)
            if res is not None:
                kw.update(res=res.to(dev), res_a=0.6 ** 0.5, res_b=0.4 ** 0.5)
            if perm is not None:
                kw["out_index"] = perm.int().to(dev)
            aggr = None
            if ex.get("aggr"):
                aggr = torch.zeros(n_tab, outs[-1], device=dev)
                rowptr = torch.zeros(n_tab + 1, dtype=torch.int32)
                rowptr[1:] = torch.cumsum(torch.bincount(tidx[0], minlength=n_tab), 0).int()
                kw.update(aggr=aggr, seg_id=tidx[0].int().to(dev), rowptr=rowptr.to(dev))
            for _ in range(2):  # twice: the second launch runs with warm caches and a reused TMEM
                if aggr is not None:
                    aggr.zero_()
                out = ops.fused_mlp(blocks, n, packed, **kw)
            torch.cuda.synchronize()
            got = out.cpu().double()
            if perm is not None:
                got = got[perm]
            scale = max(1.0, float(ref_out.abs().max()))
            err = float((got - ref_out).abs().max()) / scale
            msg = f"{iname}: rel_err {err:.2e}"
            if aggr is not None:
                ea = float((aggr.cpu().double() - ref_aggr).abs().max()) / max(1.0, float(ref_aggr.abs().max()))
                msg += f" aggr {ea:.2e}"
            if iname == "tc":
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                for _ in range(5):
                    ops.fused_mlp(blocks, n, packed, **kw)
                ev1.record()
                torch.cuda.synchronize()
                msg += f" {ev0.elapsed_time(ev1) / 5 * 1e3:.0f}us"
                if os.environ.get("EW_PROF"):
                    L = _lib.lib()
                    L.gtb_debug_tc_profile(2, None)
                    ops.fused_mlp(blocks, n, packed, **kw)
                    torch.cuda.synchronize()
                    buf = (_lib.C.c_longlong * 32)()
                    L.gtb_debug_tc_profile(2, buf)
                    L.gtb_debug_tc_profile(0, None)
                    tiles = max(1, buf[31])
                    names = {0: "c0:pre", 1: "c0:wait e", 2: "c0:body", 3: "e0:pre", 4: "e0:wait Pj", 5: "e0:wait d", 6: "e0:body",
                             7: "e1:pre", 8: "e1:wait d", 9: "e1:body", 10: "e2:pre", 11: "e2:wait d", 12: "e2:body",
                             13: "ag:pre", 14: "ag:wait out", 15: "ag:body", 16: "ag:refill Pj", 17: "prod:loop",
                             18: "prod:wait free", 19: "prod:issue", 24: "mma:loop", 25: "mma:wait a", 26: "mma:issue"}
                    msg += "\n    ew prof (cycles per PAIR of tiles, warp 0 / CTA 0, both contexts; prod / mma: context A; %d pairs): " % tiles + ", ".join(
                        f"{nm} {buf[i] / tiles:.0f}" for i, nm in names.items())
                if os.environ.get("TC_PROF"):
                    L = _lib.lib()
                    L.gtb_debug_tc_profile(1, None)
                    ops.fused_mlp(blocks, n, packed, **kw)
                    torch.cuda.synchronize()
                    buf = (_lib.C.c_longlong * 32)()
                    L.gtb_debug_tc_profile(0, buf)
                    tiles = max(1, buf[31])
                    names = ["prologue", "wait chunk", "convert", "mma issue", "issue item", "pre loads", "mma wait l0",
                             "wait add", "epilogue l0", "mma issue", "issue add", "mma wait l>0", "epilogue l>0",
                             "mma wait out", "epilogue out", "store", "aggregate", "tail"]
                    msg += "\n    prof (cycles per tile of team 0 / CTA 0, %d tiles): " % tiles + ", ".join(
                        f"{nm} {buf[i] / (1 if i == 0 else tiles):.0f}" for i, nm in enumerate(names))
                    msg += f" | total/tile {sum(buf[1:18]) / tiles:.0f}"
            line.append(msg)
        except Exception as e:  # noqa: BLE001
            line.append(f"{iname}: EXC {type(e).__name__}: {str(e)[:150]}")
    flag = _lib.C.c_int(0)
    try:
        _lib.lib().gtb_debug_tc_timeout(_lib.C.byref(flag))
        line.append(f"timeout_flag={flag.value}")
    except Exception:  # noqa: BLE001
        pass
    print(" | ".join(line), flush=True)


def main() -> None:
    if len(sys.argv) > 1:
        run_case(int(sys.argv[1]))
        return
    for i in range(len(CASES)):
        try:
            r = subprocess.run([sys.executable, __file__, str(i)], capture_output=True, text=True, timeout=180)
            out = r.stdout.strip() or f"case {i}: no output; rc={r.returncode}; stderr tail: {r.stderr.strip()[-300:]}"
        except subprocess.TimeoutExpired:
            out = f"case {i}: TIMEOUT (180 s)"
        print(out, flush=True)


if __name__ == "__main__":
    main()
