"""Backward pass of the fused row MLP (``ops.fused_mlp`` behind ``models.mlp.run_linears``).

Recompute-based: the forward keeps only its inputs and its output.  The backward recomputes the
hidden activations with the forward tiles (prefix launches), propagates the activation gradients
``dX = dY W`` through the SAME fused kernel with transposed weights (ReLU masks applied by its
``gate`` epilogue) and gets weight / bias gradients from ``gtb_rows_atb_f32``; gathers turn into
``gtb_rows_scatter_add_f32``.  All GEMM-shaped work runs in ``libgtb200.so``; torch is used for
O(rows x 64) elementwise glue (scalar scaling, masks, adding two gradient contributions).

It mirrors what torch autograd derives for the reference's op chain (models/mlp.py:59-62,
models/interaction_network.py:67-103, models/resin.py:17-42, models/edge_classifier.py:108-117);
``tests/test_gpu_backward.py`` checks every gradient against autograd through the CPU oracle.
"""
from __future__ import annotations

import os
from typing import Sequence

import torch
from torch import Tensor

from . import ops
from .ops import ACT_RELU, ACT_SIGMOID_AFFINE, Block


class _BwdPacks:
    """Packed weights the backward needs, cached per weight version next to the forward packs."""

    def __init__(self):
        self._key = None
        self._d: dict = {}

    def get(self, linears, name, make, sig=()):
        """``sig``: the calling pattern the packs depend on besides the weights (block widths, which blocks
        are gathered / activated on load): the same MLP called with another block layout re-packs."""
        key = (sig,) + tuple((p.data_ptr(), p._version) for lin in linears for p in lin.parameters())
        if key != self._key:
            self._key, self._d = key, {}
        if name not in self._d:
            self._d[name] = make()
        return self._d[name]


def _gather_rows(t: Tensor, index: Tensor | None) -> Tensor:
    return t if index is None else ops.rows_gather(t.contiguous(), index)


class FusedMLPFunction(torch.autograd.Function):
    """``out (, aggr) = fused_mlp(blocks, linears, epilogue)`` with gradients for the block tensors,
    the Linear weights / biases and the residual."""

    @staticmethod
    def forward(ctx, cfg: dict, *tensors: Tensor):
        nb, nl = len(cfg["blocks"]), len(cfg["linears"])
        hidden: list = []  # the dedicated 64-wide kernels hand out their two hidden activations: no recompute
        with torch.no_grad():
            out, aggr = cfg["runner"](cfg, tensors[:nb], tensors[-1] if cfg["has_res"] else None,
                                      hidden_out=hidden if nl == 3 and not os.environ.get("GTB_NO_SAVE_HIDDEN") else None)
        ctx.cfg = cfg
        ctx.save_for_backward(*tensors, out, *hidden)
        ctx.nb, ctx.nl, ctx.n_hidden = nb, nl, len(hidden)
        if aggr is None:
            return out
        return out, aggr

    @staticmethod
    def backward(ctx, g_out, g_aggr=None):
        cfg, nb, nl = ctx.cfg, ctx.nb, ctx.nl
        saved = ctx.saved_tensors
        saved_hidden = list(saved[len(saved) - ctx.n_hidden:]) if ctx.n_hidden else []
        saved = saved[:len(saved) - ctx.n_hidden]
        tensors, out = saved[:-1], saved[-1]
        blocks_t = tensors[:nb]
        metas = cfg["blocks"]            # (index, relu, unique_index)
        linears = cfg["linears"]
        weights = [lin.weight.detach() for lin in linears]
        n_rows, dev = cfg["n_rows"], out.device
        packs: _BwdPacks = cfg["bwd_packs"]
        dims = [weights[0].size(1)] + [w.size(0) for w in weights]

        # ---- gradient w.r.t. the stored rows, back in launch-row order
        if g_out is None:
            dy = torch.zeros((n_rows, dims[-1]), dtype=torch.float32, device=dev)
        else:
            dy = _gather_rows(g_out.reshape(-1, dims[-1]).to(torch.float32), cfg["out_index"])
        if g_aggr is not None and cfg["seg_id"] is not None:
            if dy is g_out or dy.data_ptr() == g_out.data_ptr():
                dy = dy.clone()  # the incoming gradient is not ours to write into
            dy = ops.rows_gather_add(g_aggr.contiguous(), cfg["seg_id"], dy.contiguous())
        grads: list = [None] * len(tensors)
        if cfg["has_res"]:
            if ctx.needs_input_grad[1 + len(tensors) - 1]:
                grads[-1] = dy * cfg["res_a"]
            dy = dy * cfg["res_b"]
        elif cfg["res_b"] != 1.0:
            dy = dy * cfg["res_b"]
        if cfg["final_act"] == ACT_RELU:
            a_rows = _gather_rows(out.reshape(-1, dims[-1]), cfg["out_index"])
            dy = torch.ops.aten.threshold_backward(dy, a_rows, 0.0)  # dy where the activation was positive, else 0
        elif cfg["final_act"] == ACT_SIGMOID_AFFINE:
            eps = cfg["act_eps"]
            s = (_gather_rows(out.reshape(-1, dims[-1]), cfg["out_index"]) - eps) / (1.0 - 2.0 * eps)
            dy = dy * ((1.0 - 2.0 * eps) * s * (1.0 - s))
        dy = dy.contiguous()

        # ---- recompute the hidden activations (post-ReLU) in launch-row order
        fwd_blocks = [Block(t, m[0], m[1]) for t, m in zip(blocks_t, metas)]
        widths = [t.size(1) if t.dim() > 1 else 1 for t in blocks_t]
        sig = (tuple(widths), tuple((m[0] is not None, bool(m[1])) for m in metas))
        hidden = saved_hidden  # [h0, h1] from the forward launch, or recomputed below
        if nl >= 2 and not hidden:
            p0 = packs.get(linears, "prefix0", lambda: ops.pack_linears([weights[0]], [linears[0].bias], ops.default_impl(),
                                                                        block_widths=widths), sig)
            hidden.append(ops.fused_mlp(fwd_blocks, n_rows, p0, final_act=ACT_RELU))
            for l in range(1, nl - 1):
                pl = packs.get(linears, f"layer{l}", lambda l=l: ops.pack_linears([weights[l]], [linears[l].bias], ops.default_impl()), sig)
                hidden.append(ops.fused_mlp([Block(hidden[-1])], n_rows, pl, final_act=ACT_RELU))

        # ---- layers L-1 .. 1: weight / bias gradients, then dX = dY W gated by the ReLU mask
        dz = dy
        for l in range(nl - 1, 0, -1):
            a_in = hidden[l - 1]
            gw = torch.zeros((dims[l], dims[l + 1]), dtype=torch.float32, device=dev)
            gb = torch.zeros(dims[l + 1], dtype=torch.float32, device=dev) if linears[l].bias is not None else None
            ops.rows_atb(a_in, dz, gw, colsum=gb)
            grads[nb + l] = gw.t()
            if gb is not None:
                grads[nb + nl + l] = gb
            pt = packs.get(linears, f"layerT{l}", lambda l=l: ops.pack_linears([weights[l].t().contiguous()], [None], ops.default_impl()), sig)
            dz = ops.fused_mlp([Block(dz)], n_rows, pt, gate=a_in)

        # ---- first Linear: per source block
        gw0 = torch.zeros((dims[0], dims[1]), dtype=torch.float32, device=dev)
        gb0 = torch.zeros(dims[1], dtype=torch.float32, device=dev) if linears[0].bias is not None else None
        off = 0
        bias_done = gb0 is None
        for i, (t, (index, relu, unique)) in enumerate(zip(blocks_t, metas)):
            w = widths[i]
            t2 = t if t.dim() > 1 else t.unsqueeze(1)
            small_table = index is not None and 2 * t2.size(0) <= n_rows
            need_t = ctx.needs_input_grad[1 + i]
            wslice_t = packs.get(linears, f"block{i}T", lambda off=off, w=w: ops.pack_linears(
                [weights[0][:, off:off + w].t().contiguous()], [None], ops.default_impl()), sig) if need_t else None
            if small_table:
                # gather of a small table: fold the row gradients onto the table first
                dp = torch.zeros((t2.size(0), dims[1]), dtype=torch.float32, device=dev)
                ops.rows_scatter_add(dz, index, dp)
                ops.rows_atb(t2, dp, gw0[off:off + w], a_relu=relu)
                if need_t:
                    gt = ops.fused_mlp([Block(dp)], t2.size(0), wslice_t, gate=t2 if relu else None)
                    grads[i] = gt if grads[i] is None else grads[i] + gt
            else:
                ops.rows_atb(t2, dz, gw0[off:off + w], a_index=index, a_relu=relu, colsum=None if bias_done else gb0)
                bias_done = True
                if need_t:
                    if index is None:
                        gt = ops.fused_mlp([Block(dz)], n_rows, wslice_t, gate=t2 if relu else None)
                    elif unique and t2.size(0) == n_rows:
                        gt = ops.fused_mlp([Block(dz)], n_rows, wslice_t, out_index=index)
                        if relu:
                            gt = torch.ops.aten.threshold_backward(gt, t2, 0.0)
                    else:
                        rows = ops.fused_mlp([Block(dz)], n_rows, wslice_t)
                        gt = torch.zeros_like(t2)
                        ops.rows_scatter_add(rows, index, gt)
                        if relu:
                            gt = torch.ops.aten.threshold_backward(gt, t2, 0.0)
                    grads[i] = gt if grads[i] is None else grads[i] + gt
            off += w
        if not bias_done:  # every block was a small table: the bias gradient is the plain column sum
            gb0 += dz.sum(0)
        grads[nb] = gw0.t()
        if gb0 is not None:
            grads[nb + nl] = gb0
        for i, t in enumerate(blocks_t):
            if grads[i] is not None and t.dim() == 1:
                grads[i] = grads[i].squeeze(1)
        return (None, *grads)


def fused_mlp_autograd(runner, cache_bwd: _BwdPacks, linears: Sequence[torch.nn.Linear], blocks: Sequence[Block],
                       n_rows: int, *, final_act: int, act_eps: float = 0.0, res: Tensor | None = None,
                       res_a: float = 0.0, res_b: float = 1.0, out_index: Tensor | None = None,
                       aggr_rows: int | None = None, seg_id: Tensor | None = None, rowptr: Tensor | None = None):
    """Differentiable ``run_linears``.  ``runner(cfg, block_tensors, res) -> (out, aggr)`` is the
    no-grad implementation.  Returns ``out`` or ``(out, aggr)``."""
    if len(linears) > 3:
        raise NotImplementedError("backward of Linear chains longer than 3 layers is not implemented")
    for b in blocks:
        if b.extend is not None:
            raise NotImplementedError("backward through the halo exchange of a node-partitioned graph is not implemented")
        if b.projected:
            raise NotImplementedError("backward of caller-projected blocks is not implemented")
    cfg = {
        "blocks": [(b.index, bool(b.relu), bool(getattr(b, "unique_index", False))) for b in blocks],
        "sorted": [bool(b.sorted_index) for b in blocks],
        "linears": list(linears), "n_rows": n_rows, "final_act": final_act, "act_eps": act_eps,
        "has_res": res is not None, "res_a": float(res_a), "res_b": float(res_b), "out_index": out_index,
        "aggr_rows": aggr_rows, "seg_id": seg_id, "rowptr": rowptr, "runner": runner, "bwd_packs": cache_bwd,
    }
    tensors = [b.tensor for b in blocks] + [lin.weight for lin in linears]
    tensors += [lin.bias if lin.bias is not None else torch.zeros(0, device=linears[0].weight.device) for lin in linears]
    if res is not None:
        tensors.append(res)
    return FusedMLPFunction.apply(cfg, *tensors)


class GatherRows(torch.autograd.Function):
    """``table[index]`` (``gtb_rows_gather_f32``); the gradient is the scatter-add of the row
    gradients (``gtb_rows_scatter_add_f32``)."""

    @staticmethod
    def forward(ctx, table: Tensor, index: Tensor):
        ctx.save_for_backward(index)
        ctx.n = table.size(0)
        return ops.rows_gather(table.contiguous(), index)

    @staticmethod
    def backward(ctx, g):
        (index,) = ctx.saved_tensors
        out = torch.zeros((ctx.n, g.size(1)), dtype=torch.float32, device=g.device)
        ops.rows_scatter_add(g.contiguous(), index, out)
        return out, None


class L2NormalizeRows(torch.autograd.Function):
    """``F.normalize(x, dim=1)`` of ``ResFCNN.forward`` (models/mlp.py:115-116) with the row norms
    from ``gtb_rows_inv_l2norm_f32``; ``dx = (g - y <y, g>) / max(|x|, eps)``."""

    @staticmethod
    def forward(ctx, x: Tensor, eps: float):
        inv = ops.rows_inv_l2norm([Block(x)], x.size(0), eps)
        y = x * inv.unsqueeze(1)
        ctx.save_for_backward(y, inv)
        return y

    @staticmethod
    def backward(ctx, g):
        y, inv = ctx.saved_tensors
        return (g - y * (y * g).sum(1, keepdim=True)) * inv.unsqueeze(1), None
