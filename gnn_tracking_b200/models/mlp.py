"""``MLP`` / ``ResFCNN`` with the reference's constructor arguments and state_dict
names (reference models/mlp.py:18-120), evaluated by the fused row-MLP kernel."""
from __future__ import annotations

import math
import os
from typing import Sequence

import torch
from torch import Tensor, nn

from .. import ops
from ..ops import ACT_NONE, ACT_RELU, Block, PackedMLP


def _stream_dims(linears, widths, projected):
    k0 = sum(w for w, pr in zip(widths, projected) if not pr)
    return [k0] + [lin.out_features for lin in linears], [w for w, pr in zip(widths, projected) if not pr]


class PackedCache:
    """Packed copies of a chain of ``nn.Linear`` layers for one calling pattern (widths of
    the concatenated source blocks, which of them are pre-projected), re-packed when a weight
    changes (optimizer step, ``load_state_dict``, ``.to(device)``)."""

    def __init__(self):
        self._key = None
        self._packed: list[PackedMLP] = []
        self._proj: dict[int, PackedMLP] = {}

    def invalidate(self) -> None:
        """Forget the packed copies (and the backward's, ``autograd._BwdPacks``).  The cache key is
        (storage pointer, ``_version``) per parameter, which sees optimizer steps, ``load_state_dict``
        and ``.to()``; writes through ``param.data`` (``p.data.clamp_()``, EMA / SWA copies, legacy
        optimizers) do NOT bump ``_version`` -- call ``invalidate_packed_weights(model)`` after them."""
        self._key = None
        self._packed, self._proj = [], {}
        if hasattr(self, "bwd"):
            del self.bwd

    @staticmethod
    def _groups(linears, widths, projected, impl):
        """<= 3 Linear layers per launch.  With the tcgen05 tiles a 3-layer chain whose weights
        would leave less than two staging slots in shared memory (the W head: 4 x 64 edge-embedding
        columns in front of the first Linear) runs as 2 + 1 layers instead: one more pass over
        E x H activations, but the gathers of a tile overlap the previous tile's MMAs again."""
        linears = list(linears)
        if len(linears) == 3 and impl != ops.IMPL_FFMA:
            dims3, bw = _stream_dims(linears, widths, projected)
            if 0 < ops.tc_slots(dims3, bw) < 2 and ops.tc_slots(dims3[:3], bw) >= 2:
                return [linears[:2], linears[2:]]
        return [linears[i:i + 3] for i in range(0, len(linears), 3)]

    def get(self, linears: Sequence[nn.Linear], widths: Sequence[int] | None = None,
            projected: Sequence[bool] | None = None, impl: int | None = None):
        """Returns (packed groups of <= 3 Linear layers, {block position: packed projection}).
        The first group's first Linear only keeps the columns of the non-projected blocks;
        each projected block gets its own bias-free single-Linear MLP ``[width -> N0]``."""
        impl = ops.default_impl() if impl is None else impl
        widths = tuple(widths) if widths is not None else (linears[0].in_features,)
        projected = tuple(projected) if projected is not None else (False,) * len(widths)
        key = (impl, widths, projected) + tuple((p.data_ptr(), p._version) for lin in linears for p in lin.parameters())
        if key != self._key:
            groups = self._groups(linears, widths, projected, impl)
            self._packed, self._proj = [], {}
            for gi, g in enumerate(groups):
                ws, bs = [l.weight for l in g], [l.bias for l in g]
                bw = None
                if gi == 0:
                    offs = [0]
                    for w in widths:
                        offs.append(offs[-1] + w)
                    if offs[-1] != g[0].in_features:
                        raise AssertionError(f"blocks of widths {widths} do not match the {g[0].in_features} input features")
                    w0 = g[0].weight.detach()
                    for i, pr in enumerate(projected):
                        if pr:
                            self._proj[i] = ops.pack_linears([w0[:, offs[i]:offs[i + 1]].contiguous()], [None], impl)
                    keep = [i for i, pr in enumerate(projected) if not pr]
                    bw = [widths[i] for i in keep]
                    if len(keep) != len(widths):
                        ws[0] = torch.cat([w0[:, offs[i]:offs[i + 1]] for i in keep], dim=1).contiguous()
                self._packed.append(ops.pack_linears(ws, bs, impl, block_widths=bw))
            self._key = key
        return self._packed, self._proj


def invalidate_packed_weights(module: nn.Module) -> None:
    """Drops every packed-weight cache below ``module``: the next forward re-packs from the live
    parameters.  Needed after in-place writes through ``param.data`` (see ``PackedCache.invalidate``);
    ``.to()`` / ``.cuda()`` / ``load_state_dict`` call it themselves."""
    for m in module.modules():
        for v in vars(m).values():
            for c in (v if isinstance(v, (list, tuple)) else (v,)):
                if isinstance(c, PackedCache):
                    c.invalidate()


class _PackedWeightsMixin:
    """``_apply`` (``.to`` / ``.cuda`` / ``.float``) and ``load_state_dict`` invalidate the packed copies."""

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        invalidate_packed_weights(self)
        return out

    def _hook_state_dict(self) -> None:
        self.register_load_state_dict_post_hook(lambda module, incompatible_keys: invalidate_packed_weights(module))


def autocast_bf16() -> bool:
    """True inside ``torch.autocast("cuda", dtype=torch.bfloat16)``: the reference's mixed-precision mode
    (BASELINE config 3).  Linear layers then take bf16 operands and return bf16, as autocast makes
    ``nn.Linear`` do; see DESIGN.md "bf16 path" for which launches are native in that mode."""
    return torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16


def _run_autocast(linears: Sequence[nn.Linear], blocks: Sequence[Block], n_rows: int, *, final_act: int = ACT_NONE,
                  act_eps: float = 0.0, res: Tensor | None = None, res_a: float = 0.0, res_b: float = 1.0,
                  out_scale: Tensor | None = None, row_scale: Tensor | None = None, out_index: Tensor | None = None,
                  aggr_rows: int | None = None, seg_id: Tensor | None = None, **unused):
    """A Linear / ReLU chain over concatenated column blocks under bf16 autocast, op for op what the
    reference's modules run there (models/mlp.py:59-62, edge_classifier.py:108-117): library GEMMs on
    bf16 operands with bf16 results.  Used for everything around the Interaction-Network edge kernel in
    bf16 mode (encoders, object model, heads); the edge kernel itself is ``ops.in_edge_bf16``."""
    cols = []
    for b in blocks:
        if b.extend is not None or b.projected:
            raise NotImplementedError("bf16 autocast over halo / pre-projected blocks is not implemented")
        t = b.tensor if b.tensor.dim() > 1 else b.tensor.unsqueeze(1)
        if b.index is not None:
            t = t[b.index.long()]
        cols.append(torch.relu(t) if b.relu else t)
    h = cols[0] if len(cols) == 1 else torch.cat(cols, dim=1)
    if row_scale is not None:
        h = h * row_scale.unsqueeze(1)
    for i, lin in enumerate(linears):
        h = nn.functional.linear(h, lin.weight, lin.bias)  # autocast: bf16 x bf16 -> bf16
        if i + 1 < len(linears):
            h = torch.relu(h)
    if final_act == ACT_RELU:
        h = torch.relu(h)
    elif final_act == ops.ACT_SIGMOID_AFFINE:
        h = act_eps + (1.0 - 2.0 * act_eps) * torch.sigmoid(h)
    if res is not None:
        h = res_a * res + res_b * h
    elif res_b != 1.0:
        h = res_b * h
    if out_scale is not None:
        h = h * out_scale.to(h.dtype)  # the reference's product with the latent scale stays bf16 under autocast
    aggr = None
    if aggr_rows is not None:
        aggr = torch.zeros((aggr_rows, h.size(1)), dtype=torch.float32, device=h.device).index_add_(0, seg_id.long(), h.float())
    if out_index is not None:
        out = torch.empty_like(h)
        out[out_index.long()] = h
        h = out
    return h if aggr_rows is None else (h, aggr)


def _run_nograd(cache: PackedCache, linears: Sequence[nn.Linear], blocks: Sequence[Block], n_rows: int,
                *, final_act: int = ACT_NONE, tables: dict | None = None, save_hidden: list | None = None,
                **epilogue) -> Tensor | None:
    """Linear/ReLU chain over concatenated column blocks; <= 3 Linear layers per
    fused launch, longer chains are split with the intermediate kept in HBM.

    Gathered blocks of a smaller table (``x[dst]``, ``x[src]`` with E >> N) are multiplied by
    their columns of the first Linear once per table row and the gathered products added
    behind the first Linear -- same sum, 2*Dn*H fewer multiply-adds per edge."""
    blocks = list(blocks)
    widths = [b.tensor.size(1) if b.tensor.dim() > 1 else 1 for b in blocks]
    n0 = linears[0].out_features
    can_project = min(len(linears), 3) >= 2 and n0 % 4 == 0
    projected = [can_project and b.index is not None and 2 * b.tensor.size(0) <= n_rows and not b.projected
                 for b in blocks]
    if all(projected):
        projected[-1] = False
    packed, proj = cache.get(linears, widths, projected)
    cur: list = [None] * len(blocks)
    pending = []
    # ``tables``: {block position: its pre-projected table}, already computed by the producer of the block's
    # tensor (``ops.in_node_fused``: the previous layer's node kernel) -- no projection launch here
    for i, table in (tables or {}).items():
        if not projected[i]:
            raise AssertionError("a pre-projected table was supplied for a block this call does not project")
        # (a block that crosses a halo exchange comes with its table already extended by the owned + halo rows)
        cur[i] = Block(table, blocks[i].index, False, projected=True, sorted_index=blocks[i].sorted_index)
    # blocks that cross the halo exchange of a node-partitioned graph first: their transfer runs under the
    # launches of the other blocks' projections
    for i in sorted(range(len(blocks)), key=lambda i: blocks[i].extend is None):
        b = blocks[i]
        if cur[i] is not None:
            continue
        ext = b.extend
        start = getattr(ext, "__self__", None).start if hasattr(getattr(ext, "__self__", None), "start") else None
        if projected[i]:
            out = None
            if start is not None:  # the projection writes the owned rows of the extended table in place
                out = ext.__self__.buffer(proj[i].dims[-1], b.tensor)
            table = ops.fused_mlp([Block(b.tensor, None, b.relu)], b.tensor.size(0), proj[i], out=out)
            if ext is not None:  # halo rows of the projected table from their owners
                if start is not None:
                    pending.append((i, start(table), b))
                    continue
                table = ext(table)
            cur[i] = Block(table, b.index, False, projected=True, sorted_index=b.sorted_index)
        elif ext is not None:
            cur[i] = Block(ext(b.tensor), b.index, b.relu, sorted_index=b.sorted_index)
        else:
            cur[i] = b
    for i, handle, b in pending:
        cur[i] = Block(b.extend.__self__.finish(handle), b.index, False, projected=True, sorted_index=b.sorted_index)
    for i, p in enumerate(packed):
        if i + 1 < len(packed):
            h = ops.fused_mlp(cur, n_rows, p, final_act=ACT_RELU)
            cur = [Block(h)]
        else:
            if save_hidden is not None and len(packed) == 1 and p.n_layers == 3:
                epilogue = dict(epilogue, save_hidden=save_hidden)  # see ops.fused_mlp
            return ops.fused_mlp(cur, n_rows, p, final_act=final_act, **epilogue)
    raise AssertionError("unreachable")


def projection_pattern(widths: Sequence[int], n_projected: int = 2) -> tuple[tuple[int, ...], tuple[bool, ...]]:
    """The calling pattern ``_run_nograd`` derives for a consumer whose first ``n_projected`` blocks are gathered
    rows of a small table (x[dst], x[src]) and whose other blocks are streamed: the key of its ``PackedCache``."""
    widths = tuple(widths)
    return widths, (True,) * n_projected + (False,) * (len(widths) - n_projected)


def projection_packs(mlp: "MLP", widths: Sequence[int]):
    """Packed single-Linear projections of the first two blocks of ``mlp``'s first Linear (the node column
    blocks a relational model / the W head gathers), or None when they are not 64 -> 64 tensor-core packs."""
    if len(mlp.linears) < 2:
        return None
    w, pr = projection_pattern(widths)
    packed, proj = mlp._cache.get(mlp.linears, w, pr)
    pa, pb = proj.get(0), proj.get(1)
    if pa is None or pb is None or any(p.impl != ops.IMPL_TCGEN05 or p.dims != (64, 64) for p in (pa, pb)):
        return None
    return pa, pb


def run_linears(cache: PackedCache, linears: Sequence[nn.Linear], blocks: Sequence[Block], n_rows: int,
                *, final_act: int = ACT_NONE, aggr_rows: int | None = None, **epilogue):
    """``_run_nograd`` plus autograd (``autograd.FusedMLPFunction``, recompute-based backward through
    the same kernels) when a block, a weight or the residual requires a gradient.  With
    ``aggr_rows`` the per-destination sum is returned as well: ``(out, aggr)``."""
    if autocast_bf16():
        if epilogue.get("tables"):
            raise AssertionError("pre-projected tables belong to the fp32 no-grad path")
        return _run_autocast(linears, blocks, n_rows, final_act=final_act, aggr_rows=aggr_rows, **epilogue)
    res = epilogue.get("res")
    needs_grad = torch.is_grad_enabled() and (
        any(b.tensor.requires_grad for b in blocks) or any(p.requires_grad for lin in linears for p in lin.parameters())
        or (res is not None and res.requires_grad))
    if not needs_grad:
        if aggr_rows is None:
            return _run_nograd(cache, linears, blocks, n_rows, final_act=final_act, **epilogue)
        dev = blocks[0].tensor.device
        aggr = torch.zeros((aggr_rows, linears[-1].out_features), dtype=torch.float32, device=dev)
        out = _run_nograd(cache, linears, blocks, n_rows, final_act=final_act, aggr=aggr, **epilogue)
        return out, aggr
    from ..autograd import _BwdPacks, fused_mlp_autograd
    if any(b.extend is not None for b in blocks):
        # node-partitioned graph under autograd: the raw rows cross the halo through a differentiable
        # exchange (its backward is the reverse all-to-all-v), the rest is the single-GPU backward
        from ..partition import halo_extend
        blocks = [b if b.extend is None else
                  Block(halo_extend(b.tensor, b.extend.__self__), b.index, b.relu, sorted_index=b.sorted_index)
                  for b in blocks]
    bad = [k for k in ("row_scale", "out_scale", "out", "gate", "aggr", "tables") if epilogue.get(k) is not None]
    if bad or epilogue.get("want_out") is False:
        raise NotImplementedError(f"backward with the epilogue options {bad or ['want_out=False']} is not implemented")
    if not hasattr(cache, "bwd"):
        cache.bwd = _BwdPacks()

    def runner(cfg, block_tensors, res_t, hidden_out=None):
        bl = [Block(t, m[0], m[1], sorted_index=s, unique_index=m[2]) for t, m, s in zip(block_tensors, cfg["blocks"], cfg["sorted"])]
        dev = block_tensors[0].device
        aggr = None
        kw = {}
        if cfg["aggr_rows"] is not None:
            aggr = torch.zeros((cfg["aggr_rows"], linears[-1].out_features), dtype=torch.float32, device=dev)
            kw.update(aggr=aggr, seg_id=cfg["seg_id"], rowptr=cfg["rowptr"])
        if res_t is not None:
            kw.update(res=res_t, res_a=cfg["res_a"])
        out = _run_nograd(cache, linears, bl, cfg["n_rows"], final_act=cfg["final_act"], act_eps=cfg["act_eps"],
                          res_b=cfg["res_b"], out_index=cfg["out_index"], save_hidden=hidden_out, **kw)
        return out, aggr

    return fused_mlp_autograd(runner, cache.bwd, linears, blocks, n_rows, final_act=final_act,
                              act_eps=epilogue.get("act_eps", 0.0), res=res, res_a=epilogue.get("res_a", 0.0),
                              res_b=epilogue.get("res_b", 1.0), out_index=epilogue.get("out_index"), aggr_rows=aggr_rows,
                              seg_id=epilogue.get("seg_id"), rowptr=epilogue.get("rowptr"))


class MLP(_PackedWeightsMixin, nn.Module):
    def __init__(self, input_size: int, output_size: int, hidden_dim: int | None, L: int = 3, *,
                 bias: bool = True, include_last_activation: bool = False):
        """Linear/ReLU chain: 1 input layer, ``L - 2`` hidden layers, 1 output layer.
        ``hidden_dim=None`` picks ``max(input_size, output_size)`` (reference
        mlp.py:42-43).  Parameters live at ``layers.{0,2,4,...}`` as in the reference."""
        super().__init__()
        if hidden_dim is None:
            hidden_dim = max(input_size, output_size)
        widths = [input_size] + [hidden_dim] * max(L - 1, 1) + [output_size]
        mods: list[nn.Module] = []
        for i in range(len(widths) - 1):
            if i:
                mods.append(nn.ReLU())
            mods.append(nn.Linear(widths[i], widths[i + 1], bias=bias))
        if include_last_activation:
            mods.append(nn.ReLU())
        self.layers = nn.ModuleList(mods)
        self._last_act = include_last_activation
        self._cache = PackedCache()
        self._hook_state_dict()

    @property
    def linears(self) -> list[nn.Linear]:
        return [m for m in self.layers if isinstance(m, nn.Linear)]

    @property
    def in_features(self) -> int:
        return self.linears[0].in_features

    @property
    def out_features(self) -> int:
        return self.linears[-1].out_features

    def reset_parameters(self) -> None:
        for m in self.layers:
            if hasattr(m, "reset_parameters"):
                m.reset_parameters()

    def forward_blocks(self, blocks: Sequence[Block], n_rows: int, *, final_act: int | None = None,
                       **epilogue) -> Tensor | None:
        if final_act is None:
            final_act = ACT_RELU if self._last_act else ACT_NONE
        return run_linears(self._cache, self.linears, blocks, n_rows, final_act=final_act, **epilogue)

    def forward(self, x: Tensor) -> Tensor:
        return self.forward_blocks([Block(x)], x.size(0))

    # ---- the 4 -> 64 -> 64 shape of the edge classifier's edge encoder: one warp-specialised launch
    def k4_ok(self, x: Tensor) -> bool:
        lin = self.linears
        if (len(lin) != 2 or lin[0].in_features != 4 or lin[0].out_features != 64 or lin[1].out_features != 64
                or os.environ.get("GTB_NO_ENC_WS") or autocast_bf16() or ops.default_impl() == ops.IMPL_FFMA
                or x.dim() != 2 or x.size(1) != 4 or x.dtype != torch.float32):
            return False
        return not (torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())))

    def forward_k4(self, x: Tensor, index: Tensor | None, n_rows: int, *, final_relu: bool) -> Tensor:
        """``act(MLP(x[index]))`` for the shape ``k4_ok`` accepts (``ops.edge_encoder``): the K = 4 first Linear
        on the CUDA cores, the 64 x 64 Linear on the tensor core, rows gathered by TMA."""
        if "_cache_k4" not in self.__dict__:
            self.__dict__["_cache_k4"] = PackedCache()
        lin = self.linears
        packed = self._cache_k4.get([lin[1]], impl=ops.IMPL_TCGEN05)[0][0]
        return ops.edge_encoder(x, index, n_rows, lin[0].weight, lin[0].bias, packed, final_relu or self._last_act)


class ResFCNN(_PackedWeightsMixin, nn.Module):
    def __init__(self, *, in_dim: int, hidden_dim: int, out_dim: int, depth: int, alpha: float = 0.6,
                 bias: bool = True):
        """L2-normalise rows -> encoder -> (depth-1) residual hidden layers
        ``sqrt(a) x + sqrt(1-a) layer(relu(x))`` -> decoder(relu(.)); parameter names
        ``_encoder``, ``_layers.{i}``, ``_decoder`` and the normal initialisation of
        reference mlp.py:95-113."""
        super().__init__()
        if depth < 1:
            raise ValueError("Depth must be at least 1")
        self._encoder = nn.Linear(in_dim, hidden_dim, bias=bias)
        self._decoder = nn.Linear(hidden_dim, out_dim, bias=bias)
        self._layers = nn.ModuleList([nn.Linear(hidden_dim, hidden_dim, bias=bias) for _ in range(depth - 1)])
        self._init(self._encoder, 1 / in_dim)
        for lay in self._layers:
            self._init(lay, 2 / hidden_dim)
        self._init(self._decoder, 2 / hidden_dim)
        self._alpha = alpha
        self._cache = PackedCache()
        self._layer_caches = [PackedCache() for _ in range(depth)]
        self._hook_state_dict()

    @staticmethod
    def _init(layer: nn.Linear, var: float) -> None:
        layer.reset_parameters()  # keeps the RNG stream aligned with the reference (mlp.py:108-111)
        for p in layer.parameters():
            nn.init.normal_(p.data, mean=0.0, std=math.sqrt(var))

    def forward_blocks(self, blocks: Sequence[Block], n_rows: int, *, final_act: int = ACT_NONE) -> Tensor:
        if autocast_bf16():  # reference mlp.py:115-120 under autocast: library GEMMs, bf16 results
            cols = [(torch.relu(b.tensor) if b.relu else b.tensor) if b.index is None else
                    (torch.relu(b.tensor[b.index.long()]) if b.relu else b.tensor[b.index.long()]) for b in blocks]
            x = nn.functional.normalize(cols[0] if len(cols) == 1 else torch.cat(cols, 1), p=2.0, dim=1, eps=1e-12)
            x = self._encoder(x)
            a, b2 = math.sqrt(self._alpha), math.sqrt(1 - self._alpha)
            for lay in self._layers:
                x = a * x + b2 * lay(torch.relu(x))
            x = self._decoder(torch.relu(x))
            return torch.relu(x) if final_act == ACT_RELU else x
        if torch.is_grad_enabled() and (any(b.tensor.requires_grad for b in blocks)
                                        or any(p.requires_grad for p in self.parameters())):
            return self._forward_grad(blocks, n_rows, final_act)
        inv = ops.rows_inv_l2norm(blocks, n_rows, 1e-12)
        if len(self._layers) == 0:
            p = self._cache.get([self._encoder, self._decoder], [b.tensor.size(1) for b in blocks])[0][0]
            return ops.fused_mlp(blocks, n_rows, p, row_scale=inv, final_act=final_act)
        x = ops.fused_mlp(blocks, n_rows, self._layer_caches[0].get([self._encoder], [b.tensor.size(1) for b in blocks])[0][0], row_scale=inv)
        a, b = math.sqrt(self._alpha), math.sqrt(1 - self._alpha)
        for lay, cache in zip(self._layers, self._layer_caches[1:]):
            x = ops.fused_mlp([Block(x, relu=True)], n_rows, cache.get([lay])[0][0], res=x, res_a=a, res_b=b)
        return ops.fused_mlp([Block(x, relu=True)], n_rows, self._cache.get([self._decoder])[0][0],
                             final_act=final_act)

    def _forward_grad(self, blocks: Sequence[Block], n_rows: int, final_act: int) -> Tensor:
        """Differentiable variant: the normalised input is materialised once (its gradient needs it),
        every Linear then runs through ``run_linears`` and its recompute-based backward."""
        from ..autograd import GatherRows, L2NormalizeRows
        cols = []
        for b in blocks:
            t = b.tensor if b.tensor.dim() > 1 else b.tensor.unsqueeze(1)
            if b.extend is not None or b.projected:
                raise NotImplementedError("ResFCNN backward over halo / pre-projected blocks is not implemented")
            t = t if b.index is None else GatherRows.apply(t, b.index)
            cols.append(torch.relu(t) if b.relu else t)
        x = L2NormalizeRows.apply(cols[0] if len(cols) == 1 else torch.cat(cols, 1), 1e-12)
        if len(self._layers) == 0:
            return run_linears(self._cache, [self._encoder, self._decoder], [Block(x)], n_rows, final_act=final_act)
        x = run_linears(self._layer_caches[0], [self._encoder], [Block(x)], n_rows)
        a, b = math.sqrt(self._alpha), math.sqrt(1 - self._alpha)
        for lay, cache in zip(self._layers, self._layer_caches[1:]):
            x = run_linears(cache, [lay], [Block(x, relu=True)], n_rows, res=x, res_a=a, res_b=b)
        return run_linears(self._cache, [self._decoder], [Block(x, relu=True)], n_rows, final_act=final_act)

    def forward(self, x: Tensor, **ignore) -> Tensor:
        return self.forward_blocks([Block(x)], x.size(0))
