"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Interaction-Network hot path.

A functional (state_dict driven) restatement, in plain torch CPU ops, of the
reference's IN / ResIN / edge-classifier / GraphTCN forward.  It exists so the
CUDA path can be checked on a box where ``/root/reference`` is absent.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU baseline legs may
import it; the product package ``gnn_tracking_b200`` never does.

Parity status: PINNED -- ``tests/test_oracle_vs_golden.py`` checks every function
here against golden vectors produced by the reference's own classes
(``tests/golden/make_golden.py``, run where ``/root/reference`` exists) and,
when the reference is importable, against the reference directly.  The
reference's own tests hold no numerical pin for the IN/EC forward
(SURVEY.md section 4), so the pin is "outputs of the reference itself run here".

All ``file:line`` citations are relative to ``/root/reference/src/gnn_tracking``.
Weights use the reference's state_dict names and ``nn.Linear`` ``[out, in]`` layout.
"""
from __future__ import annotations

import math
from itertools import pairwise

import torch
from torch import Tensor


# ----------------------------------------------------------------------------- MLP
def mlp(x: Tensor, sd: dict, prefix: str) -> Tensor:
    """``MLP.forward`` models/mlp.py:59-62 with the layer list built at :44-51:
    Linear at even indices, ReLU in between, no activation after the last
    Linear.  Bias is used when present in the state_dict."""
    idx = 0
    lins = []
    while f"{prefix}layers.{idx}.weight" in sd:
        lins.append(idx)
        idx += 2
    for n, i in enumerate(lins):
        w = sd[f"{prefix}layers.{i}.weight"]
        b = sd.get(f"{prefix}layers.{i}.bias")
        x = torch.nn.functional.linear(x, w, b)  # nn.Linear: under torch.autocast the bias joins the bf16 GEMM
        if n + 1 < len(lins):
            x = torch.relu(x)
    return x


def res_fcnn(x: Tensor, sd: dict, prefix: str, alpha: float) -> Tensor:
    """``ResFCNN.forward`` models/mlp.py:115-120: L2-normalise rows (eps 1e-12),
    encoder, residual hidden layers, decoder(relu(.))."""
    nrm = x.norm(p=2, dim=1, keepdim=True).clamp_min(1e-12)
    x = x / nrm
    x = _linear(x, sd, f"{prefix}_encoder.")
    i = 0
    while f"{prefix}_layers.{i}.weight" in sd:
        x = math.sqrt(alpha) * x + math.sqrt(1 - alpha) * _linear(torch.relu(x), sd, f"{prefix}_layers.{i}.")
        i += 1
    return _linear(torch.relu(x), sd, f"{prefix}_decoder.")


def _linear(x, sd, prefix):
    return torch.nn.functional.linear(x, sd[prefix + "weight"], sd.get(prefix + "bias"))


# ------------------------------------------------------------ Interaction network
def interaction_network(x: Tensor, edge_index: Tensor, edge_attr: Tensor, sd: dict, prefix: str):
    """``InteractionNetwork.forward`` models/interaction_network.py:54-72.

    propagate (PyG >= 2.3, see oracle/shims.py): x_j = x[edge_index[0]] (source),
    x_i = x[edge_index[1]] (target); message :75-89 = relational MLP over
    cat[x_i, x_j, edge_attr]; sum aggregation onto edge_index[1]; update :92-103 =
    object MLP over cat[x, aggr].  Returns (x_tilde, e_tilde)."""
    src, dst = edge_index[0], edge_index[1]
    x_i = x.index_select(0, dst)
    x_j = x.index_select(0, src)
    m = torch.cat([x_i, x_j, edge_attr], dim=1)
    e_tilde = mlp(m, sd, prefix + "relational_model.")
    aggr = e_tilde.new_zeros(x.size(0), e_tilde.size(1))
    aggr.scatter_add_(0, dst.view(-1, 1).expand_as(e_tilde), e_tilde)
    x_tilde = mlp(torch.cat([x, aggr], dim=1), sd, prefix + "object_model.")
    return x_tilde, e_tilde


def sqconvex(delta: Tensor, residue: Tensor | None, alpha: float) -> Tensor:
    """``sqconvex_combination`` models/resin.py:17-42."""
    if residue is None or math.isclose(alpha, 0.0):
        return delta
    return math.sqrt(alpha) * residue + math.sqrt(1 - alpha) * delta


def _n_layers(sd, prefix):
    n = 0
    while f"{prefix}layers.{n}.relational_model.layers.0.weight" in sd:
        n += 1
    return n


def resin(x, edge_index, edge_attr, sd, prefix, *, alpha=0.5, residual_type="skip1", collect=False,
          connect_to=1, add_bn=False, bn_training=False):
    """``ResIN.forward`` models/resin.py:292-295 dispatching to the residual
    networks :99-114 (skip1), :153-175 (skip2, overlapping ``pairwise`` pairs
    reproduced literally), :197-216 (skip_top).  ``prefix`` points at
    ``...network.``.  Returns (x, edge_attr, list_of_edge_attrs | None)."""
    L = _n_layers(sd, prefix)
    edge_attrs = [edge_attr] if collect else None
    if residual_type == "skip1":
        for i in range(L):
            xin, ein = (x, edge_attr) if i == 0 else (torch.relu(x), torch.relu(edge_attr))
            dx, edge_attr = interaction_network(xin, edge_index, ein, sd, f"{prefix}layers.{i}.")
            x = sqconvex(dx, x, alpha)
            if collect:
                edge_attrs.append(edge_attr)
    elif residual_type == "skip2":
        def bn(kind, i, t):  # nn.BatchNorm1d of resin.py:141-151 (eps 1e-5; batch statistics when `training`)
            if not add_bn:
                return t
            pre = f"{prefix}_{kind}_batch_norms.{i}."
            return torch.nn.functional.batch_norm(t, sd[pre + "running_mean"].clone(), sd[pre + "running_var"].clone(),
                                                  sd[pre + "weight"], sd[pre + "bias"], training=bn_training, momentum=0.1, eps=1e-5)
        for i0, i1 in pairwise(range(L)):
            xn, en = bn("node", i0, x), bn("edge", i0, edge_attr)
            xin, ein = (xn, en) if i0 == 0 else (torch.relu(xn), torch.relu(en))
            hx, he = interaction_network(xin, edge_index, ein, sd, f"{prefix}layers.{i0}.")
            dx, edge_attr = interaction_network(torch.relu(bn("node", i1, hx)), edge_index, torch.relu(bn("edge", i1, he)), sd,
                                                f"{prefix}layers.{i1}.")
            x = sqconvex(dx, x, alpha)
            if collect:
                edge_attrs.append(edge_attr)
    elif residual_type == "skip_top":
        x_res = None
        for i in range(L):
            if i == connect_to:
                x_res = x
            xin, ein = (x, edge_attr) if i == 0 else (torch.relu(x), torch.relu(edge_attr))
            dx, edge_attr = interaction_network(xin, edge_index, ein, sd, f"{prefix}layers.{i}.")
            x = sqconvex(dx, x_res, alpha) if x_res is not None else dx
            if collect:
                edge_attrs.append(edge_attr)
    else:
        raise ValueError(residual_type)
    return x, edge_attr, edge_attrs


# ------------------------------------------------------------------ Edge classifier
def ec_forward(x, edge_index, edge_attr, sd, prefix="", *, alpha=0.5, residual_type="skip1",
               use_intermediate_edge_embeddings=True, use_node_embedding=True, residual_kwargs=None):
    """``ECForGraphTCN.forward`` models/edge_classifier.py:89-121."""
    rk = dict(residual_kwargs or {})
    h = torch.relu(mlp(x, sd, prefix + "ec_node_encoder."))
    e = torch.relu(mlp(edge_attr, sd, prefix + "ec_edge_encoder."))
    h, e, es = resin(h, edge_index, e, sd, prefix + "ec_resin.network.", alpha=alpha,
                     residual_type=residual_type, collect=use_intermediate_edge_embeddings,
                     connect_to=rk.get("connect_to", 1))
    w_in = torch.cat(es, dim=1) if use_intermediate_edge_embeddings else e
    if use_node_embedding:
        w_in = torch.cat([h[edge_index[0]], h[edge_index[1]], w_in], dim=1)
    eps = 0.001
    w = eps + (1 - 2 * eps) * torch.sigmoid(mlp(w_in, sd, prefix + "W."))
    return {"W": w.squeeze(), "node_embedding": h, "edge_embedding": e}


# ------------------------------------------------------------------------ GraphTCN
def graph_tcn_forward(x, edge_index, edge_attr, sd, prefix="_gtcn.", *, alpha_ec=0.5, alpha_hc=0.5,
                      ec_threshold=0.5, mask_orphan_nodes=False, use_ec_embeddings_for_hc=False,
                      feed_edge_weights=False, alpha_latent=0.0, n_embedding_coords=0):
    """``ModularGraphTCN.forward`` models/track_condensation_networks.py:236-308
    for the ``GraphTCN`` composition (:362-381): EC -> threshold -> edge_subgraph
    (-> orphan pruning + relabel) -> HC encoders -> hc_in ResIN -> beta / H heads."""
    ec = ec_forward(x, edge_index, edge_attr, sd, prefix + "ec.", alpha=alpha_ec)
    w = ec["W"].reshape(-1)
    edge_mask = w > ec_threshold
    ei = edge_index[:, edge_mask]
    ea = edge_attr[edge_mask]
    ew = w[edge_mask].reshape(-1, 1)
    ee = ec["edge_embedding"][edge_mask]
    ne = ec["node_embedding"]
    xs = x
    n = x.size(0)
    if mask_orphan_nodes:
        connected = ei.flatten().unique()
        hit_mask = torch.zeros(n, dtype=torch.bool)
        hit_mask[connected] = True
        relabel = torch.full((n,), -1, dtype=torch.long)
        relabel[connected] = torch.arange(connected.numel())
        ei = relabel[ei]
        xs = x[connected]
        ne = ne[connected]
    else:
        hit_mask = torch.ones(n, dtype=torch.bool)
    xin, ein = [xs], [ea]
    if use_ec_embeddings_for_hc:
        xin.append(ne)
        ein.append(ee)
    if feed_edge_weights:
        ein.append(ew)
    xin = torch.cat(xin, dim=1)
    ein = torch.cat(ein, dim=1)
    h = torch.relu(res_fcnn(xin, sd, prefix + "hc_node_encoder.", alpha=0.0))
    e = torch.relu(mlp(ein, sd, prefix + "hc_edge_encoder."))
    h, _, _ = resin(h, ei, e, sd, prefix + "hc_in.network.", alpha=alpha_hc)
    beta = torch.sigmoid(mlp(h, sd, prefix + "p_beta."))
    eps = 1e-6
    beta = eps + (1 - 2 * eps) * beta
    H = mlp(h, sd, prefix + "p_cluster.")
    if alpha_latent:
        nec = n_embedding_coords
        res = torch.nn.functional.pad(xs[:, :nec], (0, H.shape[1] - nec))
        H = math.sqrt(alpha_latent) * res + math.sqrt(1 - alpha_latent) * H
    H = H * sd[prefix + "_latent_normalization"]
    return {"W": w, "H": H, "B": beta.squeeze(), "ec_hit_mask": hit_mask, "ec_edge_mask": edge_mask}


# ---------------------------------------------------------------- graph plan (ints)
def plan(edge_index: Tensor, n_nodes: int):
    """Integer oracle for the destination-sorted edge plan: stable argsort of
    edge_index[1], CSR row pointers, sorted src / dst.  Bit-exact contract."""
    dst = edge_index[1]
    perm = torch.sort(dst, stable=True).indices
    counts = torch.bincount(dst, minlength=n_nodes)
    rowptr = torch.zeros(n_nodes + 1, dtype=torch.long)
    rowptr[1:] = torch.cumsum(counts, 0)
    return perm, rowptr, edge_index[0][perm], dst[perm]
