"""B200-native Interaction-Network hot path of gnn-tracking/gnn_tracking behind the
reference's own module interface.

``install()`` swaps the drop-in classes into an importable ``gnn_tracking`` package so that the
reference's YAML configs (``class_path: gnn_tracking.models...``) and checkpoints run unchanged.
"""
from __future__ import annotations

__version__ = "0.1.0"

# reference module -> {attribute name: (our module, our attribute)}; resin.py:14 and
# track_condensation_networks.py:17 bind the IN class under their own names at import time
_PATCHES = {
    "gnn_tracking.models.interaction_network": {"InteractionNetwork": ("models.interaction_network", "InteractionNetwork")},
    "gnn_tracking.models.resin": {"InteractionNetwork": ("models.interaction_network", "InteractionNetwork"),
                                  "ResIN": ("models.resin", "ResIN")},
    "gnn_tracking.models.mlp": {"MLP": ("models.mlp", "MLP"), "ResFCNN": ("models.mlp", "ResFCNN")},
    "gnn_tracking.models.edge_classifier": {"ECForGraphTCN": ("models.edge_classifier", "ECForGraphTCN"),
                                            "PerfectEdgeClassification": ("models.edge_classifier", "PerfectEdgeClassification")},
    "gnn_tracking.models.track_condensation_networks": {
        "IN": ("models.interaction_network", "InteractionNetwork"),
        "ModularGraphTCN": ("models.track_condensation_networks", "ModularGraphTCN"),
        "GraphTCN": ("models.track_condensation_networks", "GraphTCN"),
        "PreTrainedECGraphTCN": ("models.track_condensation_networks", "PreTrainedECGraphTCN"),
        "PerfectECGraphTCN": ("models.track_condensation_networks", "PerfectECGraphTCN")},
    "gnn_tracking.metrics.losses.ec": {"EdgeWeightBCELoss": ("metrics.losses.ec", "EdgeWeightBCELoss"),
                                       "EdgeWeightFocalLoss": ("metrics.losses.ec", "EdgeWeightFocalLoss"),
                                       "HaughtyFocalLoss": ("metrics.losses.ec", "HaughtyFocalLoss")},
    "gnn_tracking.metrics.losses.oc": {"CondensationLossTiger": ("metrics.losses.oc", "CondensationLossTiger"),
                                       "CondensationLossRG": ("metrics.losses.oc", "CondensationLossRG")},
    "gnn_tracking.postprocessing.fastrescanner": {"DBSCANFastRescan": ("postprocessing.dbscan", "DBSCANFastRescan")},
    "gnn_tracking.postprocessing.dbscanscanner": {"DBSCANFastRescan": ("postprocessing.dbscan", "DBSCANFastRescan")},
    "gnn_tracking.metrics.losses.metric_learning": {
        "GraphConstructionHingeEmbeddingLoss": ("metrics.losses.metric_learning", "GraphConstructionHingeEmbeddingLoss")},
}


def install(strict: bool = False) -> list[str]:
    """Replace the hot-path classes of an installed ``gnn_tracking`` by the B200 ones.  Returns the
    patched ``module.attribute`` names; modules that cannot be imported are skipped unless
    ``strict``.  The forward is CUDA sm_100a only: there is no CPU fallback behind these classes."""
    import importlib

    done = []
    for ref_mod, attrs in _PATCHES.items():
        try:
            mod = importlib.import_module(ref_mod)
        except Exception:  # noqa: BLE001 - the reference (or one of its dependencies) is absent
            if strict:
                raise
            continue
        for name, (our_mod, our_name) in attrs.items():
            ours = getattr(importlib.import_module(f"{__name__}.{our_mod}"), our_name, None)
            if ours is None:
                if strict:
                    raise AttributeError(f"{__name__}.{our_mod}.{our_name}")
                continue
            setattr(mod, name, ours)
            done.append(f"{ref_mod}.{name}")
    return done
