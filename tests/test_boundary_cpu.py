"""Host-side checks that need no GPU: the C-ABI library loads and exports every symbol
include/gtb200.h declares, the modules keep the reference's constructor / hparams /
state_dict contract, and CPU tensors are refused (no fallback)."""
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    from gnn_tracking_b200 import _lib
    header = (ROOT / "include" / "gtb200.h").read_text()
    declared = set(re.findall(r"\b(gtb_[a-z0-9_]+)\s*\(", header))
    declared -= {"gtb_src_t", "gtb_mlp_desc_t"}
    assert declared, "no declarations found"
    handle = _lib.lib()
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} declared in gtb200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert handle.gtb_version() >= 100


def test_desc_struct_matches_header_layout():
    from gnn_tracking_b200 import _lib
    assert _lib.C.sizeof(_lib.Src) == 32
    assert _lib.MlpDesc.srcs.offset == 16
    assert _lib.MlpDesc.dims.offset == 16 + 32 * 16
    assert _lib.MlpDesc.packed.offset == 544
    assert _lib.MlpDesc.gate.offset == 648 and _lib.MlpDesc.hidden_ld.offset == 660
    assert _lib.MlpDesc.hidden0.offset == 664 and _lib.MlpDesc.hidden1.offset == 672 and _lib.C.sizeof(_lib.MlpDesc) == 680


def test_packed_bytes_host_logic():
    from gnn_tracking_b200 import _lib
    import ctypes as C
    L = _lib.lib()

    def nbytes(dims, impl, blocks=None):
        d = (C.c_int32 * len(dims))(*dims)
        b = (C.c_int32 * len(blocks))(*blocks) if blocks else None
        return L.gtb_mlp_packed_bytes(len(dims) - 1, d, len(blocks) if blocks else 0, b, impl)

    assert nbytes((192, 64, 64, 64), _lib.IMPL_FFMA) == 4 * (192 * 64 + 64 + 64 * 64 + 64 + 64 * 64 + 64)
    # narrow last layer padded to 8 columns, K0 to 32
    assert nbytes((14, 64, 64, 4), _lib.IMPL_FFMA) == 4 * (32 * 64 + 64 + 64 * 64 + 64 + 64 * 8 + 8)
    assert nbytes((14, 300, 64, 4), _lib.IMPL_FFMA) == 0  # unsupported width
    # tcgen05 layout: per layer hi + lo K-major tiles [npad][32] fp32 per 32-wide K tile, + 64-float bias rows
    assert nbytes((64, 64, 64, 64), _lib.IMPL_TCGEN05) == 3 * 2 * (2 * 64 * 128) + 3 * 256
    # blocks are padded to 8 columns each (5 -> 8, 4 -> 8: K0 = 16 -> one K tile); N = 5 -> 16 rows
    assert nbytes((9, 64, 64, 5), _lib.IMPL_TCGEN05, (5, 4)) == 2 * (64 * 128) + 2 * (2 * 64 * 128) + 2 * (2 * 16 * 128) + 3 * 256
    # a last Linear with <= 4 outputs behind a hidden layer is kept as plain fp32 rows (CUDA-core dot products)
    assert nbytes((64, 64, 64, 1), _lib.IMPL_TCGEN05) == 2 * 2 * (2 * 64 * 128) + 64 * 4 + 2 * 256 + 16
    assert L.gtb_mlp_tc_slots(3, (C.c_int32 * 4)(256, 64, 64, 1), 4, (C.c_int32 * 4)(64, 64, 64, 64)) == 2  # the W head: two teams
    assert nbytes((9, 64, 64, 5), _lib.IMPL_TCGEN05, (5, 5)) == 0      # blocks do not sum to K0
    assert nbytes((64, 128, 128, 64), _lib.IMPL_TCGEN05) == 0          # Linear wider than 64: FFMA path
    assert nbytes((18, 64, 64, 1), _lib.IMPL_TCGEN05, (18,)) == 0      # neither 16-byte rows nor narrow
    assert nbytes((384, 64, 64, 1), _lib.IMPL_TCGEN05, (64,) * 6) == 0  # weights + one staging slot exceed 227 KB
    assert nbytes((256, 64, 64, 1), _lib.IMPL_TCGEN05, (64,) * 4) > 0


@pytest.mark.parametrize("kind", ["in", "resin", "ec", "tcn"])
def test_state_dict_and_hparams_contract(kind, golden_models):
    """Reference checkpoints load with strict=True: same parameter names and shapes."""
    from gnn_tracking_b200.models.edge_classifier import ECForGraphTCN
    from gnn_tracking_b200.models.interaction_network import InteractionNetwork
    from gnn_tracking_b200.models.resin import ResIN
    from gnn_tracking_b200.models.track_condensation_networks import GraphTCN
    cls = {"in": InteractionNetwork, "resin": ResIN, "ec": ECForGraphTCN, "tcn": GraphTCN}[kind]
    n = 0
    for name, c in golden_models.items():
        if c["kind"] != kind:
            continue
        kw = {k: (dict(v) if isinstance(v, dict) else v) for k, v in c["kwargs"].items()}
        m = cls(**kw)
        missing, unexpected = m.load_state_dict(c["state_dict"], strict=True)
        assert not missing and not unexpected
        assert list(m.state_dict().keys()) == list(c["state_dict"].keys()), name
        for k, v in c["kwargs"].items():
            assert m.hparams[k] == v or kind == "tcn"
        n += 1
    assert n > 0


def test_hparams_attribute_dict_is_deepcopy_safe():
    import copy
    from gnn_tracking_b200.models.interaction_network import InteractionNetwork
    m = InteractionNetwork(node_indim=3, edge_indim=2)
    assert m.hparams.node_indim == 3 and m.hparams.edge_outdim == 4 and m.hparams.aggr == "add"
    m2 = copy.deepcopy(m)
    assert m2.hparams.node_hidden_dim == 40


def test_no_cpu_fallback():
    from gnn_tracking_b200.models.edge_classifier import ECForGraphTCN
    m = ECForGraphTCN(node_indim=3, edge_indim=2, L_ec=1)

    class D:
        x = torch.zeros(4, 3)
        edge_index = torch.zeros(2, 2, dtype=torch.long)
        edge_attr = torch.zeros(2, 2)

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(D())


def test_width_mismatch_is_an_assertion_error():
    from gnn_tracking_b200.models.interaction_network import InteractionNetwork
    m = InteractionNetwork(node_indim=3, edge_indim=2)
    with pytest.raises(AssertionError):
        m(torch.zeros(4, 5), torch.zeros(2, 2, dtype=torch.long), torch.zeros(2, 2))


def test_install_patches_the_reference_package():
    """`gnn_tracking_b200.install()` rebinds the hot-path classes inside the reference package
    (authoring container only: needs /root/reference through the oracle's shims)."""
    from oracle import reference_loader as rl
    if not rl.available():
        pytest.skip("reference sources not present")
    rl.load()
    import gnn_tracking_b200
    import importlib
    saved = {}
    for ref_mod, attrs in gnn_tracking_b200._PATCHES.items():
        try:
            mod = importlib.import_module(ref_mod)
        except ModuleNotFoundError as exc:  # a third-party dependency of that reference module is not installed here
            assert not exc.name.startswith("gnn_tracking"), exc
            continue
        saved[ref_mod] = {k: getattr(mod, k) for k in attrs}
    try:
        done = gnn_tracking_b200.install()
        assert {f"{m}.{k}" for m, attrs in saved.items() for k in attrs} <= set(done)
        assert "gnn_tracking.postprocessing.fastrescanner.DBSCANFastRescan" in done
        assert "gnn_tracking.models.resin.InteractionNetwork" in done
        import gnn_tracking.models.resin as ref_resin
        import gnn_tracking.models.track_condensation_networks as ref_tcn
        from gnn_tracking_b200.models.interaction_network import InteractionNetwork
        assert ref_resin.InteractionNetwork is InteractionNetwork and ref_tcn.IN is InteractionNetwork
        # a reference-side ResIN now builds B200 layers with the reference's ctor arguments
        m = ref_resin.ResIN(node_dim=4, edge_dim=3, object_hidden_dim=8, relational_hidden_dim=8, n_layers=2)
        assert all(isinstance(l, InteractionNetwork) for l in m.network.layers)
    finally:
        for ref_mod, attrs in saved.items():
            mod = importlib.import_module(ref_mod)
            for k, v in attrs.items():
                setattr(mod, k, v)


def test_loss_wrappers_host_logic():
    """DummyMultiLoss / LossClones (reference metrics/losses/__init__.py:44-131) are host logic."""
    import torch

    from gnn_tracking_b200.metrics.losses import DummyMultiLoss, LossClones

    r = DummyMultiLoss()(x=torch.arange(4.0), other=1)
    assert float(r.loss) == 6.0 and list(r.loss_dct) == ["dummy"]

    seen = []

    class Probe(torch.nn.Module):
        def forward(self, *, w, y, extra, **kw):
            seen.append(sorted(kw))
            return (w - y).sum() + extra

    out = LossClones(Probe())(w_1=torch.ones(2), y_1=torch.zeros(2), w_0=torch.zeros(2), y_0=torch.zeros(2), extra=1.0,
                              w=torch.full((2,), 9.0))
    assert list(out) == ["0", "1"] and float(out["0"]) == 1.0 and float(out["1"]) == 3.0
    assert seen[0] == ["w_1", "y_1"] and seen[1] == ["w_0", "y_0"]  # the other clone's tensors pass through renamed-free


def test_perfect_edge_classification():
    """The reference's own cases (tests/test_edge_classifier.py:18-39); host logic, runs on any device."""
    import torch

    from gnn_tracking_b200.models.edge_classifier import PerfectEdgeClassification

    class MockData:
        def __init__(self, y, pt=None):
            self.y = torch.Tensor(y)
            self.pt = torch.Tensor(pt) if pt is not None else torch.full_like(self.y, 0.5)

    y = [True, False, True, False]
    assert (PerfectEdgeClassification().forward(MockData(y))["W"] == torch.Tensor(y)).all()
    w = PerfectEdgeClassification(false_below_pt=0.5).forward(MockData(y, pt=[0, 0, 1, 1]))["W"]
    assert (w == torch.Tensor([False, False, True, False])).all()
    torch.manual_seed(0)
    assert 35 < PerfectEdgeClassification(tpr=0.5).forward(MockData(torch.full((100,), True)))["W"].sum() < 65


def test_skip2_batch_norm_state_dict_contract():
    """``add_bn=True`` keeps the reference's module tree: its state_dict (weights, running statistics,
    num_batches_tracked of every BatchNorm1d) loads with ``strict=True``."""
    from gnn_tracking_b200.models.resin import ResIN
    from tests.golden.common import load
    for case in load("resin_bn").values():
        kw = {k: (dict(v) if isinstance(v, dict) else v) for k, v in case["kwargs"].items()}
        m = ResIN(**kw)
        m.load_state_dict(case["state_dict"], strict=True)
        assert sorted(m.state_dict().keys()) == sorted(case["state_dict"].keys())
