"""TEST INFRASTRUCTURE ONLY -- import the *real* reference classes from
``/root/reference/src`` (authoring container only; the path does not exist on the
GPU box).  Used to pin ``oracle/*.py`` and to generate ``tests/golden``.
"""
from __future__ import annotations

import os
import sys

REFERENCE_SRC = os.environ.get("GT_REFERENCE_SRC", "/root/reference/src")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "gnn_tracking"))


def load():
    """Returns the imported ``gnn_tracking`` package of the reference."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_SRC}")
    os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")  # bypass @torch.compile in losses/oc.py
    from oracle import shims

    shims.install()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import gnn_tracking  # noqa: F401

    return gnn_tracking
