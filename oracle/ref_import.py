"""Import the real reference modules from /root/reference — BUILD-CONTAINER ONLY.

The reference tree is not importable as shipped (network/_deeplab.py:8-13 imports
matplotlib, src.utils, src.datasets; utils/__init__.py:2 pulls in visdom). This module
installs empty stand-ins for exactly those names (none is touched by model code) and puts
the reference on sys.path. /root/reference does not exist on the GPU box, so nothing under
tests/ -m gpu, smoke() or bench.py may import this file; only oracle/gen_golden.py and the
`-m "not gpu"` pinning tests (which skip when the tree is absent) do.
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("ISWM_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "network"))


def install_stubs() -> None:
    def mod(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    try:
        import matplotlib  # noqa: F401
    except Exception:
        mp = mod("matplotlib")
        mp.pyplot = mod("matplotlib.pyplot")
    src = mod("src")
    src.utils = mod("src.utils", ext_transforms=types.ModuleType("ext_transforms"))
    src.datasets = mod("src.datasets", FeatureVisDataset=object)
    mod("visdom", Visdom=object)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


def reference_modules():
    """Returns (network.modeling, metrics) of the reference."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    install_stubs()
    # the reference's top-level package names (network, metrics, utils) are generic: make sure
    # we import THEM and not something else already on the path
    for name in ("network", "metrics"):
        m = sys.modules.get(name)
        if m is not None and not getattr(m, "__file__", "").startswith(REF_ROOT):
            del sys.modules[name]
    import network.modeling as modeling  # type: ignore
    import metrics as ref_metrics  # type: ignore
    return modeling, ref_metrics
