"""Make the UNMODIFIED reference (``/root/reference/src/aspire``) importable.

Test infrastructure only (see ``oracle/__init__.py``).  Used by
``oracle/make_golden.py`` and by the ``not gpu`` tests that pin the numpy
restatement against the reference itself; skipped automatically when
``/root/reference`` is absent (the GPU box).

The reference needs ``array_api_compat``, ``array_api_extra``, ``h5py``,
``matplotlib`` and ``orng`` at import time; none is installed here.  scipy
vendors the first two, the other three are only touched by I/O, plotting and
RNG-wrapper code that is off the hot path, so they are stubbed.
"""

from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_SRC = "/root/reference/src"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "aspire"))


def _alias_vendored(prefix_from: str, prefix_to: str) -> None:
    for name in list(sys.modules):
        if name == prefix_from or name.startswith(prefix_from + "."):
            sys.modules.setdefault(prefix_to + name[len(prefix_from):], sys.modules[name])


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__dict__["__graft_stub__"] = True
        sys.modules[name] = mod
    for k, v in attrs.items():
        if not hasattr(mod, k):
            setattr(mod, k, v)
    return mod


def install() -> None:
    """Install aliases and stubs; idempotent."""
    os.environ.setdefault("SCIPY_ARRAY_API", "1")
    try:
        import array_api_compat  # noqa: F401
    except ImportError:
        importlib.import_module("scipy._external.array_api_compat")
        importlib.import_module("scipy._external.array_api_compat.numpy")
        importlib.import_module("scipy._external.array_api_compat.torch")
        importlib.import_module("scipy._external.array_api_compat.common._typing")
        _alias_vendored("scipy._external.array_api_compat", "array_api_compat")
    try:
        import array_api_extra  # noqa: F401
    except ImportError:
        importlib.import_module("scipy._external.array_api_extra")
        _alias_vendored("scipy._external.array_api_extra", "array_api_extra")

    try:
        import h5py  # noqa: F401
    except ImportError:
        class _File:  # pragma: no cover - never used on the hot path
            def __init__(self, *a, **k):
                raise RuntimeError("h5py is stubbed in the oracle shim")

        _stub("h5py", File=_File, Group=object, Dataset=object,
              string_dtype=lambda *a, **k: object)
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        _stub("matplotlib")
        _stub("matplotlib.figure", Figure=type("Figure", (), {}))
        _stub("matplotlib.pyplot")
    try:
        import orng  # noqa: F401
    except ImportError:
        class ArrayRNG:
            """Constructible placeholder (``samplers/mcmc.py:47`` builds one
            even when ``rng=`` is passed)."""

            def __init__(self, backend=None, **kwargs):
                self.backend = backend

        _stub("orng", ArrayRNG=ArrayRNG)
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)


def import_reference():
    """Return the reference ``aspire`` package (unmodified)."""
    if not reference_available():
        raise RuntimeError("/root/reference is not available on this machine")
    install()
    return importlib.import_module("aspire")
