"""Import shim for the UNMODIFIED reference (test infrastructure only).

This file is part of the oracle tooling: it is used by ``oracle/gen_golden.py`` in the
build container (where ``/root/reference`` exists) to run the reference's own code
and record golden input/output vectors under ``tests/golden/``.  Nothing on the
product path, and nothing that runs on the GPU box, may import it.

Recipe: the reference imports ``sacred``, ``librosa``,
``soundfile`` and ``matplotlib`` at module-import time
(/root/reference/src/models/wavernn.py:14-17, config.py:2-3, utils.py:3-9,
quantization/cb_func.py:10-12) although none of them is *called* on the hot path.
They are absent from this image, so inert stubs are planted in ``sys.modules``
before ``/root/reference/src`` is put on ``sys.path``.  ``models.wavernn.device``
is a module global looked up at call time (wavernn.py:20,177), so setting it to
``'cpu'`` makes ``Wavernn.encoder`` run on the host.
"""
import os
import sys
import types

REFERENCE_SRC = os.environ.get("FPC_REFERENCE_SRC", "/root/reference/src")
_LOADED = None


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "models", "wavernn.py"))


class _Ingredient:
    """Inert stand-in for sacred.Experiment / sacred.Ingredient."""

    def __init__(self, *a, **k):
        pass

    def _identity(self, fn=None, **_k):
        if fn is None:
            return lambda f: f
        return fn

    config = capture = automain = main = command = named_config = _identity

    def add_config(self, *a, **k):
        pass


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package so "import x.y" works
    sys.modules[name] = m
    return m


def load_reference():
    """Returns (wavernn_module, vq_func_module, cb_func_module) of the reference."""
    if not reference_available():
        raise RuntimeError("reference sources not present at %s" % REFERENCE_SRC)
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    for name in ("models", "quantization", "config", "utils"):
        if name in sys.modules:
            raise RuntimeError(
                "module %r already imported from elsewhere; run the reference shim "
                "in its own process" % name)
    _stub("sacred", Experiment=_Ingredient, Ingredient=_Ingredient)
    _stub("soundfile")
    lib = _stub("librosa")
    _stub("librosa.display")
    lib.display = sys.modules["librosa.display"]
    mpl = _stub("matplotlib", use=lambda *a, **k: None)
    _stub("matplotlib.pyplot")
    mpl.pyplot = sys.modules["matplotlib.pyplot"]
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import models.wavernn as W  # noqa: E402
    from quantization import vq_func, cb_func  # noqa: E402
    W.device = "cpu"
    _LOADED = (W, vq_func, cb_func)
    return _LOADED
