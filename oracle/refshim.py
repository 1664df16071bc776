"""TEST INFRASTRUCTURE ONLY — never imported by the product package `fastmot_b200`.

Imports the *unmodified* reference (GeekAlexis/FastMOT) from /root/reference inside THIS
container so that (a) the numpy restatements in `oracle/` can be pinned against it and
(b) golden fixtures under tests/golden/ can be generated (oracle/make_goldens.py).

The reference tree does not exist on the GPU box: nothing that runs under `-m gpu`,
`__graft_entry__.smoke()` or `bench.py` may call `load_reference()`.

Recipe follows SURVEY.md Appendix B: cupy / cupyx / tensorrt are absent here and are
replaced by inert stub modules; `cupyx.empty_pinned` maps to `np.empty`.
"""
import os
import sys
import types
import tempfile

import numpy as np

REF_ENV = "FASTMOT_REF"
DEFAULT_REF = "/root/reference"


class _Any(types.ModuleType):
    """Module stub that tolerates attribute chains, calls and int() conversion."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        child = _Any(f"{self.__name__}.{name}")
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        return _Any(self.__name__ + "()")

    def __int__(self):
        return 0

    def __index__(self):
        return 0

    def __iter__(self):
        return iter(())


def reference_available():
    root = os.environ.get(REF_ENV, DEFAULT_REF)
    return os.path.isdir(os.path.join(root, "fastmot"))


def load_reference():
    """Returns the imported reference package `fastmot` (cached in sys.modules)."""
    if "fastmot" in sys.modules and hasattr(sys.modules["fastmot"], "MultiTracker"):
        return sys.modules["fastmot"]
    root = os.environ.get(REF_ENV, DEFAULT_REF)
    if not os.path.isdir(os.path.join(root, "fastmot")):
        raise RuntimeError(f"reference tree not found at {root}")
    # keep numba / python caches out of the (root-writable) reference tree
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(tempfile.gettempdir(), "fastmot_ref_nbcache"))
    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    sys.dont_write_bytecode = True
    for name in ("cupy", "cupyx", "cupyx.scipy", "cupyx.scipy.ndimage", "tensorrt"):
        if name not in sys.modules:
            sys.modules[name] = _Any(name)
    cupyx = sys.modules["cupyx"]
    cupyx.empty_pinned = lambda shape, dtype=float: np.empty(shape, dtype)
    cupyx.empty_like_pinned = np.empty_like
    cupyx.scipy = sys.modules["cupyx.scipy"]
    cupyx.scipy.ndimage = sys.modules["cupyx.scipy.ndimage"]
    if root not in sys.path:
        sys.path.insert(0, root)
    import fastmot  # noqa: E402
    return fastmot


def reference_config():
    """cfg/mot.json of the reference decoded the way app.py does (app.py:57-58)."""
    import json
    from types import SimpleNamespace
    fastmot = load_reference()
    root = os.environ.get(REF_ENV, DEFAULT_REF)
    with open(os.path.join(root, "cfg", "mot.json")) as f:
        return json.load(f, cls=fastmot.utils.ConfigDecoder,
                         object_hook=lambda d: SimpleNamespace(**d))
