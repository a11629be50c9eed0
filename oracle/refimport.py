"""Import shim for the UNMODIFIED reference (auniquesun/PPT) hot-path modules.

TEST INFRASTRUCTURE ONLY.  Works only where /root/reference exists (the build
container).  Nothing on the GPU box may import this: the `-m gpu` tests, smoke()
and bench.py use the committed fixtures under tests/golden/ instead.

The hot-path modules import a few plotting / logging packages at module scope
(models/pointbert/misc.py:2-3, models/pointbert/checkpoint.py:7,
models/pointbert/point_encoder.py:4) that are absent here; they are never
touched by the functions we call, so inert stand-ins are registered in
sys.modules before the import.
"""
import contextlib
import os
import sys
import types
from unittest import mock

REFERENCE_ROOT = os.environ.get("PPT_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models", "pointbert"))


def _install_stubs():
    import torch.nn as nn

    def _mod(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    try:
        import matplotlib  # noqa: F401
    except Exception:
        mpl = _mod("matplotlib")
        mpl.pyplot = _mod("matplotlib.pyplot")
    try:
        import mpl_toolkits.mplot3d  # noqa: F401
    except Exception:
        tk = _mod("mpl_toolkits")
        tk.mplot3d = _mod("mpl_toolkits.mplot3d", Axes3D=object)
    try:
        import termcolor  # noqa: F401
    except Exception:
        _mod("termcolor", colored=lambda s, *a, **k: s)
    try:
        from timm.models.layers import DropPath  # noqa: F401
    except Exception:
        class DropPath(nn.Module):
            def __init__(self, p=0.0):
                super().__init__()

            def forward(self, x):
                return x

        timm = _mod("timm")
        timm.models = _mod("timm.models")
        timm.models.layers = _mod("timm.models.layers", DropPath=DropPath,
                                  trunc_normal_=nn.init.trunc_normal_)


def load():
    """Returns a namespace with the reference modules on the hot path."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # the reference tree is read-only
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    ns = types.SimpleNamespace()
    ns.misc = importlib.import_module("models.pointbert.misc")
    ns.dvae = importlib.import_module("models.pointbert.dvae")
    ns.pb_pn2 = importlib.import_module("models.pointbert.pointnet2_utils")
    ns.pn2 = importlib.import_module("models.pointnet2.pointnet2_utils")
    return ns


@contextlib.contextmanager
def fixed_fps_start(start=0):
    """Pins the random FPS start index (models/pointbert/misc.py:59,
    models/pointnet2/pointnet2_utils.py:75) without editing the reference."""
    import torch

    def _randint(low, high, size, **kw):
        kw.pop("generator", None)
        return torch.full(size, start, dtype=kw.get("dtype", torch.long),
                          device=kw.get("device"))

    with mock.patch("torch.randint", _randint):
        yield
