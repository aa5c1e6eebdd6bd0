"""Import the UNMODIFIED reference module (build container only).

``/root/reference/EKFGPSSLAM.py`` imports pyproj, matplotlib and tkinter at module
top (EKFGPSSLAM.py:6-15); none is installed here.  This loader registers stub
modules for them in ``sys.modules`` (pyproj.Proj -> oracle.utm_kruger.KruegerProj)
and then imports the reference as-is, so its hot-path functions can be called to
pin the restatement in oracle/fusion_oracle.py and to generate tests/golden/.

The reference does not exist on the GPU box: nothing in ``-m gpu`` tests,
``smoke()`` or ``bench.py`` calls this loader.  ``reference_available()`` lets
CPU tests skip cleanly elsewhere.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import types

REFERENCE_DIR = os.environ.get("GSF_REFERENCE_DIR", "/root/reference")
_cached = None


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "EKFGPSSLAM.py"))


def _install_stubs() -> None:
    from .utm_kruger import KruegerProj

    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    class CRSError(Exception):
        pass

    class _Dummy:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, name):
            return _Dummy()

        def __call__(self, *a, **k):
            return _Dummy()

    if "pyproj" not in sys.modules:
        mod("pyproj", Proj=KruegerProj)
        mod("pyproj.exceptions", CRSError=CRSError)
    if "matplotlib" not in sys.modules:
        mod("matplotlib")
        mod("matplotlib.pyplot", **{"__getattr__": lambda name: _Dummy()})
        mod("matplotlib.widgets", CheckButtons=_Dummy)
    if "mpl_toolkits" not in sys.modules:
        mod("mpl_toolkits")
        mod("mpl_toolkits.mplot3d", Axes3D=_Dummy)
    if "tkinter" not in sys.modules:
        fd = mod("tkinter.filedialog")
        mb = mod("tkinter.messagebox")
        mod("tkinter", Tk=_Dummy, filedialog=fd, messagebox=mb)


def load_reference():
    """Return the reference module object (cached)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_DIR}")
    _install_stubs()
    spec = importlib.util.spec_from_file_location(
        "_gsf_reference_EKFGPSSLAM", os.path.join(REFERENCE_DIR, "EKFGPSSLAM.py"))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    _cached = module
    return module


@contextlib.contextmanager
def quiet():
    """The reference prints on every call; swallow it."""
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(buf):
        yield buf
