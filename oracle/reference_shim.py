"""Import the UNMODIFIED reference (``/root/reference/{overiva,auxiva_pca,ive}.py``) in-process.

TEST INFRASTRUCTURE ONLY (see ``oracle/overiva_oracle.py``).  The reference cannot be imported
as-is in this image:

1. it imports ``pyroomacoustics.bss.projection_back`` (overiva.py:25, ive.py:30,
   auxiva_pca.py:27) and pyroomacoustics is not installed -> a stub module exposing the
   oracle's ``projection_back`` formula is placed in ``sys.modules``;
2. ``overiva.py:182`` calls ``np.linalg.solve(A (F,M,M), b (F,M))`` with numpy-1.x "stack of
   vectors" semantics, which numpy >= 2 rejects -> ``numpy.linalg.solve`` is wrapped for the
   duration of each reference call.

The reference tree stays read-only and no source is copied from it.  ``/root/reference`` does
not exist on the GPU box; ``oracle/build_ref.py`` byte-compiles the three modules into the
git-ignored ``oracle/_ref/*.refbc`` (a build output, like a compiled C reference), which travels
with the snapshot and is imported here when the source tree is absent.  ``available()`` says
whether the real implementation can be used (``source()`` tells which form).
"""
from __future__ import annotations

import contextlib
import importlib.machinery
import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_DIR = os.environ.get("OVERIVA_REFERENCE_DIR", "/root/reference")
COMPILED_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

_modules = {}


def source():
    """'tree' (the .py files under REFERENCE_DIR), 'compiled' (oracle/_ref/*.refbc built from them) or None."""
    if os.path.isfile(os.path.join(REFERENCE_DIR, "overiva.py")):
        return "tree"
    if all(os.path.isfile(os.path.join(COMPILED_DIR, m + ".refbc")) for m in ("overiva", "auxiva_pca", "ive")):
        return "compiled"
    return None


def available() -> bool:
    return source() is not None


class _BytecodeLoader(importlib.machinery.SourcelessFileLoader):
    """SourcelessFileLoader for a .pyc image stored under another extension."""

    def get_code(self, fullname):
        import marshal

        data = self.get_data(self.get_filename(fullname))
        if data[:4] != importlib.util.MAGIC_NUMBER:
            raise ImportError("%s was compiled by another Python version; rebuild oracle/_ref" % self.path)
        return marshal.loads(data[16:])


def _install_pra_stub():
    if "pyroomacoustics" in sys.modules and not getattr(
        sys.modules["pyroomacoustics"], "_oiva_stub", False
    ):
        return  # a real pyroomacoustics is installed: use it
    from oracle.overiva_oracle import projection_back

    def _projection_back(Y, ref, clip_up=None, clip_down=None):
        return projection_back(Y, ref)

    pra = types.ModuleType("pyroomacoustics")
    pra._oiva_stub = True
    bss = types.ModuleType("pyroomacoustics.bss")
    bss.projection_back = _projection_back
    pra.bss = bss
    sys.modules["pyroomacoustics"] = pra
    sys.modules["pyroomacoustics.bss"] = bss


@contextlib.contextmanager
def numpy1_solve():
    """numpy-1.x rule: ``b.ndim == a.ndim - 1`` means a stack of vectors."""
    orig = np.linalg.solve

    def solve(a, b):
        a_ = np.asarray(a)
        b_ = np.asarray(b)
        if a_.ndim > 2 and b_.ndim == a_.ndim - 1:
            return orig(a_, b_[..., None])[..., 0]
        return orig(a, b)

    np.linalg.solve = solve
    try:
        yield
    finally:
        np.linalg.solve = orig


def _load(name):
    """Load ``/root/reference/<name>.py`` under a private module name (the repo root has its own
    drop-in ``overiva.py`` etc., so the plain import name must not be used)."""
    if name in _modules:
        return _modules[name]
    src = source()
    if src is None:
        raise RuntimeError("reference not present (neither %s nor %s)" % (REFERENCE_DIR, COMPILED_DIR))
    _install_pra_stub()
    if src == "tree":
        path = os.path.join(REFERENCE_DIR, name + ".py")
        spec = importlib.util.spec_from_file_location("_reference_" + name, path)
    else:
        path = os.path.join(COMPILED_DIR, name + ".refbc")  # a .pyc image (oracle/build_ref.py)
        loader = _BytecodeLoader("_reference_" + name, path)
        spec = importlib.util.spec_from_loader("_reference_" + name, loader, origin=path)
    mod = importlib.util.module_from_spec(spec)
    if name == "auxiva_pca":
        # auxiva_pca.py:28 does ``from overiva import overiva``: point it at the real one
        saved = sys.modules.get("overiva")
        sys.modules["overiva"] = _load("overiva")
        try:
            spec.loader.exec_module(mod)
        finally:
            if saved is not None:
                sys.modules["overiva"] = saved
            else:
                del sys.modules["overiva"]
    else:
        spec.loader.exec_module(mod)
    _modules[name] = mod
    return mod


def ref_overiva(X, **kwargs):
    with numpy1_solve():
        return _load("overiva").overiva(X, **kwargs)


def ref_auxiva_pca(X, **kwargs):
    with numpy1_solve():
        return _load("auxiva_pca").auxiva_pca(X, **kwargs)


def ref_ogive(X, **kwargs):
    with numpy1_solve():
        return _load("ive").ogive(X, **kwargs)
