"""The C ABI used from outside Python: include/overiva_b200.h compiles as plain C (gcc -std=c99 -pedantic), a C
program links against liboveriva_b200.so (CPU: build + symbol resolution; GPU: it separates a mixture through
oiva_overiva_host and matches the oracle), and the same entry point through ctypes with host numpy buffers."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, rel_err
from overiva_b200 import _lib as L

SRC = os.path.join(ROOT, "tests", "c", "host_call.c")


def _build(tmp_path):
    exe = os.path.join(tmp_path, "host_call")
    libdir = os.path.dirname(L.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L", libdir, "-loveriva_b200", "-Wl,-rpath," + libdir]
    subprocess.run(cmd, check=True, capture_output=True)
    return exe


def test_header_is_plain_c_and_program_links(tmp_path):
    L.load()
    exe = _build(str(tmp_path))
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr  # runs (shared library resolved), no GPU work without arguments


@pytest.mark.gpu
def test_c_program_separates_a_mixture(tmp_path):
    from oracle import overiva_oracle as orc
    from overiva_b200.synth import small_test_mixture

    exe = _build(str(tmp_path))
    Xs = np.stack([small_test_mixture(40 + b, 4, 2, n_samples=1600, frame=64, hop=32) for b in range(2)])
    B, T, F, M = Xs.shape
    K = 2
    fin, fout = os.path.join(tmp_path, "x.bin"), os.path.join(tmp_path, "y.bin")
    Xs.tofile(fin)
    r = subprocess.run([exe, fin, fout] + [str(v) for v in (B, T, F, M, K, 8, L.MODEL_GAUSS, L.INIT_EIG, 1)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = np.fromfile(fout, dtype=np.complex128)
    Y = raw[: B * T * F * K].reshape(B, T, F, K)
    W = raw[B * T * F * K :].reshape(B, F, M, K)
    for b in range(B):
        Yo, Wo = orc.overiva(Xs[b], n_src=K, n_iter=8, model="gauss", init_eig=True, return_filters=True)
        assert rel_err(Y[b], Yo) <= 1e-10 and rel_err(W[b], Wo) <= 1e-10


@pytest.mark.gpu
def test_host_entry_point_through_ctypes():
    from oracle import overiva_oracle as orc
    from overiva_b200.synth import small_test_mixture

    lib = L.load()
    X = small_test_mixture(50, 3, 2, n_samples=1400, frame=64, hop=32)
    T, F, M = X.shape
    K = 2
    rng = np.random.default_rng(0)
    W0 = np.zeros((F, M, K), dtype=np.complex128)
    W0[:, :K, :] = np.eye(K)
    W0 += 0.1 * (rng.standard_normal(W0.shape) + 1j * rng.standard_normal(W0.shape))
    Xc = np.ascontiguousarray(X[None])
    Y = np.empty((1, T, F, K), dtype=np.complex128)
    W = np.empty((1, F, M, K), dtype=np.complex128)
    desc = L.PlanDesc(1, T, F, 0, M, K, L.MODEL_LAPLACE, L.C128, 0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    st = lib.oiva_overiva_host(p(Xc), p(Y), p(W), p(np.ascontiguousarray(W0[None])), C.byref(desc), 6, 1, L.INIT_W0,
                               None)
    assert st == 0, L.last_error()
    Yo, Wo = orc.overiva(X, n_src=K, n_iter=6, W0=W0, return_filters=True)
    assert rel_err(Y[0], Yo) <= 1e-10 and rel_err(W[0], Wo) <= 1e-10
    # complex64 storage, no filters requested
    X32 = Xc.astype(np.complex64)
    Y32 = np.empty((1, T, F, K), dtype=np.complex64)
    desc32 = L.PlanDesc(1, T, F, 0, M, K, L.MODEL_LAPLACE, L.C64, 0)
    assert lib.oiva_overiva_host(p(X32), p(Y32), None, None, C.byref(desc32), 6, 1, L.INIT_EYE, None) == 0
    assert rel_err(Y32[0].astype(np.complex128), orc.overiva(X, n_src=K, n_iter=6)) <= 1e-3
    # numerical failure is reported through the returned status word, argument errors through a negative code
    bad = Xc.copy()
    bad[..., 2] = bad[..., 0]
    per = (C.c_int * 1)()
    assert lib.oiva_overiva_host(p(bad), p(Y), None, None, C.byref(desc), 3, 1, L.INIT_EYE, per) & L.STATUS_SINGULAR
    assert per[0] & L.STATUS_SINGULAR
    assert lib.oiva_overiva_host(p(Xc), p(Y), None, None, C.byref(desc), 3, 1, L.INIT_W0, None) == L.ERR_INVALID  # W0 missing
