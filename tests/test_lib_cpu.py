"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/overiva_b200.h
declares, its host-side helpers and argument validation work without a GPU, and the Python entry points
fail loudly (no CPU fallback) when CUDA is absent."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from overiva_b200 import _lib as L

HEADER = os.path.join(ROOT, "include", "overiva_b200.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.isfile(L.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    return L.load()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(oiva_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
        assert n in L.SIGNATURES, "no ctypes signature for " + n
    assert sorted(L.SIGNATURES) == names  # nothing bound that the header does not declare


def test_header_cites_the_reference_interfaces():
    src = open(HEADER).read()
    for cite in ("overiva.py:179", ":181-182", "overiva.py:87", "overiva.py:152-173", "overiva.py:192-199",
                 "auxiva_pca.py:79-81", "216-241"):
        assert cite in src, cite


def test_layout_helpers(lib):
    # bins are grouped by 32 (lane <-> bin); the last group of a mixture is zero padded
    assert lib.oiva_bin_groups(2049) == 65 and lib.oiva_bin_groups(32) == 1 and lib.oiva_bin_groups(33) == 2
    assert lib.oiva_grouped_bytes(1, 116, 2049, 4, L.C128) == 65 * 32 * 116 * 4 * 16
    assert lib.oiva_grouped_bytes(3, 467, 10, 8, L.C64) == 3 * 32 * 467 * 8 * 8
    assert lib.oiva_frame_pitch(116) == 128 and lib.oiva_frame_pitch(128) == 128 and lib.oiva_frame_pitch(467) == 480
    assert lib.oiva_grouped_cov_bytes(2, 2049, 6, 2) == 2 * 65 * 2 * 21 * 32 * 16


def test_argument_validation_without_gpu(lib):
    h = C.c_void_p()
    bad = L.PlanDesc(1, 100, 10, 0, 17, 2, L.MODEL_LAPLACE, L.C128, 0)  # 17 channels
    assert lib.oiva_plan_create(C.byref(h), C.byref(bad)) == -1
    assert b"n_chan" in lib.oiva_last_error()
    bad = L.PlanDesc(1, 100, 10, 0, 4, 5, L.MODEL_LAPLACE, L.C128, 0)  # more sources than channels
    assert lib.oiva_plan_create(C.byref(h), C.byref(bad)) == -1
    ok = L.PlanDesc(2, 116, 2049, 0, 4, 2, L.MODEL_LAPLACE, L.C128, 0)
    assert lib.oiva_plan_create(C.byref(h), C.byref(ok)) == 0
    nbytes = lib.oiva_plan_workspace_bytes(h)
    assert nbytes > 2 * 116 * 2049 * 4 * 16
    assert lib.oiva_plan_iterate(h, 1, None) == -3  # no workspace bound: state error, not a crash
    lib.oiva_plan_destroy(h)
    assert lib.oiva_relayout(None, None, 1, 1, 1, 1, 0, None) == -1
    # the one-pass relayout + covariance: up to 8 channels; complex64 needs 16-byte aligned rows (even F * M)
    assert lib.oiva_relayout_cov_supported(2049, 6, L.C128) == 1 and lib.oiva_relayout_cov_supported(2049, 9, L.C128) == 0
    assert lib.oiva_relayout_cov_supported(2049, 6, L.C64) == 1 and lib.oiva_relayout_cov_supported(2049, 5, L.C64) == 0
    assert lib.oiva_relayout_cov(None, None, None, None, 0, 1, 1, 1, 1, 0, None) == -1
    # scratch for the deterministic frame-split sums: none for big batches, a few slots of |Vg| for single mixtures
    assert lib.oiva_weighted_cov_scratch_bytes(512, 116, 2049, 6, 2) == 0
    assert lib.oiva_weighted_cov_scratch_bytes(1, 116, 2049, 6, 2) == 64 * lib.oiva_grouped_cov_bytes(1, 2049, 6, 2)


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import overiva_b200 as ob

    X = np.zeros((10, 5, 3), dtype=np.complex128)
    for fn, kw in ((ob.overiva, {}), (ob.auxiva, {}), (ob.ogive, {}), (ob.auxiva_pca, dict(proj_back=True))):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            fn(X, **kw)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "overiva_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
    for f in ("overiva.py", "auxiva_pca.py", "ive.py"):
        txt = open(os.path.join(ROOT, f)).read()
        assert "oracle" not in txt


def test_drop_in_modules_expose_reference_signatures():
    import inspect

    import auxiva_pca as m_pca
    import ive as m_ive
    import overiva as m_ov

    sig = inspect.signature(m_ov.overiva)
    assert list(sig.parameters) == ["X", "n_src", "n_iter", "proj_back", "W0", "model", "init_eig", "return_filters",
                                    "callback"]  # overiva.py:28-38
    assert sig.parameters["n_iter"].default == 20 and sig.parameters["model"].default == "laplace"
    sig = inspect.signature(m_ive.ogive)
    assert list(sig.parameters) == ["X", "n_iter", "step_size", "tol", "update", "proj_back", "W0", "model", "init_eig",
                                    "return_filters", "callback"]  # ive.py:33-45
    assert sig.parameters["n_iter"].default == 4000 and sig.parameters["tol"].default == 1e-3
    sig = inspect.signature(m_pca.auxiva_pca)
    assert list(sig.parameters) == ["X", "n_src", "kwargs"]  # auxiva_pca.py:30
    # the drivers' baseline: pra.bss.ilrma(X, n_iter=..., n_components=2, proj_back=True, callback=...)
    # (overiva_oneshot.py:331-339); the leading parameters and defaults are pyroomacoustics'
    from overiva_b200 import ilrma

    sig = inspect.signature(ilrma)
    assert list(sig.parameters)[:8] == ["X", "n_src", "n_iter", "proj_back", "W0", "n_components", "return_filters",
                                        "callback"]
    assert sig.parameters["n_iter"].default == 20 and sig.parameters["n_components"].default == 2
    assert sig.parameters["proj_back"].default is False


def test_host_pipeline_chunk_schedule():
    """The host pipelines start and end with small chunks (only the first copy in and the last loop + copy out are not
    overlapped by anything); every schedule covers the batch exactly, in order, with chunks no larger than asked."""
    from overiva_b200.core import _chunk_schedule

    for B in (1, 5, 31, 32, 33, 100, 191, 192, 193, 512, 4096):
        for chunk in (1, 4, 8, 32, 64):
            s = _chunk_schedule(B, chunk)
            assert sum(s) == B and all(0 < n <= chunk for n in s), (B, chunk, s)
    assert _chunk_schedule(512, 32)[:5] == [4, 4, 8, 16, 32] and _chunk_schedule(512, 32)[-4:] == [16, 8, 4, 4]
    assert _chunk_schedule(100, 32) == [32, 32, 32, 4]  # short batches: plain chunks


def test_status_words_map_to_the_reference_error_behaviour():
    """Host logic of the per-mixture status words: singular -> LinAlgError (np.linalg.solve in overiva.py:98,182),
    non-finite alone -> a RuntimeWarning and NaN results like the reference, a stalled single-launch loop or a corrupt
    word -> RuntimeError."""
    import warnings

    from overiva_b200.core import raise_for_status

    raise_for_status(np.zeros(4, dtype=np.int32))
    with pytest.raises(np.linalg.LinAlgError, match="Singular matrix"):
        raise_for_status(np.array([L.STATUS_SINGULAR], dtype=np.int32))
    with pytest.raises(np.linalg.LinAlgError, match="mixture 2 of 3"):
        raise_for_status(np.array([0, L.STATUS_NONFINITE, L.STATUS_SINGULAR | L.STATUS_NONFINITE], dtype=np.int32))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        raise_for_status(np.array([0, L.STATUS_NONFINITE], dtype=np.int32))
    assert len(w) == 1 and issubclass(w[0].category, RuntimeWarning)
    with pytest.raises(RuntimeError, match="stalled"):
        raise_for_status(np.array([0, L.STATUS_STALLED | L.STATUS_SINGULAR], dtype=np.int32))
    with pytest.raises(RuntimeError, match="corrupt"):
        raise_for_status(np.array([8], dtype=np.int32))
    src = open(HEADER).read()
    for name, val in (("OIVA_STATUS_SINGULAR", L.STATUS_SINGULAR), ("OIVA_STATUS_NONFINITE", L.STATUS_NONFINITE),
                      ("OIVA_STATUS_STALLED", L.STATUS_STALLED)):
        assert re.search(r"#define %s %d\b" % (name, val), src), name
