"""ctypes binding of ``liboveriva_b200.so`` (the C ABI declared in ``include/overiva_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C overiva_b200/csrc``.  There is no
fallback: if the shared object is missing, loading raises and every entry point of the package fails.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# (OVERIVA_B200_LIB: an alternative build of the same library, used by tuning scripts to compare kernel variants)
LIB_PATH = os.environ.get("OVERIVA_B200_LIB") or os.path.join(_HERE, "lib", "liboveriva_b200.so")

OK = 0
ERR_INVALID, ERR_NOMEM, ERR_STATE, ERR_CUDA, ERR_UNSUPPORTED = -1, -2, -3, -4, -5
C128, C64 = 0, 1
MODEL_LAPLACE, MODEL_GAUSS, MODEL_NONE, MODEL_OGIVE_LAPLACE, MODEL_OGIVE_GAUSS = 0, 1, 2, 3, 4
INIT_EYE, INIT_EIG, INIT_W0 = 0, 1, 2
STATUS_SINGULAR, STATUS_NONFINITE, STATUS_STALLED = 1, 2, 4


class PlanDesc(C.Structure):
    _fields_ = [
        ("n_batch", C.c_int),
        ("n_frames", C.c_int),
        ("n_freq", C.c_int),
        ("n_freq_total", C.c_int),
        ("n_chan", C.c_int),
        ("n_src", C.c_int),
        ("model", C.c_int),
        ("dtype", C.c_int),
        ("flags", C.c_int),
    ]


_i, _p, _d, _sz, _ll = C.c_int, C.c_void_p, C.c_double, C.c_size_t, C.c_longlong

# name -> (restype, argtypes); every symbol declared in include/overiva_b200.h
SIGNATURES = {
    "oiva_version": (_i, []),
    "oiva_last_error": (C.c_char_p, []),
    "oiva_bin_groups": (_i, [_i]),
    "oiva_frame_pitch": (_i, [_i]),
    "oiva_grouped_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "oiva_grouped_cov_bytes": (_sz, [_i, _i, _i, _i]),
    "oiva_relayout": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "oiva_relayout_cov_supported": (_i, [_i, _i, _i]),
    "oiva_relayout_cov": (_i, [_p, _p, _p, _p, _sz, _i, _i, _i, _i, _i, _p]),
    "oiva_weighted_cov": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "oiva_weighted_cov_scratch_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "oiva_weighted_cov_ws": (_i, [_p, _p, _p, _p, _sz, _i, _i, _i, _i, _i, _i, _p]),
    "oiva_unpack_cov": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "oiva_demix_power": (_i, [_p, _p, _i, _i, _p, _i, _i, _i, _i, _i, _i, _p]),
    "oiva_group_rows": (_i, [_p, _p, _i, _i, _i, _p]),
    "oiva_ungroup_rows": (_i, [_p, _p, _i, _i, _i, _p]),
    "oiva_sum_partials": (_i, [_p, _i, _p, _i, _i, _i, _p]),
    "oiva_source_model": (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _i, _p]),
    "oiva_source_model_ws": (_i, [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "oiva_ip_update": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "oiva_cov_ip_update": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "oiva_init_demix": (_i, [_p, _p, _p, _p, _i, _p, _i, _i, _i, _i, _p]),
    "oiva_init_demix_grouped": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "oiva_eigh": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "oiva_projback_filters": (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _p]),
    "oiva_demix_output": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "oiva_demix_output_grouped": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "oiva_project_rows": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "oiva_compose_filters": (_i, [_p, _p, _p, _i, _i, _i, _i, _p]),
    "oiva_ogive_update": (_i, [_p, _p, _p, _p, _p, _p, _p, _d, _p, _i, _i, _p]),
    "oiva_ogive_update_gated": (_i, [_p, _p, _p, _p, _p, _p, _p, _d, _p, _i, _d, _i, _i, _p]),
    "oiva_ogive_iterate": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _sz, _p, _p, _p, _p, _d, _p, _i, _i, _d, _i, _i, _i, _i, _i, _p]),
    "oiva_ogive_setup": (_i, [_p, _p, _p, _p, _i, _i, _p]),
    "oiva_ogive_a_from_w": (_i, [_p, _p, _p, _i, _i, _p]),
    "oiva_ogive_switching": (_i, [_p, _p, _p, _p, _i, _i, _p]),
    "oiva_stft_twiddle_bytes": (_sz, [_i]),
    "oiva_stft_twiddles": (_i, [_p, _i, _p]),
    "oiva_stft_num_frames": (_i, [_ll, _i, _i, _ll, _ll]),
    "oiva_stft_analysis": (_i, [_p, _i, _ll, _ll, _ll, _ll, _ll, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "oiva_stft_scratch_bytes": (_sz, [_i, _i, _i, _i]),
    "oiva_stft_synthesis": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "oiva_gram_scratch_bytes": (_sz, [_i, _ll]),
    "oiva_gram": (_i, [_p, _ll, _ll, _i, _p, _ll, _ll, _i, _ll, _p, _p, _p]),
    "oiva_xcorr_scratch_bytes": (_sz, [_i, _i, _ll, _i]),
    "oiva_xcorr": (_i, [_p, _ll, _ll, _i, _p, _ll, _ll, _i, _ll, _i, _p, _p, _p]),
    "oiva_loop_resident_sync_bytes": (_sz, [_i, _i]),
    "oiva_loop_resident": (_i, [_p, _p, _p, _p, _p, _p, _sz, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "oiva_fp64_peak": (_i, [_i, _i, _i, C.POINTER(C.c_double), _p]),
    "oiva_demix_power_full": (_i, [_p, _p, _i, _i, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "oiva_weighted_cov_binwise": (_i, [_p, _p, _p, _p, _sz, _i, _i, _i, _i, _i, _p]),
    "oiva_ilrma_vpart_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "oiva_ilrma_set_model": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "oiva_ilrma_get_model": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "oiva_ilrma_nmf": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _d, _p]),
    "oiva_ilrma_rescale": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "oiva_ilrma_iterate": (_i, [_p, _p, _p, _p, _p, _p, _p, _sz, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _d, _i, _p]),
    "oiva_ilrma_fill_scale": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "oiva_demix_output_scaled": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "oiva_plan_array": (_p, [_p, _i]),
    "oiva_plan_scratch_bytes": (_sz, [_p]),
    "oiva_plan_create": (_i, [C.POINTER(_p), C.POINTER(PlanDesc)]),
    "oiva_plan_destroy": (None, [_p]),
    "oiva_plan_workspace_bytes": (_sz, [_p]),
    "oiva_plan_bind": (_i, [_p, _p, _sz]),
    "oiva_plan_load": (_i, [_p, _p, _p]),
    "oiva_plan_adopt_samples": (_i, [_p, _p]),
    "oiva_plan_init": (_i, [_p, _i, _p, _p]),
    "oiva_plan_iterate": (_i, [_p, _i, _p]),
    "oiva_plan_power": (_i, [_p, _p]),
    "oiva_plan_update": (_i, [_p, _p]),
    "oiva_plan_r2": (_p, [_p]),
    "oiva_plan_r2_elems": (_sz, [_p]),
    "oiva_plan_output": (_i, [_p, _i, _p, _p]),
    "oiva_plan_filters": (_i, [_p, _p, _p]),
    "oiva_plan_run": (_i, [_p, _p, _i, _p, _i, _i, _p, _p, _p]),
    "oiva_plan_what": (_p, [_p]),
    "oiva_plan_cov": (_p, [_p]),
    "oiva_plan_samples": (_p, [_p]),
    "oiva_plan_status_ptr": (_p, [_p]),
    "oiva_plan_status": (_i, [_p, _p]),
    "oiva_plan_status_vector": (_i, [_p, C.POINTER(C.c_int), _p]),
    "oiva_plan_reset_status": (_i, [_p, _p]),
    "oiva_plan_launch_count": (C.c_longlong, [_p]),
    "oiva_plan_enable_timing": (_i, [_p, _i]),
    "oiva_plan_read_timing": (_i, [_p, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "oiva_overiva_host": (_i, [_p, _p, _p, _p, C.POINTER(PlanDesc), _i, _i, _i, C.POINTER(C.c_int)]),
}

_lib = None


class OverivaLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and attach the signatures.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise OverivaLibraryError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C overiva_b200/csrc -j8` (there is no CPU / PyTorch fallback)" % LIB_PATH
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().oiva_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = ""):
    """Turn a non-zero return code of a library call into an exception."""
    if rc != OK:
        raise OverivaLibraryError("%s failed (code %d): %s" % (what or "overiva_b200 call", rc, last_error()))
