"""overiva_b200 -- B200-native (sm_100a) implementation of the OverIVA / AuxIVA / OGIVE demixing loop.

Drop-in entry points (same signatures as onolab-tmu/overiva):

    from overiva_b200 import overiva, auxiva, auxiva_pca, ogive, ilrma

(``ilrma`` stands in for ``pyroomacoustics.bss.ilrma``, the baseline the reference's drivers call), plus ``overiva_batch`` for many independent mixtures and ``overiva_b200.distributed`` for the
multi-GPU drivers.  See DESIGN.md / INTEGRATION.md.
"""
from .core import (DemixPlan, auxiva, auxiva_pca, clear_plan_cache, ogive, overiva, overiva_batch,  # noqa: F401
                   raise_for_status)
from .ilrma import ilrma  # noqa: F401

__all__ = ["overiva", "auxiva", "auxiva_pca", "ogive", "ilrma", "overiva_batch", "DemixPlan", "clear_plan_cache",
           "raise_for_status"]
__version__ = "0.1.0"
