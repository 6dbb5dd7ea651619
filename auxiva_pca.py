"""Drop-in module: ``from auxiva_pca import auxiva_pca`` (onolab-tmu/overiva ``auxiva_pca.py:30``)."""
from overiva_b200.core import auxiva_pca  # noqa: F401
