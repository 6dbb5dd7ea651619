"""Drop-in module: ``from ive import ogive`` (onolab-tmu/overiva ``ive.py:33-45``)."""
from overiva_b200.core import ogive  # noqa: F401
