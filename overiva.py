"""Drop-in module: ``from overiva import overiva`` resolves to the B200 implementation
(same signature as onolab-tmu/overiva ``overiva.py:28-38``)."""
from overiva_b200.core import auxiva, overiva  # noqa: F401
