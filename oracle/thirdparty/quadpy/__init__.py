"""ORACLE / TEST INFRASTRUCTURE — stand-in for quadpy==0.16.2 (requirements.txt:23).
Only `quadpy.u3._lebedev.lebedev_071()` is used (molgym/agents/covariant/spherical_dists.py:209-212)."""
from . import u3  # noqa: F401
