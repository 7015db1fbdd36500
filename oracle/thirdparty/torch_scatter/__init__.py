"""ORACLE / TEST INFRASTRUCTURE — stand-in for torch-scatter==2.0.5 (requirements.txt:30).
Only `composite.scatter_softmax` is on the reference path (molgym/modules.py:27)."""
from . import composite  # noqa: F401
from .composite import scatter_softmax  # noqa: F401
