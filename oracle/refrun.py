"""ORACLE / TEST INFRASTRUCTURE — make the reference's own `molgym/` importable VERBATIM on top of the
restated third-party stand-ins in oracle/thirdparty/.

Only works where /root/reference exists (the build container).  Nothing on the GPU box may call
`enable(require_reference=True)`; the self-contained oracle in oracle/molgym_oracle/ is what travels.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
THIRDPARTY = os.path.join(HERE, 'thirdparty')
REFERENCE = os.environ.get('MOLGYM_REFERENCE', '/root/reference')


def enable_thirdparty():
    """Stand-ins on sys.path + the numpy-1.x aliases the reference still uses (molgym/spaces.py:27-29,
    molgym/agents/internal/zmat.py:110, molgym/agents/covariant/spherical_dists.py:131)."""
    if THIRDPARTY not in sys.path:
        sys.path.insert(0, THIRDPARTY)
    for alias, target in (('float', float), ('bool', bool), ('int', int), ('product', np.prod)):
        if not hasattr(np, alias):
            setattr(np, alias, target)


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE, 'molgym'))


def enable(require_reference=True):
    enable_thirdparty()
    if reference_available():
        if REFERENCE not in sys.path:
            sys.path.append(REFERENCE)   # at the END: only `molgym` is wanted from it (its `tests` package must not shadow this repo's)
        return True
    if require_reference:
        raise RuntimeError(f'reference tree not found at {REFERENCE}')
    return False
