"""ORACLE / TEST INFRASTRUCTURE — self-contained CPU restatement of the reference hot path
(molgym's covariant and internal actor-critics + PPO loss).  Travels to the GPU box; checked in the build
container against the reference's own code run verbatim (tests/test_oracle_vs_reference.py) and against the
committed golden vectors (tests/golden/)."""
from oracle import refrun as _refrun

_refrun.enable_thirdparty()

from .covariant import CovariantOracle, pack_observations  # noqa: E402,F401
from .internal import SchNetOracle  # noqa: E402,F401
from .ppo import ppo_loss  # noqa: E402,F401
