"""ORACLE / TEST INFRASTRUCTURE — degree-71 Lebedev rule (1730 points).  quadpy's scheme exposes
`.points` [3, 1730] and `.weights` [1730] summing to 1; scipy.integrate.lebedev_rule(71) is the same rule
with weights summing to 4*pi."""
import functools
import math
from types import SimpleNamespace

from scipy.integrate import lebedev_rule


@functools.lru_cache(maxsize=None)
def lebedev_071():
    points, weights = lebedev_rule(71)
    return SimpleNamespace(points=points, weights=weights / (4 * math.pi), degree=71, name='lebedev_071')
