"""Pins of the Clebsch-Gordan coefficients against an independent implementation (sympy.physics.quantum.cg.CG), for every
(l1 m1, l2 m2 | l m) with l1, l2, l <= 4 — the only ones the path uses (maxl = 4, molgym/tools/arg_parser.py:56):
  * the oracle's restatement of cormorant.cg_lib (oracle/thirdparty/cormorant/cg_lib.py::clebsch), on which every golden
    vector rests;
  * the product's own host-side table builder (csrc/model.cuh::clebsch_gordan through the C ABI's mgb_clebsch_gordan).
Both must equal the exact values to double precision, including the selection rules (zeros)."""
import ctypes
import itertools

import pytest

from molgym_b200 import _cabi, build
from oracle import refrun

MAXL = 4


@pytest.fixture(scope='module')
def exact():
    from sympy import S
    from sympy.physics.quantum.cg import CG
    table = {}
    for l1, l2 in itertools.product(range(MAXL + 1), repeat=2):
        for l in range(abs(l1 - l2), min(l1 + l2, MAXL) + 1):
            for m1 in range(-l1, l1 + 1):
                for m2 in range(-l2, l2 + 1):
                    m = m1 + m2
                    if abs(m) <= l:
                        table[(l1, m1, l2, m2, l, m)] = float(CG(S(l1), S(m1), S(l2), S(m2), S(l), S(m)).doit())
    return table


def test_oracle_clebsch_equals_sympy(exact):
    refrun.enable_thirdparty()
    from cormorant.cg_lib import clebsch
    assert len(exact) > 1400
    for key, value in exact.items():
        assert abs(clebsch(*key) - value) < 1e-13, key
    assert clebsch(1, 1, 1, 1, 1, 1) == 0.0 and clebsch(2, 0, 2, 1, 5, 1) == 0.0      # m1 + m2 != m, l > l1 + l2


def test_product_table_builder_equals_sympy(exact):
    lib = _cabi.bind(ctypes.CDLL(build.build_cuda()))    # host function of the nvcc-built product library: no GPU needed
    nonzero = 0
    for key, value in exact.items():
        got = lib.mgb_clebsch_gordan(*key)
        assert abs(got - value) < 1e-13, (key, got, value)
        nonzero += abs(value) > 1e-12
    assert nonzero >= 1392                                 # 1 392 non-zero coefficients for the 65 paths
    assert lib.mgb_clebsch_gordan(1, 1, 1, 1, 1, 1) == 0.0 and lib.mgb_clebsch_gordan(4, 0, 4, 0, 9, 0) == 0.0


def test_oracle_cg_dictionary_blocks_use_those_coefficients(exact):
    """cg_matrix stacks the l = |l1-l2| .. l1+l2 blocks (rows (l, m), columns (m1, m2)) that CGDict serves to cg_product."""
    refrun.enable_thirdparty()
    from cormorant.cg_lib import cg_matrix
    for l1, l2 in ((1, 1), (2, 1), (2, 2), (4, 3)):
        mat = cg_matrix(l1, l2)
        lmin = abs(l1 - l2)
        for l in range(lmin, min(l1 + l2, MAXL) + 1):
            for m1 in range(-l1, l1 + 1):
                for m2 in range(-l2, l2 + 1):
                    m = m1 + m2
                    if abs(m) <= l:
                        got = float(mat[l * l - lmin * lmin + l + m, (l1 + m1) * (2 * l2 + 1) + (l2 + m2)])
                        assert abs(got - exact[(l1, m1, l2, m2, l, m)]) < 1e-13
