"""ORACLE / TEST INFRASTRUCTURE — run the reference's OWN hot-path tests, unmodified, on the stand-ins.
Usage: python oracle/run_reference_tests.py [pytest args]   (build container only: needs /root/reference)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refrun  # noqa: E402

HOT_PATH_TESTS = [
    'tests/agents/covariant/test_sphs.py',
    'tests/agents/covariant/test_so3_tools.py',
    'tests/agents/covariant/test_spherical_distr.py',
    'tests/agents/covariant/test_gmm.py',
    'tests/agents/covariant/test_tools.py',
    'tests/agents/covariant/test_agent.py',
    'tests/agents/internal/test_zmat.py',
    'tests/test_modules.py',
    'tests/test_spaces.py',
    'tests/test_tools.py',
]

if __name__ == '__main__':
    refrun.enable(require_reference=True)
    import pytest
    os.chdir(refrun.REFERENCE)
    args = sys.argv[1:] or HOT_PATH_TESTS
    sys.exit(pytest.main(['-q', '-p', 'no:cacheprovider', '--rootdir', refrun.REFERENCE] + args))
