"""The packed (FFMA2) arithmetic of the kernels is exact: the two-lane IEEE sqrt / reciprocal agree with the scalar
operators on every one of the 2^32 inputs."""
import pytest

pytestmark = pytest.mark.gpu


def test_packed_sqrt_and_rcp_exhaustive(pt):
    bad_sqrt, bad_rcp = pt.selftest_packed_math()
    assert bad_sqrt == 0 and bad_rcp == 0
