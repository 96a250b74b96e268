"""GPU: the CUDA path, through the C ABI, against the reference's own golden vectors — bit-exact."""
import pytest

from replay import load_golden, replay

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("test", load_golden(), ids=lambda t: t["name"])
def test_b200_matches_reference_tests(test):
    from libfriendship_b200 import B200Renderer
    replay(B200Renderer(), test)
