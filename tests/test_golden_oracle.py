"""CPU: the oracle (C++ restatement of RefRenderer) against every golden vector the reference's tests hold
(tests/golden/make_reference_tests.py cites each test's file:line)."""
import pytest

from oracle.binding import OracleRenderer
from replay import load_golden, replay


@pytest.mark.parametrize("test", load_golden(), ids=lambda t: t["name"])
def test_oracle_matches_reference_tests(test):
    replay(OracleRenderer(), test)


def test_golden_has_all_eleven():
    names = {t["name"] for t in load_golden()}
    assert len(names) == 11
