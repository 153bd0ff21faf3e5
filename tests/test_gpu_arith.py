"""The iteration kernel replaces div.rn / sqrt.rn by their own fast paths with a shared
reciprocal and one range guard per pixel; here those sequences are compared with the IEEE
operators on ~10^9 random operand triples (bit-exact, zero mismatches allowed)."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("elo,ehi,seed", [(-12, 12, 1), (-40, 40, 2), (-59, 59, 3), (-3, 3, 4), (-126, 127, 5)])
def test_fast_paths_match_ieee(gpu, elo, ehi, seed):
    assert gpu.selftest_arith(1 << 28, seed=seed, elo=elo, ehi=ehi) == 0
