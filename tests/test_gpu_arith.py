"""The iteration kernels replace div.rn / sqrt.rn by their own fast paths (shared reciprocal, fp64
Goldschmidt hypot, fp32-only hypot with a tie guard) and one range test per row; here those sequences
are compared with the IEEE operators on ~10^9 random operand triples per exponent range (bit-exact,
zero mismatches allowed), and the fp32 hypot must vouch for all but a sliver of the operands."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("elo,ehi,seed", [(-12, 12, 1), (-40, 40, 2), (-59, 59, 3), (-3, 3, 4), (-126, 127, 5),
                                          (-24, 2, 6), (0, 0, 7)])
def test_fast_paths_match_ieee(gpu, elo, ehi, seed):
    n = 1 << 28
    bad, unvouched = gpu.selftest_arith(n, seed=seed, elo=elo, ehi=ehi, with_unvouched=True)
    assert bad == 0
    if ehi <= 12 and elo >= -24:      # flow differences live here: the exact fallback must stay rare
        assert unvouched < n * 1e-4, unvouched / n
