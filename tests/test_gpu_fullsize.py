"""Parity at BASELINE.json's sizes: the CUDA path against the C oracle on the SAME full-size inputs.

north_star bounds, asserted for every case: mean endpoint error <= 0.01 px, max <= 0.1 px, equal
per-(level, warp) iteration counts -- and in fact bit-equal flow planes (the kernels are
operation-for-operation the oracle's, so equality is what is asserted; the EPE bounds are kept
beside it so that the tolerance north_star states is written in the test).

* configs[0]  2048^2, 5 scales, 5 warps                      oracle: ~3 s
* configs[1]  8192^2, 6 scales, 5 warps                      oracle: ~40-60 s on 16 host cores
* configs[2]  4096^2 chained slices through tvl1_stack_run   oracle: 2 pairs, ~10 s each
* configs[4]  6144^2 pair, matches only + flow               oracle: ~25 s
* configs[3]  8 scales, 10 warps, ~11 px displacement: a 16384 x 2048 strip (full 16384 width, so
              the strip tiling, pitches and 32-bit index arithmetic of the largest config are
              exercised) against the oracle, ~60 s; the full 16384^2 pair runs all schedules against
              each other, and against the oracle when TVL1_PARITY_16K=1 (~6 min of oracle; run once per
              round, record in profiles/).
* a config[3]-style pair at 768^2 with a displacement the pyramid cannot recover: the warp kernel's
  wide-window fallback, 300-iteration (level, warp)s.
"""
import os

import numpy as np
import pytest

from fibsem_optflow_b200 import synth

pytestmark = pytest.mark.gpu


def test_config0_2048_exact(gpu, orc):
    I0, I1 = synth.make_pair(2048, 2048, seed=7)
    s = gpu.Solver(gpu.default_params(lambda_=0.15, nscales=5, inner_iterations=30, outer_iterations=10))
    u, v = s.calc(I0, I1)
    ou, ov, oit, olev = orc.tvl1_calc(I0, I1, **{"lambda": 0.15, "nscales": 5})
    assert s.stats.levels == olev == 5
    assert np.array_equal(s.stats.iters_array(), oit[:olev])
    assert np.array_equal(u, ou) and np.array_equal(v, ov)
    s.close()


def test_config1_8192_properties(gpu):
    n = 8192
    I0, I1 = synth.make_pair(n, n, seed=7, shear=4.0 / n)
    kw = dict(lambda_=0.15, nscales=6, inner_iterations=30, outer_iterations=10)
    s = gpu.Solver(gpu.default_params(**kw))
    u, v = s.calc(I0, I1)
    it_fused = s.stats.iters_array().copy()
    assert s.stats.levels == 6
    # (1) same handle, same inputs: identical
    u2, v2 = s.calc(I0, I1)
    assert np.array_equal(u, u2) and np.array_equal(v, v2)
    assert np.array_equal(it_fused, s.stats.iters_array())
    # (2) the same passes as host-driven launch slots, and one iteration per launch everywhere: identical
    s.set_option("coop_outer", 0)
    u3, v3 = s.calc(I0, I1)
    assert np.array_equal(it_fused, s.stats.iters_array())
    assert np.array_equal(u, u3) and np.array_equal(v, v3)
    s.set_option("fused_min_px", 1e18)
    u3, v3 = s.calc(I0, I1)
    assert np.array_equal(it_fused, s.stats.iters_array())
    assert np.array_equal(u, u3) and np.array_equal(v, v3)
    del u2, v2, u3, v3
    # (3) the known displacement is recovered (margins: the border has no data to match)
    ut, vt = synth.true_flow(n, n, shear=4.0 / n)
    epe = np.hypot(u - ut, v - vt)
    assert epe.mean() < 0.08
    assert epe[16:-16, 16:-16].max() < 0.5
    # (4) the stop test really decided: every (level, warp) ended before the iteration cap
    assert it_fused.max() < 300 and it_fused.min() >= 1
    s.close()


def _assert_parity(u, v, ou, ov, it, oit, lev, olev):
    assert lev == olev
    assert np.array_equal(it, oit[:olev]), "iteration-count vectors differ"
    epe = np.hypot(u - ou, v - ov)
    assert epe.mean() <= 0.01 and epe.max() <= 0.1, (float(epe.mean()), float(epe.max()))   # north_star
    assert np.array_equal(u, ou) and np.array_equal(v, ov)                                   # in fact: bit-equal


def test_config1_8192_exact(gpu, orc):
    n = 8192
    I0, I1 = synth.make_pair(n, n, seed=7, shear=4.0 / n)
    s = gpu.Solver(gpu.default_params(lambda_=0.15, nscales=6, inner_iterations=30, outer_iterations=10))
    u, v = s.calc(I0, I1)
    ou, ov, oit, olev = orc.tvl1_calc(I0, I1, **{"lambda": 0.15, "nscales": 6})
    _assert_parity(u, v, ou, ov, s.stats.iters_array(), oit, s.stats.levels, olev)
    # match coordinates from the two flows: bit-equal
    got = s.sample_matches(I0, I1, u, v, scale=0.5, npoints=25, seed=3)
    want = orc.random_points(I0, I1, ou, ov, scale=0.5, npoints=25, seed=3)
    for k in range(6):
        assert np.array_equal(got[k], want[k])
    s.close()


def test_config2_4096_chained_stack_exact(gpu, orc):
    n = 4096
    sl = synth.make_stack(2, n, n, seed=100)            # 3 slices -> 2 chained pairs
    s = gpu.Solver(gpu.default_params(lambda_=0.15, nscales=5, inner_iterations=30, outer_iterations=10))
    res = s.run_stack(sl, flows=True, apply_mask=True, npoints=25, scale=0.5, seed=1)
    for k in range(2):
        ou, ov, oit, olev = orc.tvl1_calc(sl[k], sl[k + 1], **{"lambda": 0.15, "nscales": 5})
        orc.mask_flow(sl[k + 1], ou, ov)
        st = res["stats"][k]
        _assert_parity(res["u"][k], res["v"][k], ou, ov, st.iters_array(), oit, st.levels, olev)
        want = orc.random_points(sl[k], sl[k + 1], ou, ov, scale=0.5, npoints=25, seed=1)
        for j in range(5):
            assert np.array_equal(res["matches"][k][j], want[j])
    s.close()


def test_config4_6144_pair_exact(gpu, orc):
    n = 6144
    sl = synth.make_stack(1, n, n, seed=200)
    s = gpu.Solver(gpu.default_params(lambda_=0.15, nscales=5, inner_iterations=30, outer_iterations=10))
    # the full-volume run keeps the flow on the device and downloads matches only (h_u = NULL)
    res = s.run_stack(sl, flows=False, apply_mask=True, npoints=25, scale=0.5, seed=5)
    u, v = s.calc(sl[0], sl[1])
    ou, ov, oit, olev = orc.tvl1_calc(sl[0], sl[1], **{"lambda": 0.15, "nscales": 5})
    _assert_parity(u, v, ou, ov, s.stats.iters_array(), oit, s.stats.levels, olev)
    assert np.array_equal(res["stats"][0].iters_array(), oit[:olev])
    orc.mask_flow(sl[1], ou, ov)
    want = orc.random_points(sl[0], sl[1], ou, ov, scale=0.5, npoints=25, seed=5)
    for j in range(5):
        assert np.array_equal(res["matches"][0][j], want[j])
    s.close()


C3 = dict(seed=13, dx=9.0, dy=-6.0, margin=64, coarse=16)      # ~11 px, recoverable with 8 scales
C3_KW = {"lambda": 0.15, "nscales": 8, "warps": 10}


def test_config3_16384_wide_strip_exact(gpu, orc):
    h, w = 2048, 16384
    I0, I1 = synth.make_pair(h, w, shear=0.0005, **C3)
    s = gpu.Solver(gpu.default_params(lambda_=0.15, nscales=8, warps=10))
    u, v = s.calc(I0, I1)
    ou, ov, oit, olev = orc.tvl1_calc(I0, I1, **C3_KW)
    _assert_parity(u, v, ou, ov, s.stats.iters_array(), oit, s.stats.levels, olev)
    assert s.stats.iters_array()[0].max() < 300          # the stop test decided at the finest level
    s.close()


def test_config3_16384_square(gpu, orc):
    """The full 16384^2 pair (25 GB arena): every schedule agrees bit for bit, the displacement is
    recovered; with TVL1_PARITY_16K=1 also the direct oracle comparison (minutes of host time)."""
    n = 16384
    I0, I1 = synth.make_pair(n, n, shear=0.5 / n, **C3)
    s = gpu.Solver(gpu.default_params(lambda_=0.15, nscales=8, warps=10))
    u, v = s.calc(I0, I1)
    it = s.stats.iters_array().copy()
    assert s.stats.levels == 8 and it[0].max() < 300
    s.set_option("coop_outer", 0)
    s.set_option("fused_min_px", 1e18)                    # one iteration per launch everywhere
    u2, v2 = s.calc(I0, I1)
    assert np.array_equal(it, s.stats.iters_array())
    assert np.array_equal(u, u2) and np.array_equal(v, v2)
    del u2, v2
    s.close()
    ut, vt = synth.true_flow(n, n, dx=C3["dx"], dy=C3["dy"], shear=0.5 / n, margin=C3["margin"])
    epe = np.hypot(u - ut, v - vt)
    assert epe.mean() < 0.15, float(epe.mean())
    del ut, vt, epe
    if os.environ.get("TVL1_PARITY_16K") == "1":
        ou, ov, oit, olev = orc.tvl1_calc(I0, I1, **C3_KW)
        _assert_parity(u, v, ou, ov, it, oit, 8, olev)


def test_config3_large_displacement_exact(gpu, orc):
    h = w = 768
    I0, I1 = synth.make_pair(h, w, seed=13, dx=13.4, dy=-9.2, shear=0.008, margin=64)
    kw = {"lambda": 0.15, "nscales": 8, "warps": 10}
    s = gpu.Solver(gpu.default_params(lambda_=0.15, nscales=8, warps=10))
    u, v = s.calc(I0, I1)
    ou, ov, oit, olev = orc.tvl1_calc(I0, I1, **kw)
    assert s.stats.levels == olev
    assert np.array_equal(s.stats.iters_array(), oit[:olev])
    assert np.array_equal(u, ou) and np.array_equal(v, ov)
    s.close()
