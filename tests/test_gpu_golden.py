"""CUDA path against the committed golden fixtures (cv2 / glibc derived, tests/golden/)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_primitives_golden(gpu, orc):
    g = np.load(os.path.join(GOLD, "primitives.npz"))
    src = g["src"]
    h, w = src.shape
    assert np.array_equal(gpu.k_resize(src, inv_scale=0.8), g["down"])
    assert np.array_equal(gpu.k_resize(g["down"], dw=w, dh=h), g["up"])
    assert np.array_equal(gpu.k_median5(g["med_in"]), g["med"])
    # the warp kernel's I1w output with (I1 := src, u := map - grid) is remap(src)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    u1 = (g["mx"] - xx).astype(np.float32)
    u2 = (g["my"] - yy).astype(np.float32)
    ok = ((xx + u1) == g["mx"]) & ((yy + u2) == g["my"])   # fp32 round trip of the map
    z = np.zeros_like(src)
    iw = gpu.k_warp(z, src, u1, u2)[0]
    assert ok.mean() > 0.5
    assert np.array_equal(iw[ok], g["remap"][ok])


def test_pair_golden(gpu):
    g = np.load(os.path.join(GOLD, "pair_96x128.npz"))
    s = gpu.Solver(gpu.default_params(lambda_=0.15, nscales=5))
    u, v = s.calc(g["I0"], g["I1"])
    assert np.array_equal(s.stats.iters_array(), g["iters"])
    epe = np.hypot(u - g["u"], v - g["v"])
    assert epe.mean() <= 0.01 and epe.max() <= 0.1       # north_star tolerance
    assert np.array_equal(u, g["u"]) and np.array_equal(v, g["v"])
    s2 = gpu.Solver(gpu.default_params())                 # reference wrapper defaults
    u, v = s2.calc(g["I0"], g["I1"])
    assert np.array_equal(s2.stats.iters_array(), g["iters_ref"])
    assert np.array_equal(u, g["u_ref"]) and np.array_equal(v, g["v_ref"])


def test_round2_golden(gpu):
    """3x3 median, scale-0.5 pyramid step and the whole-pair flows with medianFiltering 3 / scaleStep 0.5
    against the cv2-made vectors of tests/golden/make_golden.py --round2"""
    g = np.load(os.path.join(GOLD, "primitives2.npz"))
    assert np.array_equal(gpu.k_median3(g["med3_in"]), g["med3"])
    for k in range(5):
        assert np.array_equal(gpu.k_resize(g["half_in_%d" % k], inv_scale=0.5), g["half_%d" % k]), k
    for tag, kw in (("med3", dict(median_filtering=3)), ("half", dict(scale_step=0.5))):
        s = gpu.Solver(gpu.default_params(lambda_=0.15, nscales=4, **kw))
        u, v = s.calc(g["I0"], g["I1"])
        assert np.array_equal(s.stats.iters_array(), g["it_" + tag]), tag
        assert np.array_equal(u, g["u_" + tag]) and np.array_equal(v, g["v_" + tag]), tag
        s.close()
