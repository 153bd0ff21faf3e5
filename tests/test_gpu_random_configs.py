"""Randomised parity: image sizes (odd, tiny, wide, tall), every solver parameter the reference's
generate_TV_args exposes (src/optflow.cpp:500-514) plus scaleStep / inner / outer / medianFiltering, with
and without zero bands, all four iteration schedules -- each case bit-exact against the oracle, iteration
counts included.  Seeds are fixed so that a failure reproduces."""
import numpy as np
import pytest

from fibsem_optflow_b200 import synth

pytestmark = pytest.mark.gpu


def draw(rng):
    h = int(rng.integers(20, 420))
    w = int(rng.integers(20, 520))
    kw = dict(
        tau=float(rng.choice([0.25, 0.2, 0.1, 0.3])),
        lambda_=float(rng.choice([0.15, 0.05, 0.3, 0.6])),
        theta=float(rng.choice([0.3, 0.2, 0.5])),
        nscales=int(rng.integers(1, 8)),
        warps=int(rng.integers(1, 7)),
        epsilon=float(rng.choice([0.01, 0.02, 0.005, 0.05])),
        scale_step=float(rng.choice([0.8, 0.7, 0.9, 0.6])),
        inner_iterations=int(rng.choice([30, 7, 12, 2, 1, 33])),
        outer_iterations=int(rng.choice([10, 3, 1, 5])),
        median_filtering=int(rng.choice([5, 5, 1])),
    )
    pair = dict(seed=int(rng.integers(1, 10_000)), dx=float(rng.uniform(-3, 3)), dy=float(rng.uniform(-3, 3)),
                shear=float(rng.uniform(-0.01, 0.01)), sigma=float(rng.choice([1.5, 2.0, 3.0])))
    bands = bool(rng.integers(0, 3) == 0)
    # drawn last, so that the cases above keep their values: the INTER_AREA pyramid, the 3x3 median, gamma
    if rng.random() < 0.2:
        kw["scale_step"] = 0.5
    if rng.random() < 0.2 and kw["median_filtering"] == 5:
        kw["median_filtering"] = 3
    if rng.random() < 0.15:
        kw["gamma"] = float(rng.choice([0.1, 0.5]))
    return h, w, kw, pair, bands


@pytest.mark.parametrize("h,w", [(1, 1), (1, 9), (2, 3), (4, 4), (3, 130), (130, 3), (5, 300), (16, 17), (15, 40), (33, 1)])
@pytest.mark.parametrize("gamma", [0.0, 0.3])
def test_degenerate_sizes(gpu, orc, h, w, gamma):
    """frames smaller than a tile, a strip, a median window or the pyramid's 16-px stop rule (one level only):
    every schedule bit-exact against the oracle"""
    rng = np.random.default_rng(h * 1000 + w)
    I0 = rng.integers(0, 256, (h, w), dtype=np.uint8)
    I1 = np.roll(I0, 1, axis=1) if w > 1 else I0.copy()
    I1 = (I1.astype(np.int32) * 7 // 8 + 3).astype(np.uint8)
    kw = dict(lambda_=0.15, nscales=5, warps=3, inner_iterations=5, outer_iterations=3, gamma=gamma)
    okw = {("lambda" if k == "lambda_" else k): v for k, v in kw.items()}
    ou, ov, oit, olev = orc.tvl1_calc(I0, I1, **okw)
    for fused_min, multi, coop in ((0, 1, 1), (0, 1, 0), (1e18, 1, 1), (1e18, 0, 1)):
        s = gpu.Solver(gpu.default_params(**kw))
        s.set_option("fused_min_px", fused_min)
        s.set_option("multi_iter", multi)
        s.set_option("coop_outer", coop)
        u, v = s.calc(I0, I1)
        assert s.stats.levels == olev
        assert np.array_equal(s.stats.iters_array(), oit[:olev]), (h, w, fused_min, multi, coop)
        assert np.array_equal(u, ou) and np.array_equal(v, ov), (h, w, fused_min, multi, coop)
        s.close()
        if gamma != 0.0:
            break          # the three-channel iteration has one schedule


@pytest.mark.parametrize("case", range(24))
def test_random_config(gpu, orc, case):
    rng = np.random.default_rng(1000 + case)
    h, w, kw, pair, bands = draw(rng)
    I0, I1 = synth.make_pair(h, w, **pair)
    if bands:
        I0 = I0.copy(); I1 = I1.copy()
        I0[: h // 4] = 0; I1[: h // 4 + 2] = 0
        I0[:, -(w // 5):] = 0; I1[:, -(w // 5):] = 0
    okw = {("lambda" if k == "lambda_" else k): v for k, v in kw.items()}
    ou, ov, oit, olev = orc.tvl1_calc(I0, I1, **okw)
    # every schedule: fused + single passes in one cooperative launch per outer iteration / the same passes
    # as host-driven launch slots / single passes in one cooperative launch / one iteration per launch
    for fused_min, multi, coop in ((0, 1, 1), (0, 1, 0), (1e18, 1, 1), (1e18, 0, 1)):
        s = gpu.Solver(gpu.default_params(**kw))
        s.set_option("fused_min_px", fused_min)
        s.set_option("multi_iter", multi)
        s.set_option("coop_outer", coop)
        u, v = s.calc(I0, I1)
        assert s.stats.levels == olev, (case, kw)
        assert np.array_equal(s.stats.iters_array(), oit[:olev]), (case, kw, h, w)
        assert np.array_equal(u, ou) and np.array_equal(v, ov), (case, kw, h, w, fused_min)
        s.close()
