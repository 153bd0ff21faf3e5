"""Whole-pair parity through the C ABI (tvl1_calc_u8_host) against the C oracle.
Tolerance from BASELINE.json north_star: mean EPE <= 0.01 px, max <= 0.1 px; in practice the
two are bit-identical because every stage is."""
import numpy as np
import pytest

from fibsem_optflow_b200 import synth

pytestmark = pytest.mark.gpu

MEAN_EPE_TOL = 0.01
MAX_EPE_TOL = 0.1


def run_pair(gpu, orc, I0, I1, **kw):
    s = gpu.Solver(gpu.default_params(**kw))
    u, v = s.calc(I0, I1)
    okw = {("lambda" if k == "lambda_" else k): val for k, val in kw.items() if k != "iterations"}
    okw.setdefault("inner_iterations", s.params.inner_iterations or 30)
    ou, ov, oit, olev = orc.tvl1_calc(I0, I1, **okw)
    return s, (u, v), (ou, ov, oit, olev)


@pytest.mark.parametrize("h,w,seed", [(256, 320, 7), (300, 200, 3), (64, 48, 5)])
def test_pair_defaults(gpu, orc, h, w, seed):
    I0, I1 = synth.make_pair(h, w, seed=seed)
    s, (u, v), (ou, ov, oit, olev) = run_pair(gpu, orc, I0, I1, lambda_=0.15, nscales=5,
                                             inner_iterations=30, outer_iterations=10)
    assert s.stats.levels == olev
    assert np.array_equal(s.stats.iters_array(), oit[:olev])
    epe = np.hypot(u - ou, v - ov)
    assert epe.mean() <= MEAN_EPE_TOL and epe.max() <= MAX_EPE_TOL
    assert np.array_equal(u, ou) and np.array_equal(v, ov)


def test_pair_reference_wrapper_defaults(gpu, orc):
    # generate_TV_args defaults: lambda .05, nscales 10, iterations 300 (src/optflow.cpp:503-511)
    I0, I1 = synth.make_pair(200, 260, seed=9)
    s = gpu.Solver(gpu.default_params())
    u, v = s.calc(I0, I1)
    ou, ov, oit, olev = orc.tvl1_calc(I0, I1, **{"lambda": 0.05, "nscales": 10})
    assert s.stats.levels == olev
    assert np.array_equal(s.stats.iters_array(), oit[:olev])
    epe = np.hypot(u - ou, v - ov)
    assert epe.mean() <= MEAN_EPE_TOL and epe.max() <= MAX_EPE_TOL


def test_pair_no_median_flat(gpu, orc):
    I0, I1 = synth.make_pair(128, 160, seed=2)
    s, (u, v), (ou, ov, oit, olev) = run_pair(gpu, orc, I0, I1, lambda_=0.15, nscales=4, warps=3,
                                             median_filtering=1, inner_iterations=40,
                                             outer_iterations=2)
    assert np.array_equal(s.stats.iters_array(), oit[:olev])
    assert np.array_equal(u, ou) and np.array_equal(v, ov)


def test_handle_reuse_and_resize(gpu, orc):
    s = gpu.Solver(gpu.default_params(lambda_=0.15, nscales=3))
    for (h, w, seed) in [(96, 128, 1), (96, 128, 2), (70, 90, 3)]:
        I0, I1 = synth.make_pair(h, w, seed=seed)
        u, v = s.calc(I0, I1)
        ou, ov, _, _ = orc.tvl1_calc(I0, I1, **{"lambda": 0.15, "nscales": 3})
        assert np.array_equal(u, ou) and np.array_equal(v, ov)


@pytest.mark.parametrize("h,w,seed,kw", [
    (192, 256, 7, dict(lambda_=0.15, nscales=5, gamma=0.25)),
    (150, 131, 3, dict(lambda_=0.15, nscales=3, gamma=1.5, inner_iterations=9, outer_iterations=4)),
    (96, 130, 5, dict(gamma=0.05, warps=2)),                                          # wrapper defaults otherwise
    (128, 100, 2, dict(lambda_=0.15, nscales=4, gamma=0.4, median_filtering=1)),
])
def test_gamma_solve_is_exact(gpu, orc, h, w, seed, kw):
    """gamma != 0 (the reference forwards the key, src/optflow.cpp:511,518): the third channel u3 / p31, p32;
    flow bit-equal to the oracle, same iteration counts; a brightness offset between the frames so that the
    illumination term has something to do; the handle goes back to gamma == 0 afterwards"""
    I0, I1 = synth.make_pair(h, w, seed=seed)
    I1 = np.clip(I1.astype(np.int32) + 7, 0, 255).astype(np.uint8)
    s = gpu.Solver(gpu.default_params(**kw))
    u, v = s.calc(I0, I1)
    okw = {("lambda" if k == "lambda_" else k): val for k, val in kw.items()}
    if "lambda" not in okw:
        okw.update({"lambda": 0.05, "nscales": 10})
    ou, ov, oit, lev = orc.tvl1_calc(I0, I1, **okw)
    assert s.stats.levels == lev and np.array_equal(s.stats.iters_array(), oit[:lev])
    assert np.array_equal(u, ou) and np.array_equal(v, ov)
    kw0 = dict(kw, gamma=0.0)
    s.set_params(gpu.default_params(**kw0))
    u0, v0 = s.calc(I0, I1)
    okw["gamma"] = 0.0
    ou0, ov0, _, _ = orc.tvl1_calc(I0, I1, **okw)
    assert np.array_equal(u0, ou0) and np.array_equal(v0, ov0)
    s.close()


@pytest.mark.parametrize("h,w,seed,kw", [
    (180, 240, 6, dict(lambda_=0.15, nscales=4, median_filtering=3)),
    (97, 133, 8, dict(median_filtering=3, warps=3)),
    (128, 160, 4, dict(lambda_=0.15, nscales=3, median_filtering=3, gamma=0.2, inner_iterations=10, outer_iterations=5)),
])
def test_median3_solve_is_exact(gpu, orc, h, w, seed, kw):
    """medianFiltering = 3 (cv::medianBlur's other fp32 aperture): flow bit-equal to the oracle, same iteration counts"""
    I0, I1 = synth.make_pair(h, w, seed=seed)
    s = gpu.Solver(gpu.default_params(**kw))
    u, v = s.calc(I0, I1)
    okw = {("lambda" if k == "lambda_" else k): val for k, val in kw.items()}
    if "lambda" not in okw:
        okw.update({"lambda": 0.05, "nscales": 10})
    ou, ov, oit, lev = orc.tvl1_calc(I0, I1, **okw)
    assert s.stats.levels == lev and np.array_equal(s.stats.iters_array(), oit[:lev])
    assert np.array_equal(u, ou) and np.array_equal(v, ov)
    s.close()


@pytest.mark.parametrize("h,w,seed,kw", [
    (200, 263, 6, dict(lambda_=0.15, nscales=4, scale_step=0.5)),
    (131, 97, 8, dict(scale_step=0.5, warps=3)),                       # as many halvings as fit; odd sizes on the way
    (256, 512, 4, dict(lambda_=0.15, nscales=5, scale_step=0.5, median_filtering=3)),
])
def test_scale_half_solve_is_exact(gpu, orc, h, w, seed, kw):
    """scaleStep == 0.5: the pyramid takes OpenCV's INTER_AREA fast path (2x2 means); flow bit-equal to the oracle"""
    I0, I1 = synth.make_pair(h, w, seed=seed)
    s = gpu.Solver(gpu.default_params(**kw))
    u, v = s.calc(I0, I1)
    okw = {("lambda" if k == "lambda_" else k): val for k, val in kw.items()}
    if "lambda" not in okw:
        okw.update({"lambda": 0.05, "nscales": 10})
    ou, ov, oit, lev = orc.tvl1_calc(I0, I1, **okw)
    assert s.stats.levels == lev and np.array_equal(s.stats.iters_array(), oit[:lev])
    assert np.array_equal(u, ou) and np.array_equal(v, ov)
    s.close()


def test_errors(gpu):
    import ctypes as C
    p = gpu.default_params()
    p.nscales = 0
    h = C.c_void_p()
    assert gpu.lib().tvl1_create(C.byref(p), 0, C.byref(h)) == -1
    p = gpu.default_params(gamma=float("nan"))
    assert gpu.lib().tvl1_create(C.byref(p), 0, C.byref(h)) == -1
    p = gpu.default_params(use_initial_flow=1)
    assert gpu.lib().tvl1_create(C.byref(p), 0, C.byref(h)) == -3
    p = gpu.default_params(median_filtering=7)
    assert gpu.lib().tvl1_create(C.byref(p), 0, C.byref(h)) == -3
    s = gpu.Solver(gpu.default_params())
    with pytest.raises(gpu.Tvl1Error):
        s.calc(np.zeros((8, 8), np.uint8), np.zeros((8, 9), np.uint8))


@pytest.mark.parametrize("h,w,seed,kw", [
    (256, 320, 7, dict(lambda_=0.15, nscales=5)),
    (300, 200, 3, dict(lambda_=0.15, nscales=4, inner_iterations=7, outer_iterations=12)),   # odd inner
    (200, 260, 9, dict()),                                                                   # wrapper defaults
    (128, 160, 2, dict(lambda_=0.15, nscales=3, median_filtering=1, inner_iterations=40, outer_iterations=2)),
    (96, 130, 4, dict(lambda_=0.15, nscales=2, epsilon=0.05)),                               # stops within 1-3 iterations
])
def test_fused_schedule_is_exact(gpu, orc, h, w, seed, kw):
    """The temporally blocked schedule (two iterations per launch, device-side replay when the first
    of the two already meets the stop test) must reproduce the per-iteration stop exactly."""
    I0, I1 = synth.make_pair(h, w, seed=seed)
    s = gpu.Solver(gpu.default_params(**kw))
    s.set_option("fused_min_px", 0)                      # force it on every level
    u, v = s.calc(I0, I1)
    okw = {("lambda" if k == "lambda_" else k): val for k, val in kw.items()}
    if "lambda" not in okw:
        okw.update({"lambda": 0.05, "nscales": 10})
    ou, ov, oit, olev = orc.tvl1_calc(I0, I1, **okw)
    assert s.stats.levels == olev
    assert np.array_equal(s.stats.iters_array(), oit[:olev])
    assert np.array_equal(u, ou) and np.array_equal(v, ov)
    with pytest.raises(gpu.Tvl1Error):
        s.set_option("no_such_option", 1)


def test_zero_padded_frames_exact(gpu, orc):
    """Frames with exactly flat (zero) bands, like aligned FIB-SEM slices padded with zeros: inside the
    bands the flow decays through tiny and subnormal values, which is where the kernels' fast
    arithmetic has to hand over to (or provably agree with) the IEEE operators."""
    I0, I1 = synth.make_pair(320, 384, seed=12)
    I0 = I0.copy(); I1 = I1.copy()
    I0[:90] = 0; I1[:96] = 0
    I0[:, :70] = 0; I1[:, :64] = 0
    I0[250:, 300:] = 0; I1[250:, 300:] = 0
    for fused_min in (0, 1e18):
        s = gpu.Solver(gpu.default_params(lambda_=0.15, nscales=4))
        s.set_option("fused_min_px", fused_min)
        u, v = s.calc(I0, I1)
        ou, ov, oit, olev = orc.tvl1_calc(I0, I1, **{"lambda": 0.15, "nscales": 4})
        assert np.array_equal(s.stats.iters_array(), oit[:olev])
        assert np.array_equal(u, ou) and np.array_equal(v, ov)
        s.close()
