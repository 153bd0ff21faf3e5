"""Stage-level parity: each CUDA kernel, called through the C ABI, against the C oracle on
the same seeded inputs.  Bar: bit-exact (all stages are fp32 with one rounding per op)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIZES = [(16, 16), (37, 53), (64, 124), (65, 125), (97, 249), (130, 500), (240, 1000)]


def rnd(rng, h, w, scale=1.0):
    return (rng.standard_normal((h, w)) * scale).astype(np.float32)


def test_convert_u8(gpu, orc):
    rng = np.random.default_rng(0)
    for h, w in SIZES:
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert np.array_equal(gpu.k_convert_u8(img), img.astype(np.float32))


@pytest.mark.parametrize("h,w", SIZES + [(333, 517)])
def test_resize_down_and_up(gpu, orc, h, w):
    rng = np.random.default_rng(h * 1000 + w)
    src = (rng.random((h, w)) * 255).astype(np.float32)
    want = orc.resize_scale(src, 0.8)
    got = gpu.k_resize(src, inv_scale=0.8)
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    # flow upsample: explicit size, times 1/scaleStep
    want_up = orc.resize_to(want, w, h) * np.float32(1 / 0.8)
    got_up = gpu.k_resize(want, dw=w, dh=h, mul=float(np.float32(1 / 0.8)))
    assert np.array_equal(got_up, want_up)


@pytest.mark.parametrize("h,w", SIZES + [(2, 5), (3, 7), (31, 17), (33, 130)])
def test_resize_half(gpu, orc, h, w):
    """scaleStep == 0.5: OpenCV's INTER_AREA fast path (2x2 means; pairwise / sequential sums, partial border blocks)"""
    rng = np.random.default_rng(h * 1000 + w + 7)
    src = (rng.random((h, w)) * 255).astype(np.float32)
    want = orc.resize_scale(src, 0.5)
    got = gpu.k_resize(src, inv_scale=0.5)
    assert got.shape == want.shape
    assert np.array_equal(got, want)


@pytest.mark.parametrize("h,w", SIZES)
def test_centered_gradient(gpu, orc, h, w):
    rng = np.random.default_rng(1)
    src = (rng.random((h, w)) * 255).astype(np.float32)
    wx, wy = orc.centered_gradient(src)
    gx, gy = gpu.k_centered_gradient(src)
    assert np.array_equal(gx, wx) and np.array_equal(gy, wy)


@pytest.mark.parametrize("h,w", SIZES)
@pytest.mark.parametrize("amp", [0.7, 4.0, 60.0])
def test_warp(gpu, orc, h, w, amp):
    rng = np.random.default_rng(2)
    I0 = (rng.random((h, w)) * 255).astype(np.float32)
    I1 = (rng.random((h, w)) * 255).astype(np.float32)
    I1x, I1y = orc.centered_gradient(I1)
    u1, u2 = rnd(rng, h, w, amp), rnd(rng, h, w, amp)
    ww, wx, wy, g, r = orc.warp(I0, I1, I1x, I1y, u1, u2)
    gw, gx, gy, gg, gr = gpu.k_warp(I0, I1, u1, u2)
    assert np.array_equal(gw, ww)
    assert np.array_equal(gx, wx)
    assert np.array_equal(gy, wy)
    assert np.array_equal(gg, g)
    assert np.array_equal(gr, r)


@pytest.mark.parametrize("h,w", SIZES)
def test_median5(gpu, orc, h, w):
    rng = np.random.default_rng(3)
    src = rnd(rng, h, w, 3.0)
    assert np.array_equal(gpu.k_median5(src), orc.median5(src))


@pytest.mark.parametrize("h,w", SIZES + [(1, 1), (1, 9), (7, 1), (3, 2)])
def test_median3(gpu, orc, h, w):
    """the 3x3 aperture (medianFiltering = 3), ties and zeros included"""
    rng = np.random.default_rng(23)
    src = rnd(rng, h, w, 3.0)
    src[rng.random((h, w)) < 0.2] = 0
    src[rng.random((h, w)) < 0.1] = 1.5
    assert np.array_equal(gpu.k_median3(src), orc.median3(src))


def make_iter_inputs(rng, h, w):
    I1wx, I1wy = rnd(rng, h, w, 8.0), rnd(rng, h, w, 8.0)
    # some exactly-zero gradients to reach the `grad > FLT_EPSILON` branch
    z = rng.random((h, w)) < 0.05
    I1wx[z] = 0
    I1wy[z] = 0
    grad = I1wx * I1wx + I1wy * I1wy
    rho_c = rnd(rng, h, w, 20.0)
    state = [rnd(rng, h, w, 0.8) for _ in range(2)] + [rnd(rng, h, w, 0.4) for _ in range(4)]
    return (I1wx, I1wy, grad, rho_c), state


@pytest.mark.parametrize("h,w", SIZES)
@pytest.mark.parametrize("n", [1, 2, 7])
def test_iterate(gpu, orc, h, w, n):
    rng = np.random.default_rng(4)
    consts, state = make_iter_inputs(rng, h, w)
    l_t, theta, taut = np.float32(0.15 * 0.3), np.float32(0.3), np.float32(0.25 / 0.3)
    want = [s.copy() for s in state]
    werr = [orc.iterate(*consts, *want, l_t, theta, taut) for _ in range(n)]
    got = gpu.k_iterate(*consts, *state, l_t, theta, taut, n=n)
    names = ["u1", "u2", "p11", "p12", "p21", "p22"]
    for k in range(6):
        assert np.array_equal(got[k], want[k]), names[k]
    np.testing.assert_allclose(got[6], werr, rtol=1e-12)


@pytest.mark.parametrize("h,w", SIZES + [(200, 700)])
@pytest.mark.parametrize("n", [2, 6])
def test_iterate_fused2(gpu, orc, h, w, n):
    """temporally blocked kernel (two iterations per launch): same states, same per-iteration errors"""
    rng = np.random.default_rng(5)
    consts, state = make_iter_inputs(rng, h, w)
    l_t, theta, taut = np.float32(0.15 * 0.3), np.float32(0.3), np.float32(0.25 / 0.3)
    want = [s.copy() for s in state]
    werr = [orc.iterate(*consts, *want, l_t, theta, taut) for _ in range(n)]
    got = gpu.k_iterate(*consts, *state, l_t, theta, taut, n=n, fused=True)
    names = ["u1", "u2", "p11", "p12", "p21", "p22"]
    for k in range(6):
        assert np.array_equal(got[k], want[k]), names[k]
    np.testing.assert_allclose(got[6], werr, rtol=1e-12)


@pytest.mark.parametrize("h,w", SIZES + [(200, 700)])
@pytest.mark.parametrize("n", [1, 4])
def test_iterate_gamma(gpu, orc, h, w, n):
    """the three-channel iteration (gamma != 0): same states and per-iteration errors as the oracle"""
    rng = np.random.default_rng(15)
    consts, state = make_iter_inputs(rng, h, w)
    I1wx, I1wy, grad, rho_c = consts
    u3, p31, p32 = rnd(rng, h, w, 0.6), rnd(rng, h, w, 0.4), rnd(rng, h, w, 0.4)
    st = [state[0], state[1], u3] + state[2:] + [p31, p32]
    l_t, theta, taut, gamma = np.float32(0.15 * 0.3), np.float32(0.3), np.float32(0.25 / 0.3), np.float32(0.35)
    want = [s.copy() for s in st]
    werr = [orc.iterate_gamma(I1wx, I1wy, grad, rho_c, *want, l_t, theta, taut, gamma) for _ in range(n)]
    got = gpu.k_iterate_gamma(I1wx, I1wy, rho_c, *st, l_t, theta, taut, gamma, n=n)
    names = ["u1", "u2", "u3", "p11", "p12", "p21", "p22", "p31", "p32"]
    for k in range(9):
        assert np.array_equal(got[k], want[k]), names[k]
    np.testing.assert_allclose(got[9], werr, rtol=1e-12)


@pytest.mark.parametrize("h,w", SIZES[:2] + [(200, 700)])
@pytest.mark.parametrize("n", [1, 5, 8])
def test_iterate_outer(gpu, orc, h, w, n):
    """k_outer, the kernel the solver ships: n iterations in one cooperative launch (two-iteration
    passes + a single pass for odd n): same states, same per-iteration errors"""
    rng = np.random.default_rng(6)
    consts, state = make_iter_inputs(rng, h, w)
    l_t, theta, taut = np.float32(0.15 * 0.3), np.float32(0.3), np.float32(0.25 / 0.3)
    want = [s.copy() for s in state]
    werr = [orc.iterate(*consts, *want, l_t, theta, taut) for _ in range(n)]
    got = gpu.k_iterate(*consts, *state, l_t, theta, taut, n=n, fused="outer")
    names = ["u1", "u2", "p11", "p12", "p21", "p22"]
    for k in range(6):
        assert np.array_equal(got[k], want[k]), names[k]
    np.testing.assert_allclose(got[6], werr, rtol=1e-12)


def test_prescale_u8(gpu, orc):
    """tvl1_prescale_u8: the loader's 8-bit cv::resize on the device -- cv2-made golden vectors and,
    at a realistic size, the oracle."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prescale.npz"))
    for k in range(int(g["n"])):
        got = gpu.prescale_u8(g["src_%d" % k], float(g["scale_%d" % k]))
        assert got.shape == g["dst_%d" % k].shape and np.array_equal(got, g["dst_%d" % k]), k
    rng = np.random.default_rng(8)
    src = rng.integers(0, 256, size=(1531, 2049), dtype=np.uint8)
    for sc in (0.5, 0.37, 0.25, 0.9):
        scf = float(np.float32(sc))
        assert np.array_equal(gpu.prescale_u8(src, scf), orc.prescale_u8(src, scf)), sc
    with pytest.raises(gpu.Tvl1Error):
        gpu.prescale_u8(src[:4, :4], 0.01)
