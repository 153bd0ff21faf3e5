"""Subsystem (4): device-side match sampling against the oracle's literal random_points
(glibc rand() + std::random_shuffle over all mask pixels).  Bar: bit-exact positions and
bit-exact p/q given the same flow."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def frames(rng, h, w, zero_frac):
    f0 = rng.integers(2, 256, (h, w), dtype=np.uint8)
    f1 = rng.integers(2, 256, (h, w), dtype=np.uint8)
    z0 = rng.random((h, w)) < zero_frac
    z1 = rng.random((h, w)) < zero_frac
    f0[z0] = rng.integers(0, 2, int(z0.sum()), dtype=np.uint8)
    f1[z1] = rng.integers(0, 2, int(z1.sum()), dtype=np.uint8)
    return f0, f1


@pytest.mark.parametrize("h,w,zf,npts,seed", [
    (64, 80, 0.0, 25, 1), (64, 80, 0.7, 25, 12345), (37, 53, 0.3, 25, 1539000000),
    (5, 5, 0.0, 25, 7), (3, 4, 0.5, 25, 9), (1, 1, 0.0, 25, 3), (300, 333, 0.9, 100, 42),
    (512, 777, 0.2, 25, 2), (64, 64, 0.0, 1, 5), (2000, 1500, 0.5, 25, 99), (40, 40, 0.0, 1600, 11),
])
def test_sample_matches(gpu, orc, h, w, zf, npts, seed):
    rng = np.random.default_rng(seed)
    f0, f1 = frames(rng, h, w, zf)
    u = (rng.standard_normal((h, w)) * 2).astype(np.float32)
    v = (rng.standard_normal((h, w)) * 2).astype(np.float32)
    s = gpu.Solver(gpu.default_params())
    kw = dict(roi0=(3, 11), roi1=(5, 2), scale=0.5, npoints=npts, seed=seed)
    got = s.sample_matches(f0, f1, u, v, **kw)
    want = orc.random_points(f0, f1, u, v, **kw)
    assert np.array_equal(got[5], want[5])                       # positions
    for k in range(5):
        assert np.array_equal(got[k], want[k])                   # px py qx qy w (bit-exact doubles)


def test_sample_empty_mask_dummy(gpu):
    f = np.ones((20, 30), np.uint8)
    z = np.zeros((20, 30), np.float32)
    s = gpu.Solver(gpu.default_params())
    px, py, qx, qy, w, pos = s.sample_matches(f, f, z, z, npoints=25, seed=1)
    assert px.tolist() == [-1.0] and py.tolist() == [-1.0] and qx.tolist() == [-1.0]
    assert qy.tolist() == [-1.0] and w.tolist() == [0.0]


def test_mask_flow(gpu, orc):
    rng = np.random.default_rng(0)
    h, w = 45, 70
    f1 = rng.integers(0, 4, (h, w), dtype=np.uint8)
    u = rng.standard_normal((h, w)).astype(np.float32)
    v = rng.standard_normal((h, w)).astype(np.float32)
    bu = gpu.DevBuf(u.nbytes).upload(u)
    bv = gpu.DevBuf(v.nbytes).upload(v)
    bf = gpu.DevBuf(f1.nbytes).upload(f1)
    s = gpu.Solver(gpu.default_params())
    s.mask_flow_device(bf.ptr, w, w, h, bu.ptr, bv.ptr, w * 4)
    gu, gv = bu.download((h, w), np.float32), bv.download((h, w), np.float32)
    orc.mask_flow(f1, u, v)
    assert np.array_equal(gu, u) and np.array_equal(gv, v)


def test_solve_wrapper_random_points(gpu, orc):
    from fibsem_optflow_b200 import api, synth
    I0, I1 = synth.make_pair(128, 160, seed=3)
    I1[:10, :] = 0                     # a masked strip
    args = {"debug": True, "output_type": "random_points", "nscales": 4, "lambda": 0.15, "scale": 0.5}
    im = {"pId": "a", "qId": "b", "pGroupId": "1.0", "qGroupId": "2.0", "npoints": 30}
    fx, fy = api.solve_wrapper(I0, I1, im, args, roi_vec=((0, 10), (0, 10)))
    ou, ov, _, _ = orc.tvl1_calc(I0, I1, **{"lambda": 0.15, "nscales": 4})
    orc.mask_flow(I1, ou, ov)
    assert np.array_equal(fx, ou) and np.array_equal(fy, ov)
    want = orc.random_points(I0, I1, ou, ov, roi0=(0, 10), roi1=(0, 10), scale=0.5, npoints=30, seed=1)   # unseeded == srand(1)
    pm = im["point_matches"]
    assert pm["p"][0] == want[0].tolist() and pm["p"][1] == want[1].tolist()
    assert pm["q"][0] == want[2].tolist() and pm["q"][1] == want[3].tolist()
    assert pm["w"] == [1] * 30
    api.move_pm(im, args)
    assert args["point_matches"][0]["pId"] == "a"
    api.release_solvers()
