"""ctypes front end of the C oracle (oracle/tvl1_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(fibsem_optflow_b200) never imports this module.  PARITY UNPINNED for the
composition -- see oracle/tvl1_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtvl1_oracle.so")


class OrcParams(C.Structure):
    _fields_ = [
        ("tau", C.c_double), ("lambda_", C.c_double), ("theta", C.c_double),
        ("epsilon", C.c_double), ("scale_step", C.c_double), ("gamma", C.c_double),
        ("nscales", C.c_int), ("warps", C.c_int), ("inner_iterations", C.c_int),
        ("outer_iterations", C.c_int), ("median_filtering", C.c_int),
        ("error_sum_mode", C.c_int), ("nthreads", C.c_int),
    ]


def build(force=False):
    src = os.path.join(_HERE, "tvl1_oracle.c")
    if force or not os.path.exists(_SO) or (
            os.path.exists(src) and os.path.getmtime(_SO) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    L.orc_default_params.argtypes = [C.POINTER(OrcParams)]
    L.orc_scaled_size.argtypes = [C.c_int, C.c_double]
    L.orc_scaled_size.restype = C.c_int
    L.orc_resize_linear.argtypes = [_f32p, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, C.c_double]
    L.orc_convert_u8.argtypes = [_u8p, C.c_long, C.c_int, C.c_int, _f32p]
    L.orc_centered_gradient.argtypes = [_f32p, C.c_int, C.c_int, _f32p, _f32p]
    L.orc_remap_cubic.argtypes = [_f32p, C.c_int, C.c_int, _f32p, _f32p, _f32p]
    L.orc_warp.argtypes = [_f32p] * 6 + [C.c_int, C.c_int, C.c_void_p] + [_f32p] * 4
    L.orc_iterate.argtypes = [_f32p] * 10 + [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                             C.c_int]
    L.orc_iterate.restype = C.c_double
    L.orc_iterate_gamma.argtypes = [_f32p] * 13 + [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                                   C.c_int]
    L.orc_iterate_gamma.restype = C.c_double
    L.orc_median5.argtypes = [_f32p, C.c_int, C.c_int, _f32p]
    L.orc_median25_selftest.restype = C.c_long
    L.orc_median3.argtypes = [_f32p, C.c_int, C.c_int, _f32p]
    L.orc_pyramid_sizes.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, _i32p, _i32p]
    L.orc_pyramid_sizes.restype = C.c_int
    L.orc_tvl1_calc.argtypes = [C.POINTER(OrcParams), _u8p, C.c_long, _u8p, C.c_long, C.c_int,
                                C.c_int, _f32p, _f32p, C.c_void_p]
    L.orc_tvl1_calc.restype = C.c_int
    L.orc_prescale_u8.argtypes = [_u8p, C.c_long, C.c_int, C.c_int, C.c_double, _u8p, C.c_long, C.c_int, C.c_int]
    L.orc_prescale_u8.restype = None
    L.orc_mask_flow.argtypes = [_u8p, C.c_long, C.c_int, C.c_int, _f32p, _f32p]
    L.orc_random_points.argtypes = [_u8p, C.c_long, _u8p, C.c_long, _f32p, _f32p, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                    C.c_long, _f64p, _f64p, _f64p, _f64p, _f64p, C.c_void_p]
    L.orc_random_points.restype = C.c_int
    L.orc_points_at.argtypes = [_f32p, _f32p, C.c_int, C.c_int, _i32p, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_float, _f64p, _f64p, _f64p, _f64p]
    _lib = L
    return L


def default_params(**kw):
    p = OrcParams()
    lib().orc_default_params(C.byref(p))
    for k, v in kw.items():
        if k == "lambda":
            k = "lambda_"
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def scaled_size(n, f):
    return lib().orc_scaled_size(int(n), float(f))


def resize_scale(src, f):
    src = _f32(src)
    h, w = src.shape
    dw, dh = scaled_size(w, f), scaled_size(h, f)
    dst = np.empty((dh, dw), np.float32)
    lib().orc_resize_linear(src, w, h, dst, dw, dh, float(f))
    return dst


def resize_to(src, dw, dh):
    src = _f32(src)
    h, w = src.shape
    dst = np.empty((dh, dw), np.float32)
    lib().orc_resize_linear(src, w, h, dst, dw, dh, 0.0)
    return dst


def centered_gradient(src):
    src = _f32(src)
    h, w = src.shape
    dx = np.empty_like(src)
    dy = np.empty_like(src)
    lib().orc_centered_gradient(src, w, h, dx, dy)
    return dx, dy


def remap_cubic(src, mapx, mapy):
    src, mapx, mapy = _f32(src), _f32(mapx), _f32(mapy)
    h, w = src.shape
    assert mapx.shape == src.shape and mapy.shape == src.shape
    dst = np.empty_like(src)
    lib().orc_remap_cubic(src, w, h, mapx, mapy, dst)
    return dst


def warp(I0, I1, I1x, I1y, u1, u2):
    I0, I1, I1x, I1y, u1, u2 = map(_f32, (I0, I1, I1x, I1y, u1, u2))
    h, w = I0.shape
    out = [np.empty_like(I0) for _ in range(5)]
    lib().orc_warp(I0, I1, I1x, I1y, u1, u2, w, h, out[0].ctypes.data, out[1], out[2], out[3],
                   out[4])
    return tuple(out)  # I1w, I1wx, I1wy, grad, rho_c


def iterate(I1wx, I1wy, grad, rho_c, u1, u2, p11, p12, p21, p22, l_t, theta, taut, mode=0):
    """One inner iteration IN PLACE on u1,u2,p11..p22 (float32 C-contiguous). Returns error."""
    h, w = u1.shape
    return lib().orc_iterate(_f32(I1wx), _f32(I1wy), _f32(grad), _f32(rho_c), u1, u2, p11, p12,
                             p21, p22, w, h, l_t, theta, taut, mode)


def iterate_gamma(I1wx, I1wy, grad, rho_c, u1, u2, u3, p11, p12, p21, p22, p31, p32, l_t, theta, taut, gamma, mode=0):
    """One inner iteration with the third channel (gamma != 0) IN PLACE on u1..u3, p11..p32.  Returns error."""
    h, w = u1.shape
    return lib().orc_iterate_gamma(_f32(I1wx), _f32(I1wy), _f32(grad), _f32(rho_c), u1, u2, u3, p11, p12,
                                   p21, p22, p31, p32, w, h, l_t, theta, taut, gamma, mode)


def median5(src):
    src = _f32(src)
    h, w = src.shape
    dst = np.empty_like(src)
    lib().orc_median5(src, w, h, dst)
    return dst


def median3(src):
    src = _f32(src)
    h, w = src.shape
    dst = np.empty_like(src)
    lib().orc_median3(src, w, h, dst)
    return dst


def pyramid_sizes(w, h, nscales, scale_step):
    ws = np.zeros(65, np.int32)
    hs = np.zeros(65, np.int32)
    n = lib().orc_pyramid_sizes(w, h, nscales, scale_step, ws, hs)
    return [(int(ws[i]), int(hs[i])) for i in range(n)]


def tvl1_calc(I0, I1, params=None, **kw):
    """Returns (u, v, iters[nscales][warps], levels_used)."""
    p = params if params is not None else default_params(**kw)
    I0 = np.ascontiguousarray(I0, np.uint8)
    I1 = np.ascontiguousarray(I1, np.uint8)
    h, w = I0.shape
    u = np.empty((h, w), np.float32)
    v = np.empty((h, w), np.float32)
    iters = np.full((p.nscales, p.warps), -1, np.int32)
    n = lib().orc_tvl1_calc(C.byref(p), I0, w, I1, w, w, h, u, v, iters.ctypes.data)
    if n < 0:
        raise RuntimeError("orc_tvl1_calc failed: %d" % n)
    return u, v, iters, n


def prescale_u8(src, scale):
    """8-bit cv::resize(src, Size(), scale, scale) of the reference's loader (src/optflow.cpp:111,124)."""
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dw, dh = scaled_size(w, scale), scaled_size(h, scale)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_prescale_u8(src, w, w, h, float(scale), dst, dw, dw, dh)
    return dst


def mask_flow(f1, u, v):
    f1 = np.ascontiguousarray(f1, np.uint8)
    h, w = f1.shape
    lib().orc_mask_flow(f1, w, w, h, u, v)


def random_points(f0, f1, u, v, roi0=(0, 0), roi1=(0, 0), scale=0.5, npoints=25, seed=-1):
    f0 = np.ascontiguousarray(f0, np.uint8)
    f1 = np.ascontiguousarray(f1, np.uint8)
    u, v = _f32(u), _f32(v)
    h, w = f0.shape
    n = max(npoints, 1)
    px, py, qx, qy, wg = (np.zeros(n, np.float64) for _ in range(5))
    pos = np.zeros((n, 2), np.int32)
    k = lib().orc_random_points(f0, w, f1, w, u, v, w, h, roi0[0], roi0[1], roi1[0], roi1[1],
                                scale, npoints, seed, px, py, qx, qy, wg, pos.ctypes.data)
    return px[:k], py[:k], qx[:k], qy[:k], wg[:k], pos[:k]


def points_at(u, v, positions, roi0=(0, 0), roi1=(0, 0), scale=0.5):
    u, v = _f32(u), _f32(v)
    pos = np.ascontiguousarray(positions, np.int32)
    n = pos.shape[0]
    px, py, qx, qy = (np.zeros(n, np.float64) for _ in range(4))
    lib().orc_points_at(u, v, u.shape[1], n, pos, roi0[0], roi0[1], roi1[0], roi1[1], scale,
                        px, py, qx, qy)
    return px, py, qx, qy
