"""NumPy + cv2 composition of OpenCV's CPU DualTVL1.

TEST INFRASTRUCTURE ONLY (same rules as oracle/tvl1_oracle.h).  PARITY UNPINNED
for the composition: OpenCV's own DualTVL1 class is not installed anywhere here.

This second oracle exists to pin the C restatement: it calls the REAL
cv2.resize / cv2.remap(INTER_CUBIC) / cv2.medianBlur(5) -- the three library
primitives cv::DualTVL1OpticalFlow composes -- and does the elementwise steps
in NumPy fp32 (NumPy never contracts to FMA).  cv2's dispatched AVX2/FMA
kernels are switched off (cv2.setUseOptimized(False)) so that resize runs the
plain mul+add code path that OpenCV 3.4.1's SSE build has; with that the C
oracle matches this module bit for bit (tests/test_oracle_vs_cv2.py).
"""
import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

F32 = np.float32
FLT_EPSILON = np.float32(1.1920929e-07)
FLT_MAX = np.float32(3.4028235e38)


def _need_cv2():
    if cv2 is None:
        raise RuntimeError("cv2 is required for oracle/tvl1_ref.py")
    cv2.setUseOptimized(False)


def centered_gradient(I):
    # A.3 (index clamping)
    P = np.pad(I, 1, mode="edge")
    dx = F32(0.5) * (P[1:-1, 2:] - P[1:-1, :-2])
    dy = F32(0.5) * (P[2:, 1:-1] - P[:-2, 1:-1])
    return dx.astype(F32), dy.astype(F32)


def warp(I0, I1, I1x, I1y, u1, u2):
    # A.4
    _need_cv2()
    h, w = I0.shape
    xs = np.arange(w, dtype=F32)[None, :]
    ys = np.arange(h, dtype=F32)[:, None]
    m1 = (xs + u1).astype(F32)
    m2 = (ys + u2).astype(F32)
    I1w = cv2.remap(I1, m1, m2, cv2.INTER_CUBIC)
    I1wx = cv2.remap(I1x, m1, m2, cv2.INTER_CUBIC)
    I1wy = cv2.remap(I1y, m1, m2, cv2.INTER_CUBIC)
    grad = I1wx * I1wx + I1wy * I1wy
    rho_c = ((I1w - I1wx * u1) - I1wy * u2) - I0
    return I1w, I1wx, I1wy, grad, rho_c


def divergence(a, b):
    # A.5 step 3, including the association order on the first row/column
    d = np.empty_like(a)
    d[1:, 1:] = (a[1:, 1:] - a[1:, :-1]) + (b[1:, 1:] - b[:-1, 1:])
    d[0, 1:] = (a[0, 1:] - a[0, :-1]) + b[0, 1:]
    d[1:, 0] = (a[1:, 0] + b[1:, 0]) - b[:-1, 0]
    d[0, 0] = a[0, 0] + b[0, 0]
    return d


def forward_gradient(u):
    dx = np.zeros_like(u)
    dy = np.zeros_like(u)
    dx[:, :-1] = u[:, 1:] - u[:, :-1]
    dy[:-1, :] = u[1:, :] - u[:-1, :]
    return dx, dy


def hypot_f(a, b):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    return np.sqrt(a * a + b * b).astype(F32)


def iterate(I1wx, I1wy, grad, rho_c, u1, u2, p11, p12, p21, p22, l_t, theta, taut):
    """One inner iteration (A.5). Returns new (u1,u2,p11,p12,p21,p22,error)."""
    l_t, theta, taut = F32(l_t), F32(theta), F32(taut)
    rho = rho_c + (I1wx * u1 + I1wy * u2)
    lg = l_t * grad
    c1 = rho < -lg
    c2 = (~c1) & (rho > lg)
    c3 = (~c1) & (~c2) & (grad > FLT_EPSILON)
    with np.errstate(divide="ignore", invalid="ignore"):
        fi = -rho / grad
    d1 = np.zeros_like(u1)
    d2 = np.zeros_like(u2)
    d1 = np.where(c1, l_t * I1wx, d1)
    d2 = np.where(c1, l_t * I1wy, d2)
    d1 = np.where(c2, -l_t * I1wx, d1)
    d2 = np.where(c2, -l_t * I1wy, d2)
    with np.errstate(invalid="ignore", over="ignore"):
        d1 = np.where(c3, fi * I1wx, d1).astype(F32)
        d2 = np.where(c3, fi * I1wy, d2).astype(F32)
    v1 = u1 + d1
    v2 = u2 + d2
    div1 = divergence(p11, p12)
    div2 = divergence(p21, p22)
    nu1 = v1 + theta * div1
    nu2 = v2 + theta * div2
    term = (nu1 - u1) * (nu1 - u1) + (nu2 - u2) * (nu2 - u2)
    error = float(np.sum(term.astype(np.float64)))
    u1x, u1y = forward_gradient(nu1)
    u2x, u2y = forward_gradient(nu2)
    ng1 = F32(1.0) + taut * hypot_f(u1x, u1y)
    ng2 = F32(1.0) + taut * hypot_f(u2x, u2y)
    p11 = (p11 + taut * u1x) / ng1
    p12 = (p12 + taut * u1y) / ng1
    p21 = (p21 + taut * u2x) / ng2
    p22 = (p22 + taut * u2y) / ng2
    return nu1, nu2, p11, p12, p21, p22, error


def iterate_gamma(I1wx, I1wy, grad, rho_c, u1, u2, u3, p11, p12, p21, p22, p31, p32, l_t, theta, taut, gamma):
    """One inner iteration with the third channel (A.5, gamma != 0).
    Returns new (u1,u2,u3,p11,p12,p21,p22,p31,p32,error)."""
    l_t, theta, taut, gamma = F32(l_t), F32(theta), F32(taut), F32(gamma)
    rho = (rho_c + (I1wx * u1 + I1wy * u2)) + gamma * u3
    lg = l_t * grad
    c1 = rho < -lg
    c2 = (~c1) & (rho > lg)
    c3 = (~c1) & (~c2) & (grad > FLT_EPSILON)
    with np.errstate(divide="ignore", invalid="ignore"):
        fi = (-rho / grad).astype(F32)
    zero = np.zeros_like(u1)
    with np.errstate(invalid="ignore", over="ignore"):
        d = []
        for t in (I1wx, I1wy, np.full_like(u1, gamma)):
            x = np.where(c1, l_t * t, zero)
            x = np.where(c2, -l_t * t, x)
            x = np.where(c3, fi * t, x).astype(F32)
            d.append(x)
    v = [u1 + d[0], u2 + d[1], u3 + d[2]]
    div = [divergence(p11, p12), divergence(p21, p22), divergence(p31, p32)]
    nu = [v[k] + theta * div[k] for k in range(3)]
    e = [nu[0] - u1, nu[1] - u2, nu[2] - u3]
    term = (e[0] * e[0] + e[1] * e[1]) + e[2] * e[2]
    error = float(np.sum(term.astype(np.float64)))
    out_p = []
    for k, (pa, pb) in enumerate(((p11, p12), (p21, p22), (p31, p32))):
        ux, uy = forward_gradient(nu[k])
        ng = F32(1.0) + taut * hypot_f(ux, uy)
        out_p += [(pa + taut * ux) / ng, (pb + taut * uy) / ng]
    return (nu[0], nu[1], nu[2], *out_p, error)


def tvl1_calc(I0u8, I1u8, tau=0.25, lambda_=0.15, theta=0.3, nscales=5, warps=5, epsilon=0.01,
              inner_iterations=30, outer_iterations=10, scale_step=0.8, median_filtering=5, gamma=0.0):
    """Returns (u, v, iters[levels][warps])."""
    _need_cv2()
    I0s = [I0u8.astype(F32)]
    I1s = [I1u8.astype(F32)]
    for s in range(1, nscales):
        a = cv2.resize(I0s[s - 1], None, fx=scale_step, fy=scale_step,
                       interpolation=cv2.INTER_LINEAR)
        b = cv2.resize(I1s[s - 1], None, fx=scale_step, fy=scale_step,
                       interpolation=cv2.INTER_LINEAR)
        if a.shape[1] < 16 or a.shape[0] < 16:
            nscales = s
            break
        I0s.append(a)
        I1s.append(b)
    l_t = F32(lambda_ * theta)
    taut = F32(tau / theta)
    th = F32(theta)
    u1 = np.zeros_like(I0s[-1])
    u2 = np.zeros_like(I0s[-1])
    u3 = np.zeros_like(I0s[-1])          # gamma != 0 only
    iters = np.full((nscales, warps), -1, np.int32)
    for s in range(nscales - 1, -1, -1):
        I0, I1 = I0s[s], I1s[s]
        h, w = I0.shape
        scaled_eps = F32(epsilon * epsilon * (w * h))
        I1x, I1y = centered_gradient(I1)
        p11 = np.zeros_like(I0); p12 = np.zeros_like(I0)
        p21 = np.zeros_like(I0); p22 = np.zeros_like(I0)
        p31 = np.zeros_like(I0); p32 = np.zeros_like(I0)
        for wi in range(warps):
            _, I1wx, I1wy, grad, rho_c = warp(I0, I1, I1x, I1y, u1, u2)
            error = FLT_MAX
            count = 0
            no = 0
            while error > scaled_eps and no < outer_iterations:
                if median_filtering > 1:
                    u1 = cv2.medianBlur(u1, median_filtering)
                    u2 = cv2.medianBlur(u2, median_filtering)
                ni = 0
                while error > scaled_eps and ni < inner_iterations:
                    if gamma != 0.0:
                        u1, u2, u3, p11, p12, p21, p22, p31, p32, e = iterate_gamma(
                            I1wx, I1wy, grad, rho_c, u1, u2, u3, p11, p12, p21, p22, p31, p32, l_t, th, taut, gamma)
                    else:
                        u1, u2, p11, p12, p21, p22, e = iterate(I1wx, I1wy, grad, rho_c, u1, u2,
                                                                p11, p12, p21, p22, l_t, th, taut)
                    error = F32(e)
                    count += 1
                    ni += 1
                no += 1
            iters[s, wi] = count
        if s == 0:
            break
        ph, pw = I0s[s - 1].shape
        up = F32(1 / scale_step)
        u1 = cv2.resize(u1, (pw, ph), interpolation=cv2.INTER_LINEAR) * up
        u2 = cv2.resize(u2, (pw, ph), interpolation=cv2.INTER_LINEAR) * up
        if gamma != 0.0:
            u3 = cv2.resize(u3, (pw, ph), interpolation=cv2.INTER_LINEAR)   # zoomed, not scaled
    return u1, u2, iters
