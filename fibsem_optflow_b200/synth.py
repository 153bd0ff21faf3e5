"""Synthetic FIB-SEM-like slice pairs.

Band-limited 8-bit noise with a known sub-pixel translation + shear.  Used by the
tests and by bench.py; plain NumPy/SciPy so that it runs identically here and on
the GPU box (there is no dataset access and the reference ships no images).
"""
import numpy as np

try:
    from scipy.ndimage import gaussian_filter as _gauss
except Exception:  # pragma: no cover
    _gauss = None


def _blur(a, sigma):
    if _gauss is not None:
        return _gauss(a, sigma, mode="wrap")
    # separable FIR fallback
    r = int(4 * sigma + 0.5)
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2).astype(np.float32)
    k /= k.sum()
    for ax in (0, 1):
        acc = np.zeros_like(a)
        for i, kv in enumerate(k):
            acc += kv * np.roll(a, i - r, axis=ax)
        a = acc
    return a


def _cr(t):
    """Catmull-Rom weights for fractional offset t in [0,1): taps at -1,0,1,2."""
    t = np.asarray(t, np.float32)
    t2, t3 = t * t, t * t * t
    return (-0.5 * t3 + t2 - 0.5 * t, 1.5 * t3 - 2.5 * t2 + 1.0,
            -1.5 * t3 + 2.0 * t2 + 0.5 * t, 0.5 * t3 - 0.5 * t2)


def texture(h, w, seed, sigma=2.0, coarse=0):
    """float32 canvas in [0,255], band-limited.  coarse = f > 1 adds a second band of structure f times
    larger (noise drawn on an f-times coarser grid, blurred there, repeated f times, then blurred with
    the fine band): what a pyramid needs to recover displacements of many pixels (configs[3])."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(h, w), dtype=np.uint8).astype(np.float32)
    if coarse and coarse > 1:
        f = int(coarse)
        hc, wc = (h + f - 1) // f, (w + f - 1) // f
        c = rng.integers(0, 256, size=(hc, wc), dtype=np.uint8).astype(np.float32)
        c = _blur(c, 1.0)
        c = np.repeat(np.repeat(c, f, axis=0), f, axis=1)[:h, :w]
        # equal contrast per band after the blur below: white noise loses ~ 1 / (2 sigma sqrt(pi)) of its std
        a += c * np.float32(1.0 / (2.0 * sigma * np.sqrt(np.pi)) * 2.0)
    a = _blur(a, sigma)
    lo, hi = float(a.min()), float(a.max())
    a -= lo
    a *= 255.0 / (hi - lo)
    return a


def shift_shear(canvas, dx, dy, shear):
    """canvas sampled at (x + dx + shear*y, y + dy), cubic, periodic border."""
    h, w = canvas.shape
    # vertical: constant shift
    iy = int(np.floor(dy))
    wy = _cr(dy - iy)
    v = np.zeros_like(canvas)
    for k in range(4):
        v += np.float32(wy[k]) * np.roll(canvas, -(iy + k - 1), axis=0)
    # horizontal: per-row shift
    s = dx + shear * np.arange(h, dtype=np.float64)
    ix = np.floor(s).astype(np.int64)
    wx = _cr((s - ix).astype(np.float32))
    out = np.zeros_like(canvas)
    cols = np.arange(w, dtype=np.int64)[None, :]
    rows = np.arange(h, dtype=np.int64)[:, None]
    for k in range(4):
        idx = (cols + ix[:, None] + (k - 1)) % w
        out += wx[k][:, None] * v[rows, idx]
    return out


def to_u8(a):
    return np.clip(np.rint(a), 0, 255).astype(np.uint8)


def make_pair(h, w, seed=7, dx=1.3, dy=-0.7, shear=0.002, sigma=2.0, margin=32, coarse=0):
    """Returns (I0, I1) uint8 with I1(x, y) = I0(x + dx + shear*y', y + dy) (y' in canvas
    rows), i.e. the true flow from I0 to I1 is u = -(dx + shear*y'), v = -dy."""
    H, W = h + 2 * margin, w + 2 * margin
    c = texture(H, W, seed, sigma, coarse)
    m = shift_shear(c, dx, dy, shear)
    sl = (slice(margin, margin + h), slice(margin, margin + w))
    return to_u8(c[sl]), to_u8(m[sl])


def true_flow(h, w, dx=1.3, dy=-0.7, shear=0.002, margin=32):
    ys = np.arange(h, dtype=np.float32)[:, None] + margin
    u = -(dx + shear * ys) * np.ones((1, w), np.float32)
    v = -dy * np.ones((h, w), np.float32)
    return u.astype(np.float32), v


def make_stack(n, h, w, seed=11, sigma=2.0, margin=32):
    """n+1 chained slices: slice k+1 = slice k's canvas shifted by a per-k sub-pixel offset
    (config 3/5: adjacent slices, pairs (k, k+1))."""
    H, W = h + 2 * margin, w + 2 * margin
    c = texture(H, W, seed, sigma)
    rng = np.random.default_rng(seed + 1)
    sl = (slice(margin, margin + h), slice(margin, margin + w))
    out = [to_u8(c[sl])]
    for _ in range(n):
        dx, dy = rng.uniform(-1.5, 1.5, size=2)
        c = shift_shear(c, float(dx), float(dy), 0.0)
        out.append(to_u8(c[sl]))
    return out
