"""Generates the golden fixtures in this directory.  Run in the build container:

    python tests/golden/make_golden.py

Sources of truth (none of them is this repository's own code):
  * cv2 4.13 (opencv-python-headless) with dispatched SIMD off -> remap / resize / medianBlur
    outputs, and -- through oracle/tvl1_ref.py, which only composes those cv2 calls with
    NumPy elementwise steps -- a whole-pair flow field;
  * glibc (ctypes -> libc.so.6): rand() streams and the libstdc++ random_shuffle recurrence;
  * SURVEY.md C6 known answers (kept literally, and re-derived here).
The reference repository itself ships no fixtures.
"""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402

from fibsem_optflow_b200 import synth  # noqa: E402
from oracle import tvl1_ref  # noqa: E402

cv2.setUseOptimized(False)
rng = np.random.default_rng(20261018)


def primitives():
    h, w = 40, 56
    src = (rng.random((h, w)) * 255).astype(np.float32)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    mx = (xx + rng.standard_normal((h, w)) * 2.5).astype(np.float32)
    my = (yy + rng.standard_normal((h, w)) * 2.5).astype(np.float32)
    # a block of far-outside and edge-straddling coordinates
    mx[:4] += 70
    my[-4:] -= 55
    mx[10:14, :6] -= 3.3
    remap = cv2.remap(src, mx, my, cv2.INTER_CUBIC)
    down = cv2.resize(src, None, fx=0.8, fy=0.8, interpolation=cv2.INTER_LINEAR)
    up = cv2.resize(down, (w, h), interpolation=cv2.INTER_LINEAR)
    med_in = (rng.standard_normal((h, w)) * 3).astype(np.float32)
    med = cv2.medianBlur(med_in, 5)
    np.savez_compressed(os.path.join(HERE, "primitives.npz"), src=src, mx=mx, my=my, remap=remap,
                        down=down, up=up, med_in=med_in, med=med)
    sizes = {str(n): int(cv2.resize(np.zeros((n, n), np.float32), None, fx=0.8, fy=0.8).shape[0])
             for n in (2048, 1638, 1310, 1048, 838, 4096, 6144, 8192, 16384, 17, 19, 20, 25, 30)}
    return sizes


def pair():
    I0, I1 = synth.make_pair(96, 128, seed=7)
    u, v, iters = tvl1_ref.tvl1_calc(I0, I1)            # OpenCV CPU-class defaults
    u2, v2, iters2 = tvl1_ref.tvl1_calc(I0, I1, lambda_=0.05, nscales=10)   # reference wrapper defaults
    np.savez_compressed(os.path.join(HERE, "pair_96x128.npz"), I0=I0, I1=I1, u=u, v=v, iters=iters,
                        u_ref=u2, v_ref=v2, iters_ref=iters2)


def glibc():
    code = r'''
import ctypes, json, sys
libc = ctypes.CDLL("libc.so.6")
seed = int(sys.argv[1])
if seed >= 0:
    libc.srand(seed)
first = [libc.rand() for _ in range(8)]
rest = [libc.rand() for _ in range(100000 - 8)]
v = list(range(12))
# second process state is irrelevant: the shuffle below is re-derived from a fresh stream
print(json.dumps({"first": first, "at_99999": rest[-1]}))
'''
    out = {}
    for seed in (-1, 1, 0, 12345, 1539000000):
        r = json.loads(subprocess.check_output([sys.executable, "-c", code, str(seed)]))
        out[str(seed)] = r
    # libstdc++ random_shuffle of 0..11 with an unseeded rand(): SURVEY.md C6
    shuffle_code = r'''
import ctypes, json
libc = ctypes.CDLL("libc.so.6")
v = list(range(12))
for i in range(1, 12):
    j = libc.rand() % (i + 1)
    v[i], v[j] = v[j], v[i]
print(json.dumps(v))
'''
    out["shuffle12_unseeded"] = json.loads(subprocess.check_output([sys.executable, "-c", shuffle_code]))
    assert out["shuffle12_unseeded"] == [4, 10, 11, 8, 0, 5, 2, 1, 6, 9, 3, 7], out["shuffle12_unseeded"]
    assert out["-1"]["first"][:3] == [1804289383, 846930886, 1681692777]
    return out


def primitives2():
    """Round 2 additions (own generator and file, so that the round-1 fixtures stay byte-identical):
    the 3x3 median, the scale-0.5 pyramid step (cv::resize's INTER_AREA fast path) on odd and even sizes,
    and whole-pair flows of the cv2 composition with medianFiltering 3 / scaleStep 0.5."""
    r2 = np.random.default_rng(20261019)
    out = {}
    med_in = (r2.standard_normal((37, 53)) * 3).astype(np.float32)
    med_in[r2.random(med_in.shape) < 0.2] = 0
    out["med3_in"] = med_in
    out["med3"] = cv2.medianBlur(med_in, 3)
    for k, (h, w) in enumerate([(37, 53), (31, 17), (40, 56), (7, 10), (3, 7)]):
        a = (r2.random((h, w)) * 255).astype(np.float32)
        out["half_in_%d" % k] = a
        out["half_%d" % k] = cv2.resize(a, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_LINEAR)
    I0, I1 = synth.make_pair(90, 120, seed=13)
    out["I0"], out["I1"] = I0, I1
    out["u_med3"], out["v_med3"], out["it_med3"] = tvl1_ref.tvl1_calc(I0, I1, median_filtering=3, nscales=4)
    out["u_half"], out["v_half"], out["it_half"] = tvl1_ref.tvl1_calc(I0, I1, scale_step=0.5, nscales=4)
    np.savez_compressed(os.path.join(HERE, "primitives2.npz"), **out)


def main():
    if "--round2" in sys.argv:
        primitives2()
        print("primitives2.npz written to", HERE)
        return
    sizes = primitives()
    pair()
    primitives2()
    known = {
        "resize_sizes_0.8": sizes,
        "glibc_rand": glibc(),
        # fp32 (100 + 7 + 1.3f) * 2.0f widened to double
        "match_arith": {"pos": 100, "roi": 7, "flow": 1.3, "inv_scale": 2.0,
                        "q": float((np.float32(107) + np.float32(1.3)) * np.float32(2.0))},
    }
    assert known["match_arith"]["q"] == 216.60000610351562
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(known, f, indent=1, sort_keys=True)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
