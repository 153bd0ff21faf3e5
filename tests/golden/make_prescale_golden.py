"""Golden vectors for the 8-bit prescale (reference src/optflow.cpp:111,124: cv::resize on the decoded
frame).  Made with the installed cv2 (SIMD dispatch on, the way a user's OpenCV runs), not with this
repository's code.  Run from the repo root: python tests/golden/make_prescale_golden.py"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(20261018)
    out = {}
    k = 0
    for (h, w) in [(97, 131), (64, 64), (33, 100), (120, 7), (50, 51)]:
        src = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
        src[: h // 4] = 0                      # a masked band, like padded FIB-SEM frames
        for sc in (0.5, 0.25, 0.3, 0.75, 0.8, 1.5):
            scf = float(np.float32(sc))        # the reference holds `scale` in a float
            dst = cv2.resize(src, None, fx=scf, fy=scf)
            out["src_%d" % k] = src
            out["scale_%d" % k] = np.float64(scf)
            out["dst_%d" % k] = dst
            k += 1
    out["n"] = np.int64(k)
    np.savez_compressed(os.path.join(HERE, "prescale.npz"), **out)
    print("prescale.npz:", k, "cases, cv2", cv2.__version__)


if __name__ == "__main__":
    main()
