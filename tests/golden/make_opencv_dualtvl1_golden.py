"""Pins the oracle's COMPOSITION to OpenCV's own class -- when an OpenCV build that has it exists.

    python tests/golden/make_opencv_dualtvl1_golden.py        # writes opencv_dualtvl1.npz

The reference solves through OpenCV 3.4.1's DualTVL1 (reference src/optflow.cpp:516-520), which is
not vendored; the cv2 of this image (opencv-python-headless 4.13) has no `optflow` contrib module, so
the class cannot run here and `oracle/` is pinned primitive by primitive only.
A maintainer with opencv-contrib-python (cv2.optflow.DualTVL1OpticalFlow_create, 4.x) or an OpenCV
3.4 build (cv2.createOptFlow_DualTVL1 / cv2.DualTVL1OpticalFlow_create) runs this script once: it
solves the seeded synthetic pairs below with OpenCV's class and stores inputs, parameters and flows.
tests/test_oracle_primitives.py::test_oracle_vs_opencv_dualtvl1 then compares the C oracle with the
stored flows (or with the live class if importable), and is skipped -- stating "parity unpinned" --
while neither exists.  Nothing here touches /root/reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

OUT = os.path.join(HERE, "opencv_dualtvl1.npz")

# (h, w, seed, dx, dy, shear, parameters) -- small enough to commit, varied enough to pin the loop
# structure (outer/inner counts, median on/off, stop test), the pyramid stop rule and the borders
CASES = [
    (96, 128, 7, 1.3, -0.7, 0.002, dict(nscales=5, warps=5)),
    (120, 90, 11, -0.6, 0.9, 0.0, dict(nscales=3, warps=2, lambda_=0.05)),
    (64, 200, 3, 2.2, 0.4, 0.004, dict(nscales=10, warps=3, medianFiltering=1)),
    (150, 150, 5, 0.3, 0.3, 0.0, dict(nscales=4, warps=4, innerIterations=7, outerIterations=3, epsilon=0.02)),
]
DEFAULTS = dict(tau=0.25, lambda_=0.15, theta=0.3, nscales=5, warps=5, epsilon=0.01, innerIterations=30,
                outerIterations=10, scaleStep=0.8, gamma=0.0, medianFiltering=5)


def opencv_factory():
    """Returns a callable creating OpenCV's CPU DualTVL1 object, or None if this cv2 has none."""
    try:
        import cv2
    except Exception:
        return None
    cands = []
    if hasattr(cv2, "optflow"):
        cands += [getattr(cv2.optflow, n, None) for n in ("DualTVL1OpticalFlow_create", "createOptFlow_DualTVL1")]
    cands += [getattr(cv2, n, None) for n in ("DualTVL1OpticalFlow_create", "createOptFlow_DualTVL1")]
    for c in cands:
        if c is not None:
            return c
    return None


def opencv_solve(factory, I0, I1, prm):
    import cv2
    cv2.setUseOptimized(False)      # the scalar code path: what a 3.4.1 SSE2 build computes (no FMA contraction)
    cv2.setNumThreads(1)            # the CPU class sums the error serially per stripe: one stripe = the literal order
    o = factory()
    for key, setter in (("tau", "setTau"), ("lambda_", "setLambda"), ("theta", "setTheta"), ("nscales", "setScalesNumber"),
                        ("warps", "setWarpingsNumber"), ("epsilon", "setEpsilon"), ("innerIterations", "setInnerIterations"),
                        ("outerIterations", "setOuterIterations"), ("scaleStep", "setScaleStep"), ("gamma", "setGamma"),
                        ("medianFiltering", "setMedianFiltering")):
        getattr(o, setter)(prm[key])
    o.setUseInitialFlow(False)
    flow = o.calc(I0, I1, None)
    return np.ascontiguousarray(flow[..., 0]), np.ascontiguousarray(flow[..., 1])


def cases():
    from fibsem_optflow_b200 import synth
    for k, (h, w, seed, dx, dy, shear, kw) in enumerate(CASES):
        prm = dict(DEFAULTS)
        prm.update(kw)
        I0, I1 = synth.make_pair(h, w, seed=seed, dx=dx, dy=dy, shear=shear)
        yield k, I0, I1, prm


def main():
    f = opencv_factory()
    if f is None:
        print("this cv2 has no DualTVL1 class (needs opencv-contrib-python or an OpenCV 3.4 build): nothing written; "
              "the oracle's composition stays unpinned")
        return 1
    import cv2
    out = {"n": len(CASES), "cv2_version": cv2.__version__}
    for k, I0, I1, prm in cases():
        u, v = opencv_solve(f, I0, I1, prm)
        out["I0_%d" % k], out["I1_%d" % k], out["u_%d" % k], out["v_%d" % k] = I0, I1, u, v
        out["prm_%d" % k] = np.array([prm[x] for x in sorted(prm)], np.float64)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, "from cv2", cv2.__version__)
    return 0


if __name__ == "__main__":
    sys.exit(main())
