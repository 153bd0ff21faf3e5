"""Feature pre-alignment of the reference (src/features.cpp:46-167, call site src/optflow.cpp:366-377),
restated with the CPU classes of the installed cv2 -- the same OpenCV algorithms the reference runs through
their cv::cuda variants (ORB, brute-force Hamming 2-NN, ratio test, RANSAC homography, warpAffine).

TEST INFRASTRUCTURE ONLY (tests/, never the product).  cv::cuda::ORB and cv::findHomography are not
bit-reproducible across builds (keypoint ties, RANSAC's RNG), and the product's descriptor is its own, so
tests compare TRANSFORMS (displacement of the frame corners under the two affines), not keypoints.
"""
import numpy as np

ORB_TYPE, SURF_TYPE = 1, 2          # src/features.h:13-14

ORB_DEFAULTS = dict(nfeatures=5000, scaleFactor=1.2, nlevels=8, edgeThreshold=31, firstLevel=0, WTA_K=2,
                    patchSize=31, fastThreshold=20)          # orb_defaults, src/features.cpp:19-32


def find_alignment(frame_q, frame_t, ratio=0.8, ransac=5.0, homo=8, **orb):
    """find_alignment(frame1, frame0, ...) as solve_rois calls it (src/optflow.cpp:373): `frame_q` is the
    frame to be moved (the pair's frame1), `frame_t` the fixed one.  Returns (affine 2x3 float32 that maps
    frame_q coordinates to frame_t coordinates, n_matches, n_good)."""
    import cv2
    p = dict(ORB_DEFAULTS)
    p.update(orb)
    det = cv2.ORB_create(nfeatures=p["nfeatures"], scaleFactor=p["scaleFactor"], nlevels=p["nlevels"],
                         edgeThreshold=p["edgeThreshold"], firstLevel=p["firstLevel"], WTA_K=p["WTA_K"],
                         scoreType=cv2.ORB_HARRIS_SCORE, patchSize=p["patchSize"], fastThreshold=p["fastThreshold"])
    k0, d0 = det.detectAndCompute(frame_q, None)
    k1, d1 = det.detectAndCompute(frame_t, None)
    ident = np.array([[1, 0, 0], [0, 1, 0]], np.float32)
    if d0 is None or d1 is None or len(k1) < 2:
        return ident, 0, 0
    matches = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(d0, d1, k=2)
    good = []
    for i in range(min(d1.shape[0] - 1, len(matches))):          # the reference's loop bound, src/features.cpp:105
        m = matches[i]
        if 0 < len(m) <= 2 and len(m) == 2 and m[0].distance < ratio * m[1].distance:
            good.append(m[0])
    good.sort(key=lambda m: m.distance)
    if len(good) <= 10:
        return ident, len(matches), len(good)
    p0 = np.float32([k0[m.queryIdx].pt for m in good])
    p1 = np.float32([k1[m.trainIdx].pt for m in good])
    H, _ = cv2.findHomography(p0, p1, homo, ransac)
    if H is None or abs(1 - H[0, 0]) > 0.20 or abs(1 - H[1, 1]) > 0.20:
        return ident, len(matches), len(good)
    return H[:2, :].astype(np.float32), len(matches), len(good)


def warp_affine(img, affine, size):
    """cv::cuda::warpAffine(src, dst, affine, size, INTER_LINEAR, BORDER_CONSTANT, 0) (src/optflow.cpp:374,
    431-432) through the CPU function: dst(x) = src(A^-1 x)."""
    import cv2
    return cv2.warpAffine(img, np.asarray(affine, np.float64), (int(size[0]), int(size[1])), flags=cv2.INTER_LINEAR,
                          borderMode=cv2.BORDER_CONSTANT, borderValue=0)


def corner_error(a, b, w, h):
    """largest distance between the images of the frame corners (and centre) under two affines"""
    pts = np.array([[0, 0, 1], [w, 0, 1], [0, h, 1], [w, h, 1], [w / 2, h / 2, 1]], np.float64).T
    da = np.asarray(a, np.float64) @ pts - np.asarray(b, np.float64) @ pts
    return float(np.sqrt((da ** 2).sum(0)).max())
