"""The job driver (fibsem_optflow_b200/host/optflow_b200, N1): same job JSON as the reference CLI,
outputs checked against the oracle -- float TIFF planes for "flow"/"map", match records for
"random_points" (debug mode: one unseeded rand() stream over all ROIs and pairs, in the
reference's alphabetical ROI order)."""
import gzip
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from fibsem_optflow_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "fibsem_optflow_b200", "host", "optflow_b200")


def build_cli():
    subprocess.check_call(["make", "-C", os.path.dirname(EXE), "-s"])
    return EXE


def write_png(path, a):
    import cv2
    assert cv2.imwrite(path, a)


def write_pgm(path, a):
    with open(path, "wb") as f:
        f.write(b"P5\n# test\n%d %d\n255\n" % (a.shape[1], a.shape[0]))
        f.write(a.tobytes())


def read_tiff_f32(path):
    from PIL import Image
    return np.array(Image.open(path), dtype=np.float32)


def test_flow_and_map_tiffs(gpu, orc, tmp_path):
    exe = build_cli()
    sl = synth.make_stack(2, 256, 320, seed=21)         # 3 slices, prescaled by 0.5 in the driver
    sl[1][:16, :] = 0
    names = []
    for k, a in enumerate(sl):
        p = str(tmp_path / ("s%d.%s" % (k, "png" if k != 1 else "pgm")))
        (write_png if k != 1 else write_pgm)(p, a)
        names.append(p)
    job = """{
      /* a comment, as in docs/example.json */
      "debug": true, "style": 1, "features": false,
      "images": [ {"p": "%s", "q": "%s", "output_name": "a~b", "output_type": "flow"},
                  {"p": "%s", "q": "%s", "output_name": "b~c", "output_type": "map", "nscales": 3}, ],
      "rois": {"top": 40, "bottom": 48},
      "lambda": 0.15, "nscales": 4, "scale": 0.5, // trailing comma above is tolerated too
      "output_dir": "%s"
    }""" % (names[0], names[1], names[1], names[2], str(tmp_path))
    jf = str(tmp_path / "job.json.gz")
    with gzip.open(jf, "wt") as f:
        f.write(job)
    subprocess.check_call([exe, jf])
    half = [((a[0::2, 0::2].astype(np.int32) + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2).astype(np.uint8)
            for a in sl]
    H, W = half[0].shape
    for (name, k, nsc, is_map) in (("a~b", 0, 4, False), ("b~c", 1, 3, True)):
        for key, (y0, hh) in (("top", (0, 40)), ("bottom", (H - 48, 48))):
            f0, f1 = half[k][y0:y0 + hh], half[k + 1][y0:y0 + hh]
            ou, ov, _, _ = orc.tvl1_calc(f0, f1, **{"lambda": 0.15, "nscales": nsc})
            if is_map:
                ou = ou + np.arange(W, dtype=np.float32)[None, :]
                ov = ov + np.arange(hh, dtype=np.float32)[:, None]
            ou = np.where(f1 <= 1, np.float32(0), ou)
            ov = np.where(f1 <= 1, np.float32(0), ov)
            base = str(tmp_path / ("%s_0.50_%s" % (name, key)))
            assert np.array_equal(read_tiff_f32(base + "_x.tiff"), ou), (name, key)
            assert np.array_equal(read_tiff_f32(base + "_y.tiff"), ov), (name, key)


def test_random_points_job(gpu, orc, tmp_path):
    exe = build_cli()
    sl = synth.make_stack(2, 96, 128, seed=4)
    names = []
    for k, a in enumerate(sl):
        p = str(tmp_path / ("t%d.png" % k))
        write_png(p, a)
        names.append(p)
    # top-level keys exactly as support_scripts/gen_cross_file_list.py:75-99 writes them (N3); the
    # feature-matching knobs are carried along and ignored when "features" is absent
    job = {"style": 1, "debug": True, "homo": 4, "ratio": 0.7, "ransac": 5, "hessianThreshold": 1600,
           "host": "render.example.org", "port": 8080, "matchCollection": "test_v1", "owner": "flyem",
           "output_type": "random_points", "scale": 1.0, "lambda": 0.15, "nscales": 3,
           "npoints": 7, "output_dir": str(tmp_path), "rois": {"top": 32, "bottom": 40},
           "images": [{"p": names[0], "q": names[1], "pId": "t0", "qId": "t1", "pGroupId": "1.0", "qGroupId": "2.0"},
                      {"p": names[1], "q": names[2], "pId": "t1", "qId": "t2", "pGroupId": "2.0", "qGroupId": "3.0",
                       "rois": {"custom": [8, 16, 100, 60]}}]}
    jf = str(tmp_path / "job.json")
    json.dump(job, open(jf, "w"))
    subprocess.check_call([exe, jf], stdout=subprocess.DEVNULL)
    got = json.load(open(str(tmp_path / "point_matches_000.json")))
    assert [g["pId"] for g in got] == ["t0", "t1"] and got[1]["qGroupId"] == "3.0"
    # the record move_pm builds (src/optflow.cpp:574-593), which upload_points PUTs to the Render service
    assert all(set(g) == {"pGroupId", "pId", "qGroupId", "qId", "matches"} for g in got)
    assert all(set(g["matches"]) == {"p", "q", "w"} for g in got)
    # the same job through the oracle, in a fresh process (unseeded rand() stream), ROI keys in
    # jsoncpp's alphabetical order: bottom, then top
    np.savez(str(tmp_path / "in.npz"), sl=np.stack(sl))
    code = (
        "import sys, json; sys.path.insert(0, %r)\n"
        "import numpy as np\n"
        "from oracle import oracle as O\n"
        "sl = np.load(%r)['sl']; H, W = sl[0].shape; out = []\n"
        "jobs = [(0, [(0, H - 40, W, 40), (0, 0, W, 32)]), (1, [(8, 16, 100, 60)])]\n"
        "for k, rois in jobs:\n"
        "    rec = [[], [], [], [], []]\n"
        "    for (x, y, w, h) in rois:\n"
        "        f0 = np.ascontiguousarray(sl[k][y:y+h, x:x+w]); f1 = np.ascontiguousarray(sl[k+1][y:y+h, x:x+w])\n"
        "        u, v, _, _ = O.tvl1_calc(f0, f1, **{'lambda': 0.15, 'nscales': 3})\n"
        "        O.mask_flow(f1, u, v)\n"
        "        r = O.random_points(f0, f1, u, v, roi0=(x, y), roi1=(x, y), scale=1.0, npoints=7, seed=-1)\n"
        "        for j in range(5): rec[j] += r[j].tolist()\n"
        "    out.append(rec)\n"
        "print(json.dumps(out))\n" % (ROOT, str(tmp_path / "in.npz")))
    want = json.loads(subprocess.check_output([sys.executable, "-c", code]).decode())
    for k in range(2):
        m = got[k]["matches"]
        assert m["p"][0] == want[k][0] and m["p"][1] == want[k][1]
        assert m["q"][0] == want[k][2] and m["q"][1] == want[k][3]
        assert m["w"] == [int(x) for x in want[k][4]]


def test_cli_rejects(gpu, tmp_path):
    exe = build_cli()
    jf = str(tmp_path / "bad.json")
    open(jf, "w").write('{"images": [], "features": 1, "style": 2}')
    assert subprocess.call([exe, jf], stderr=subprocess.DEVNULL) != 0
    open(jf, "w").write('{"images": [ {"p": "x" "q": "y"} ]}')     # missing comma, like docs/example.json:72
    assert subprocess.call([exe, jf], stderr=subprocess.DEVNULL) != 0


def test_general_scale_and_prefetch(gpu, orc, tmp_path):
    """N2: any `scale` (prescaled on the device exactly like the loader's 8-bit cv::resize), frames of
    the next pair decoded on a host thread; three unrelated pairs so that nothing comes from the cache."""
    exe = build_cli()
    sl = synth.make_stack(5, 200, 260, seed=33)
    names = []
    for k, a in enumerate(sl):
        p = str(tmp_path / ("g%d.png" % k))
        write_png(p, a)
        names.append(p)
    pairs = [(0, 1), (2, 3), (4, 5)]
    # a whole-frame "custom" roi: without any roi the reference pre-aligns by features (src/optflow.cpp:366)
    job = {"debug": True, "output_type": "flow", "scale": 0.3, "lambda": 0.15, "nscales": 3, "output_dir": str(tmp_path),
           "rois": {"custom": [0, 0, 78, 60]},
           "images": [{"p": names[a], "q": names[b], "output_name": "p%d" % a} for a, b in pairs]}
    jf = str(tmp_path / "job.json")
    json.dump(job, open(jf, "w"))
    subprocess.check_call([exe, jf], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    scf = float(np.float32(0.3))
    for a, b in pairs:
        f0, f1 = orc.prescale_u8(sl[a], scf), orc.prescale_u8(sl[b], scf)
        ou, ov, _, _ = orc.tvl1_calc(f0, f1, **{"lambda": 0.15, "nscales": 3})
        ou = np.where(f1 <= 1, np.float32(0), ou)
        ov = np.where(f1 <= 1, np.float32(0), ov)
        base = str(tmp_path / ("p%d_0.30" % a))
        assert np.array_equal(read_tiff_f32(base + "_x.tiff"), ou)
        assert np.array_equal(read_tiff_f32(base + "_y.tiff"), ov)


def test_sharded_random_points_cover_every_pair(gpu, tmp_path):
    """--shard R/W (one process per GPU): every rank writes its own batch files, so the union of the
    outputs holds every pair exactly once; a pair with a bad roi is logged and skipped, the job goes
    on and its pending matches are still flushed; "matches_file" collects every batch of a process."""
    exe = build_cli()
    sl = synth.make_stack(5, 64, 96, seed=9)
    names = []
    for k, a in enumerate(sl):
        p = str(tmp_path / ("h%d.png" % k))
        write_png(p, a)
        names.append(p)
    images = [{"p": names[k], "q": names[k + 1], "pId": "h%d" % k, "qId": "h%d" % (k + 1),
               "pGroupId": "%d.0" % k, "qGroupId": "%d.0" % (k + 1)} for k in range(5)]
    images[3]["rois"] = {"custom": [0, 0, 4000, 10]}            # outside the frame: skipped, not fatal
    job = {"debug": True, "output_type": "random_points", "scale": 1.0, "lambda": 0.15, "nscales": 2, "npoints": 3,
           "batch_size": 0, "output_dir": str(tmp_path), "images": images}
    jf = str(tmp_path / "job.json")
    json.dump(job, open(jf, "w"))
    for r in range(2):
        subprocess.check_call([exe, "--shard", "%d/2" % r, jf], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    seen = []
    files = sorted(f for f in os.listdir(str(tmp_path)) if f.startswith("point_matches_"))
    assert files and all("_r0of2_" in f or "_r1of2_" in f for f in files)
    for f in files:
        seen += [g["pId"] for g in json.load(open(str(tmp_path / f)))]
    assert sorted(seen) == ["h0", "h1", "h2", "h4"]             # each once; h3 skipped
    # matches_file: all batches of the process in one array
    job["matches_file"] = str(tmp_path / "all.json")
    json.dump(job, open(jf, "w"))
    subprocess.check_call([exe, jf], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    assert [g["pId"] for g in json.load(open(job["matches_file"]))] == ["h0", "h1", "h2", "h4"]


def test_run_job_launcher(gpu, tmp_path):
    """python -m fibsem_optflow_b200.run_job: one driver process per rank, side by side (here both ranks on
    device 0); the union of the ranks' outputs is the job's output"""
    from fibsem_optflow_b200 import run_job
    build_cli()
    sl = synth.make_stack(4, 64, 96, seed=19)
    names = []
    for k, a in enumerate(sl):
        p = str(tmp_path / ("g%d.png" % k))
        write_png(p, a)
        names.append(p)
    images = [{"p": names[k], "q": names[k + 1], "pId": "g%d" % k, "qId": "g%d" % (k + 1),
               "pGroupId": "%d.0" % k, "qGroupId": "%d.0" % (k + 1), "output_name": "g%d" % k} for k in range(4)]
    job = {"debug": True, "output_type": "random_points", "scale": 1.0, "lambda": 0.15, "nscales": 2, "npoints": 3,
           "rois": {"custom": [0, 0, 96, 64]}, "output_dir": str(tmp_path), "images": images}
    jf = str(tmp_path / "job.json")
    json.dump(job, open(jf, "w"))
    assert run_job.run(jf, 2, devices=[0, 0]) == [0, 0]
    seen = []
    for f in sorted(f for f in os.listdir(str(tmp_path)) if f.startswith("point_matches_")):
        assert "_r0of2_" in f or "_r1of2_" in f
        seen += [g["pId"] for g in json.load(open(str(tmp_path / f)))]
    assert sorted(seen) == ["g0", "g1", "g2", "g3"]


def test_job_without_roi_is_prealigned(gpu, tmp_path):
    """N4: a pair without any roi is aligned by features first (src/optflow.cpp:366-377), its map moved by
    the same affine (:411-444), and its matches take the `features` branch of random_points (:544-550).
    The driver and the Python mirror run the same library: bit-equal planes and records."""
    import cv2
    from fibsem_optflow_b200 import api
    exe = build_cli()
    h, w, m = 420, 560, 60
    c = synth.to_u8(synth.texture(h + 2 * m, w + 2 * m, 3, 2.0, coarse=8))
    f0 = np.ascontiguousarray(c[m:m + h, m:m + w])
    M = np.array([[1.01, -0.006, 4.2 + m], [0.006, 1.01, -2.7 + m]])
    f1 = cv2.warpAffine(c, M, (w, h), flags=cv2.INTER_CUBIC | cv2.WARP_INVERSE_MAP)
    n0, n1 = str(tmp_path / "f0.png"), str(tmp_path / "f1.png")
    write_png(n0, f0)
    write_png(n1, f1)
    common = {"debug": True, "scale": 1.0, "lambda": 0.15, "nscales": 3, "npoints": 5, "output_dir": str(tmp_path)}
    job = dict(common, images=[
        {"p": n0, "q": n1, "output_name": "m", "output_type": "map"},
        {"p": n0, "q": n1, "output_name": "f", "output_type": "flow", "features": 1, "rois": {"top": 80}},
        {"p": n0, "q": n1, "output_type": "random_points", "pId": "a", "qId": "b", "pGroupId": "1.0", "qGroupId": "2.0"}])
    jf = str(tmp_path / "job.json")
    json.dump(job, open(jf, "w"))
    subprocess.check_call([exe, jf], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    args = dict(common)
    want = api.solve_rois(f0, f1, {"default": [0, 0, w, h]}, {"output_type": "map"}, args)["default"]
    assert np.array_equal(read_tiff_f32(str(tmp_path / "m_1.00_x.tiff")), want[0])
    assert np.array_equal(read_tiff_f32(str(tmp_path / "m_1.00_y.tiff")), want[1])
    assert abs(np.median(want[0] - np.arange(w, dtype=np.float32)[None, :]) + 4.2) < 1.5
    want = api.solve_rois(f0, f1, {"top": [0, 0, w, 80]}, {"output_type": "flow", "features": 1}, args)["top"]
    assert np.array_equal(read_tiff_f32(str(tmp_path / "f_1.00_top_x.tiff")), want[0])
    assert np.array_equal(read_tiff_f32(str(tmp_path / "f_1.00_top_y.tiff")), want[1])
    im = {"output_type": "random_points", "pId": "a", "qId": "b", "pGroupId": "1.0", "qGroupId": "2.0"}
    api.solve_rois(f0, f1, {"default": [0, 0, w, h]}, im, args, seed=-1)
    got = json.load(open(str(tmp_path / "point_matches_000.json")))
    assert got == args["point_matches"]
    api.release_solvers()
