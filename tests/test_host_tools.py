"""Host-side pieces of the job driver (N1), checked without a GPU through host/hosttool: the
comment-tolerant JSON reader (the reference reads its job files with jsoncpp in non-strict mode,
src/optflow.cpp:32-58, and docs/example.json relies on comments), jsoncpp's alphabetical member order
(which fixes the order the reference walks "rois" in, :339), the PNG / PGM / TIFF readers that stand in
for cv::imread(IMREAD_GRAYSCALE) (:106,:119) and the float TIFF writer (:480-481)."""
import gzip
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "fibsem_optflow_b200", "host")
TOOL = os.path.join(HOST, "hosttool")


@pytest.fixture(scope="module")
def tool():
    subprocess.check_call(["make", "-C", HOST, "-s", "hosttool"])
    return TOOL


def test_json_comments_order_numbers(tool, tmp_path):
    text = """// job file in the style of docs/example.json
    {
      "style": 1, /* block comment */ "debug": false,
      "rois": {"top": 100, "bottom": 120, "custom": [1, 2, 3, 4]},   // alphabetical walk: bottom, custom, top
      "scale": 0.5, "lambda": 0.15, "big": 12345678901, "neg": -3, "exp": 1e-3,
      "images": [ {"p": "a.png", "q": "b.png", "output_name": "a~b", }, ],
      "s": "quote \\" and \\\\ and \\u00e9"
    }"""
    p = tmp_path / "job.json"
    p.write_text(text)
    out = subprocess.check_output([tool, "json", str(p)]).decode()
    got = json.loads(out)
    assert got["rois"] == {"bottom": 120, "custom": [1, 2, 3, 4], "top": 100}
    assert list(got["rois"]) == ["bottom", "custom", "top"]          # jsoncpp member order
    assert list(got)[:3] == sorted(got)[:3]
    assert got["big"] == 12345678901 and got["neg"] == -3 and got["exp"] == 1e-3 and got["scale"] == 0.5
    assert got["images"][0]["output_name"] == "a~b"
    assert got["s"] == 'quote " and \\ and é'
    assert '"big" : 12345678901' in out.replace("  ", " ") or "12345678901" in out   # integers stay integers


def test_json_rejects_garbage(tool, tmp_path):
    for bad in ('{"images": [ {"p": "x" "q": "y"} ]}',      # missing comma, as in docs/example.json:72
                '{"a": 1', '[1, 2', '{"a": tru}', ''):
        p = tmp_path / "bad.json"
        p.write_text(bad)
        assert subprocess.call([tool, "json", str(p)], stderr=subprocess.DEVNULL) != 0, bad


def decode(tool, path):
    raw = subprocess.check_output([tool, "image", str(path)])
    head, _, body = raw.partition(b"\n")
    w, h = (int(x) for x in head.split())
    return np.frombuffer(body, np.uint8).reshape(h, w)


@pytest.mark.parametrize("shape", [(37, 53), (128, 200), (1, 1), (5, 1000), (300, 900)])
def test_image_readers_match_cv2(tool, tmp_path, shape):
    cv2 = pytest.importorskip("cv2")
    from PIL import Image
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    a = rng.integers(0, 256, size=shape, dtype=np.uint8)
    # 8-bit grey PNG (all five row filters occur on noise), PGM, uncompressed TIFF
    assert cv2.imwrite(str(tmp_path / "g.png"), a, [cv2.IMWRITE_PNG_COMPRESSION, 6])
    with open(tmp_path / "g.pgm", "wb") as f:
        f.write(b"P5\n# c\n%d %d\n255\n" % (shape[1], shape[0]) + a.tobytes())
    Image.fromarray(a).save(str(tmp_path / "g.tiff"), compression=None)
    for name in ("g.png", "g.pgm", "g.tiff"):
        assert np.array_equal(decode(tool, tmp_path / name), a), name
    # compressed TIFFs: cv::imwrite's own (LZW + horizontal predictor), PackBits, Deflate, LZW without predictor
    assert cv2.imwrite(str(tmp_path / "cv.tiff"), a)
    assert np.array_equal(decode(tool, tmp_path / "cv.tiff"), a)
    smooth = (np.add.outer(np.arange(shape[0]), np.arange(shape[1])) // 3 % 256).astype(np.uint8)   # long LZW strings
    assert cv2.imwrite(str(tmp_path / "cvs.tiff"), smooth)
    assert np.array_equal(decode(tool, tmp_path / "cvs.tiff"), smooth)
    for comp in ("packbits", "tiff_adobe_deflate", "tiff_lzw"):
        for arr in (a, smooth):
            Image.fromarray(arr).save(str(tmp_path / "p.tiff"), compression=comp)
            assert np.array_equal(decode(tool, tmp_path / "p.tiff"), arr), comp
    # Adam7-interlaced PNGs (grey and colour) and a palette PNG, written by PIL
    import zlib, struct

    def png_bytes(arr, ctype, interlace):
        """minimal PNG writer (filter 0) so that the Adam7 path is exercised independently of PIL"""
        hh, ww = arr.shape[:2]
        ch = 1 if arr.ndim == 2 else arr.shape[2]
        def chunk(tag, data):
            return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
        if not interlace:
            rows = b"".join(b"\0" + arr[y].tobytes() for y in range(hh))
        else:
            X0, Y0 = (0, 4, 0, 2, 0, 1, 0), (0, 0, 4, 0, 2, 0, 1)
            DX, DY = (8, 8, 4, 4, 2, 2, 1), (8, 8, 8, 4, 4, 2, 2)
            rows = b""
            for k in range(7):
                sub = arr[Y0[k]::DY[k], X0[k]::DX[k]]
                if sub.shape[0] and sub.shape[1]:
                    rows += b"".join(b"\0" + np.ascontiguousarray(sub[y]).tobytes() for y in range(sub.shape[0]))
        ihdr = struct.pack(">IIBBBBB", ww, hh, 8, ctype, 0, 0, 1 if interlace else 0)
        return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", zlib.compress(rows)) + chunk(b"IEND", b"")

    rgb_i = rng.integers(0, 256, size=shape + (3,), dtype=np.uint8)
    for (arr, ctype) in ((a, 0), (rgb_i, 2)):
        (tmp_path / "i.png").write_bytes(png_bytes(arr, ctype, True))
        want_i = cv2.imread(str(tmp_path / "i.png"), cv2.IMREAD_GRAYSCALE)
        assert want_i is not None and np.array_equal(decode(tool, tmp_path / "i.png"), want_i), ctype
    Image.fromarray(a).convert("P").save(str(tmp_path / "pal.png"))
    assert np.array_equal(decode(tool, tmp_path / "pal.png"), cv2.imread(str(tmp_path / "pal.png"), cv2.IMREAD_GRAYSCALE))
    g16t = rng.integers(0, 65536, size=shape, dtype=np.uint16)
    assert cv2.imwrite(str(tmp_path / "h.tiff"), g16t)
    assert np.array_equal(decode(tool, tmp_path / "h.tiff"), cv2.imread(str(tmp_path / "h.tiff"), cv2.IMREAD_GRAYSCALE))
    # colour and 16-bit PNGs go through the same conversion cv::imread(IMREAD_GRAYSCALE) applies
    rgb = rng.integers(0, 256, size=shape + (3,), dtype=np.uint8)
    assert cv2.imwrite(str(tmp_path / "c.png"), rgb)
    want = cv2.imread(str(tmp_path / "c.png"), cv2.IMREAD_GRAYSCALE)
    assert np.array_equal(decode(tool, tmp_path / "c.png"), want)
    g16 = rng.integers(0, 65536, size=shape, dtype=np.uint16)
    assert cv2.imwrite(str(tmp_path / "h.png"), g16)
    want = cv2.imread(str(tmp_path / "h.png"), cv2.IMREAD_GRAYSCALE)
    assert np.array_equal(decode(tool, tmp_path / "h.png"), want)


def test_image_reader_errors(tool, tmp_path):
    (tmp_path / "x.png").write_bytes(b"\x89PNG\r\n\x1a\n" + b"\0" * 10)
    assert subprocess.call([tool, "image", str(tmp_path / "x.png")], stderr=subprocess.DEVNULL) != 0
    assert subprocess.call([tool, "image", str(tmp_path / "missing.png")], stderr=subprocess.DEVNULL) != 0


def test_float_tiff_writer_roundtrip(tool, tmp_path):
    from PIL import Image
    rng = np.random.default_rng(4)
    a = rng.standard_normal((31, 45)).astype(np.float32)
    a[0, 0] = -0.0
    a[1, 1] = np.float32(1e-42)     # subnormal survives
    out = tmp_path / "f.tiff"
    subprocess.run([tool, "tiff", "45", "31", str(out)], input=a.tobytes(), check=True)
    b = np.array(Image.open(str(out)))
    assert b.dtype == np.float32 and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_driver_cli_without_gpu(tmp_path):
    """The job driver itself: --help needs no device; argument errors and unreadable job files fail
    loudly before any CUDA call; and without a device the driver refuses to run (there is no CPU path)."""
    subprocess.check_call(["make", "-C", HOST, "-s", "optflow_b200"])
    exe = os.path.join(HOST, "optflow_b200")
    out = subprocess.check_output([exe, "--help"]).decode()
    assert "--shard" in out and "--prefetch" in out
    assert subprocess.call([exe], stderr=subprocess.DEVNULL, stdout=subprocess.DEVNULL) != 0            # no job file
    assert subprocess.call([exe, "--shard", "3/2", "x.json"], stderr=subprocess.DEVNULL, stdout=subprocess.DEVNULL) != 0
    bad = tmp_path / "bad.json"
    bad.write_text('{"images": [ {"p": "x" "q": "y"} ]}')
    assert subprocess.call([exe, str(bad)], stderr=subprocess.DEVNULL, stdout=subprocess.DEVNULL) != 0
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "fibsem_optflow_b200", "csrc", "libtvl1_b200.so"))
    if lib.tvl1_dev_count() <= 0:
        good = tmp_path / "job.json"
        good.write_text(json.dumps({"images": [], "output_dir": str(tmp_path)}))
        p = subprocess.run([exe, str(good)], capture_output=True)
        assert p.returncode != 0 and b"no CUDA device" in p.stderr + p.stdout
