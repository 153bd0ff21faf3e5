"""The C-ABI library loads and exports every symbol include/tvl1_b200.h declares; without a
GPU the compute entry points fail loudly instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "tvl1_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tvl1_[a-z0-9_]+)\s*\(", src)))


def test_exports_match_header(native):
    names = declared_symbols()
    assert len(names) >= 25
    L = native.lib()
    for n in names:
        assert hasattr(L, n), "libtvl1_b200.so does not export %s" % n
    assert sorted(native.EXPORTS) == names


def test_struct_sizes(native):
    # layout agreed between the header and the ctypes mirror
    assert C.sizeof(native.Params) == 6 * 8 + 10 * 4
    p = native.default_params()
    assert (p.tau, p.lambda_, p.theta, p.nscales, p.warps, p.epsilon, p.iterations, p.scale_step,
            p.gamma) == (0.25, 0.05, 0.3, 10, 5, 0.01, 300, 0.8, 0.0)   # src/optflow.cpp:503-511


def test_version(native):
    assert b"sm_100a" in native.lib().tvl1_version()


def test_pyramid_sizes_host(native, orc):
    for (w, h, n) in [(2048, 2048, 5), (8192, 8192, 6), (96, 128, 10), (19, 300, 4), (16384, 16384, 8)]:
        assert native.pyramid_sizes(w, h, n, 0.8) == orc.pyramid_sizes(w, h, n, 0.8)


def test_glibc_rand_stream(native):
    # the sampler's jump-ahead generator against the real libc, seeded and unseeded-equivalent
    libc = C.CDLL("libc.so.6")
    for seed in (1, 12345, 1539000000):
        libc.srand(seed)
        want = [libc.rand() for _ in range(5000)]
        assert native.glibc_rand(seed, 0, 64).tolist() == want[:64]
        for skip in (1, 31, 1000, 4937):
            assert native.glibc_rand(seed, skip, 40).tolist() == want[skip:skip + 40]
    assert native.glibc_rand(-1, 0, 3).tolist() == [1804289383, 846930886, 1681692777]


def test_glibc_rand_far_jump(native):
    import json
    known = json.load(open(os.path.join(ROOT, "tests", "golden", "known_answers.json")))["glibc_rand"]
    for seed, rec in known.items():
        if not isinstance(rec, dict):
            continue
        assert native.glibc_rand(int(seed), 0, 8).tolist() == rec["first"]
        assert native.glibc_rand(int(seed), 99999, 1).tolist() == [rec["at_99999"]]


def test_no_cpu_fallback(native):
    if native.device_count() > 0:
        pytest.skip("a GPU is present")
    p = native.default_params()
    h = C.c_void_p()
    rc = native.lib().tvl1_create(C.byref(p), 0, C.byref(h))
    assert rc == -2 and h.value is None
    assert b"no CPU path" in native.lib().tvl1_last_error()
    with pytest.raises(native.Tvl1Error):
        native.Solver()
    with pytest.raises(native.Tvl1Error):
        native.k_median5(np.zeros((8, 8), np.float32))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "fibsem_optflow_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                s = open(os.path.join(dp, f), errors="ignore").read()
                for line in s.splitlines():
                    t = line.strip()
                    if t.startswith(("import ", "from ", "#include")):
                        assert "oracle" not in t, "%s: %s" % (f, t)
