"""Timed jobs through the C++ driver (host/optflow_b200), the reference's `optflow <job.json>` CLI:
 (1) a gen_cross_file_list-style job: 8 chained pairs of 4096^2 PNG slices, scale 0.5, rois top/bottom
     (100-row strips at engine size), output_type random_points  -- the reference's production shape;
 (2) the same 8 pairs without rois/scale restrictions: whole 4096^2 frames at scale 1, random_points (matches only);
 (3) the same, output_type flow (two 64 MiB float TIFFs per pair).
Prints one JSON line per job: wall seconds of the driver process, ms per pair.
usage (on a GPU box): python scripts/job_probe.py [size] [pairs] > gpurun_out/job_probe.json"""
import json
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np
from fibsem_optflow_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "fibsem_optflow_b200", "host", "optflow_b200")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 8

subprocess.check_call(["make", "-C", os.path.dirname(EXE), "-s"])
tmp = tempfile.mkdtemp(prefix="job_probe_")
sl = synth.make_stack(NP, S, S, seed=300)
names = []
for k, a in enumerate(sl):
    p = os.path.join(tmp, "slice_%03d.png" % k)
    assert cv2.imwrite(p, a, [cv2.IMWRITE_PNG_COMPRESSION, 1])
    names.append(p)
png_mb = sum(os.path.getsize(p) for p in names) / 1e6


def run(tag, top):
    out = os.path.join(tmp, tag)
    os.makedirs(out, exist_ok=True)
    job = dict(top)
    job.update({"debug": True, "style": 1, "features": False, "output_dir": out,
                "images": [{"p": names[k], "q": names[k + 1], "output_name": "s%03d~s%03d" % (k, k + 1),
                            "p_tile": "t%d" % k, "q_tile": "t%d" % (k + 1), "p_group": "g", "q_group": "g"} for k in range(NP)]})
    jf = os.path.join(tmp, tag + ".json")
    with open(jf, "w") as f:
        json.dump(job, f)
    subprocess.check_call([EXE, jf], stdout=subprocess.DEVNULL)           # warm: CUDA context, page cache
    t = time.perf_counter()
    r = subprocess.run([EXE, "--timing", jf], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, check=True)
    dt = time.perf_counter() - t
    sys.stderr.write(tag + ": " + r.stderr)
    nfiles = len(os.listdir(out))
    print(json.dumps({"job": tag, "pairs": NP, "slice": [S, S], "png_mbytes": round(png_mb, 1), "wall_s": round(dt, 3),
                      "ms_per_pair": round(dt * 1e3 / NP, 1), "output_files": nfiles, "args": top}), flush=True)


run("cross_rois_scale_half_points", {"output_type": "random_points", "scale": 0.5, "rois": {"top": 100, "bottom": 100},
                                      "lambda": 0.05, "nscales": 10, "npoints": 25, "batch_size": 4})
run("whole_frame_points", {"output_type": "random_points", "scale": 1.0, "rois": {"custom": [0, 0, S, S]},
                           "lambda": 0.15, "nscales": 5, "npoints": 25, "batch_size": 4})
run("whole_frame_flow_tiffs", {"output_type": "flow", "scale": 1.0, "rois": {"custom": [0, 0, S, S]}, "lambda": 0.15, "nscales": 5})
