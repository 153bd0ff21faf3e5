"""profiles/k_outer_traffic.json from the ncu --set full capture of one k_outer launch made by
scripts/ncu_r2.sh (kbench `outer`: KB_OUTER_ITERS iterations of an N^2 level in ONE launch).
usage: python scripts/make_traffic_json.py gpurun_out/prof_k_outer_<tag>.ncu-rep <tag> [N] [iters]"""
import csv
import json
import os
import subprocess
import sys

rep, tag = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 10
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]


def val(k):
    i = hdr.index(k)
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9,
                "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}.get(u, 1.0)


rd, wr, t = val("dram__bytes_read.sum"), val("dram__bytes_write.sum"), val("gpu__time_duration.sum")
inst = val("smsp__inst_executed.sum")
px_it = float(n) * n * iters
d = {"kernel": "k_outer<4>, one launch = %d iterations of a %dx%d level (%s)" % (iters, n, n, r[hdr.index("Kernel Name")]),
     "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
     "iterations_in_launch": iters, "px": n * n,
     "dram_bytes_per_px_iteration": (rd + wr) / px_it,
     "algorithmic_bytes_per_px_iteration": 64, "compulsory_bytes_per_px_iteration_fused": 30,
     "launch_ms_under_ncu": t * 1e3, "dram_gbs_under_ncu": (rd + wr) / t / 1e9,
     "thread_instructions_per_px_iteration": inst * 32 / px_it,
     "source": "ncu --set full --clock-control none, %s (scripts/ncu_r2.sh, tag %s)" % (os.path.basename(rep), tag)}
json.dump(d, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "k_outer_traffic.json"), "w"), indent=1)
print(json.dumps(d, indent=1))
