"""One job over the GPUs of a box: `python -m fibsem_optflow_b200.run_job --gpus N job.json[.gz]`.

Slice pairs are independent (reference src/optflow.cpp:75-178 walks "images" one pair at a time), so a job
shards by pair: this launcher starts one `host/optflow_b200 --device r --shard r/N` process per GPU -- rank r
solves its contiguous block of "images" (api.shard_pairs: the slice two adjacent pairs share stays on one GPU)
and writes its own outputs (flow / map TIFFs by output name, point_matches_r<r>of<N>_<k>.json batches).
Nothing is exchanged between the ranks: no NCCL, no collective.
"""
import argparse
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "host", "optflow_b200")


def commands(job, gpus, devices=None, prefetch=None, timing=False):
    """The per-rank command lines (rank r -> device devices[r], default r)."""
    devices = list(range(gpus)) if devices is None else list(devices)
    if len(devices) != gpus:
        raise ValueError("need one device id per rank")
    out = []
    for r in range(gpus):
        c = [EXE, "--device", str(devices[r]), "--shard", "%d/%d" % (r, gpus)]
        if prefetch is not None:
            c += ["--prefetch", str(int(prefetch))]
        if timing:
            c.append("--timing")
        out.append(c + [job])
    return out


def run(job, gpus, devices=None, prefetch=None, timing=False):
    """Runs the ranks side by side; returns the list of exit codes (0 = that shard completed)."""
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-C", os.path.dirname(EXE), "-s", "optflow_b200"])
    procs = [subprocess.Popen(c) for c in commands(job, gpus, devices, prefetch, timing)]
    return [p.wait() for p in procs]


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("job")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--devices", default=None, help="comma-separated device ids, one per rank (default 0..gpus-1)")
    ap.add_argument("--prefetch", type=int, default=None)
    ap.add_argument("--timing", action="store_true")
    ap.add_argument("--dry-run", action="store_true", help="print the per-rank command lines and exit")
    a = ap.parse_args(argv)
    devs = [int(x) for x in a.devices.split(",")] if a.devices else None
    if a.dry_run:
        for c in commands(a.job, a.gpus, devs, a.prefetch, a.timing):
            print(" ".join(c))
        return 0
    rcs = run(a.job, a.gpus, devs, a.prefetch, a.timing)
    bad = [r for r, rc in enumerate(rcs) if rc != 0]
    if bad:
        print("ranks failed: %s" % bad, file=sys.stderr)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
