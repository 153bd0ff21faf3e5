"""Host-side mirror of the reference interface (no GPU): parameter resolution, match records,
pair sharding (incl. a world_size-2 gloo run)."""
import os
import sys

import numpy as np
import pytest

from fibsem_optflow_b200 import api


def test_generate_TV_args_precedence():
    # per-pair overrides global overrides default (reference src/optflow.cpp:503-512)
    args = {"lambda": 0.1, "nscales": 6, "warps": 7}
    im = {"lambda": 0.2, "iterations": 90}
    tv = api.generate_TV_args(im, args)
    assert tv["lambda"] == 0.2 and tv["nscales"] == 6 and tv["warps"] == 7 and tv["iterations"] == 90
    assert tv["tau"] == 0.25 and tv["theta"] == 0.3 and tv["epsilon"] == 0.01
    assert tv["scaleStep"] == 0.8 and tv["gamma"] == 0.0 and tv["useInitialFlow"] is False
    assert api.generate_TV_args({}, {}) == api.TV_DEFAULTS


def test_params_from_TV_args(native):
    p = api.params_from_TV_args(api.generate_TV_args({"iterations": 95, "useInitialFlow": True}, {}))
    assert p.iterations == 95 and p.use_initial_flow == 0 and p.median_filtering == 5
    p = api.params_from_TV_args(api.generate_TV_args({}, {"innerIterations": 20, "outerIterations": 3,
                                                          "medianFiltering": 1}))
    assert (p.inner_iterations, p.outer_iterations, p.median_filtering) == (20, 3, 1)


def test_move_pm():
    args = {}
    im = {"pGroupId": "1.0", "pId": "a", "qGroupId": "2.0", "qId": "b",
          "point_matches": {"p": [[1.0], [2.0]], "q": [[1.5], [2.5]], "w": [1]}}
    api.move_pm(im, args)
    api.move_pm({"pGroupId": "2.0", "pId": "b", "qGroupId": "3.0", "qId": "c",
                 "point_matches": {"p": [[], []], "q": [[], []], "w": []}}, args)
    assert len(args["point_matches"]) == 2
    assert args["point_matches"][0]["matches"]["q"] == [[1.5], [2.5]]
    assert set(args["point_matches"][0]) == {"pGroupId", "pId", "qGroupId", "qId", "matches"}
    assert im["point_matches"] == {}


@pytest.mark.parametrize("n,ws", [(512, 8), (10, 4), (3, 8), (0, 2), (7, 1)])
def test_shard_pairs_partition(n, ws):
    parts = [list(api.shard_pairs(n, ws, r)) for r in range(ws)]
    flat = [i for p in parts for i in p]
    assert flat == list(range(n))                     # contiguous blocks, in rank order
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    mine = list(api.shard_pairs(37, world, rank))
    # what bench.py does across ranks: barrier, then MAX of the per-rank time and SUM of the work
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    n = torch.tensor([float(len(mine))], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    q.put((rank, mine, t.item(), n.item()))
    dist.destroy_process_group()


def test_sharding_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] + res[1][1] == list(range(37))
    assert res[0][2] == res[1][2] == 2.0 and res[0][3] == res[1][3] == 37.0


def test_run_job_commands():
    """the multi-GPU job launcher: one driver process per GPU, contiguous shards, nothing shared"""
    from fibsem_optflow_b200 import run_job
    cs = run_job.commands("job.json", 4, devices=[2, 3, 0, 1], prefetch=6, timing=True)
    assert len(cs) == 4
    for r, c in enumerate(cs):
        assert c[0].endswith("optflow_b200") and c[-1] == "job.json"
        assert c[c.index("--shard") + 1] == "%d/4" % r
        assert c[c.index("--device") + 1] == str([2, 3, 0, 1][r])
        assert "--timing" in c and c[c.index("--prefetch") + 1] == "6"
    import pytest
    with pytest.raises(ValueError):
        run_job.commands("job.json", 2, devices=[0])
    assert run_job.main(["--gpus", "2", "--dry-run", "j.json"]) == 0
