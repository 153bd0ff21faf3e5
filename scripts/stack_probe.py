"""Developer probe: where a pair of the stack runner spends its time beyond the solve
(python scripts/stack_probe.py [size] [pairs])."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fibsem_optflow_b200 import _native as N, synth

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 24
sl = [torch.from_numpy(a).pin_memory() for a in synth.make_stack(8, S, S, seed=100)]
ring = [torch.empty((S, S), dtype=torch.float32).pin_memory() for _ in range(8)]
period = 2 * (len(sl) - 1)
idx = [(k % period) if (k % period) < len(sl) else period - (k % period) for k in range(NP + 1)]
s = N.Solver(N.default_params(lambda_=0.15, nscales=5, warps=5, inner_iterations=30, outer_iterations=10))


def run(flows, mask, npoints):
    kw = dict(slices=None, flows=flows, apply_mask=mask, npoints=npoints, scale=0.5, seed=1,
              slice_ptrs=[sl[i].data_ptr() for i in idx], pitch=S, shape=(S, S))
    if flows:
        kw.update(out_u=[ring[(2 * k) % 8].data_ptr() for k in range(NP)], out_v=[ring[(2 * k + 1) % 8].data_ptr() for k in range(NP)])
    s.run_stack(**kw)
    t = time.perf_counter()
    r = s.run_stack(**kw)
    dt = (time.perf_counter() - t) * 1e3
    solve = sum(x.ms_total for x in r["stats"])
    return dt / NP, solve / NP, sum(x.total_iterations for x in r["stats"]) / NP


for name, a in (("solve only (no flows, no mask, no matches)", (False, False, -1)),
                ("+ mask", (False, True, -1)),
                ("+ mask + 25 matches", (False, True, 25)),
                ("+ flow download", (True, False, -1)),
                ("all (the bench's stack leg)", (True, True, 25))):
    per, solve, it = run(*a)
    print("%-46s %.2f ms per pair (solve %.2f ms, %.0f iterations)" % (name, per, solve, it))
