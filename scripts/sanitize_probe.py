"""Small solves for compute-sanitizer (memcheck / racecheck): every kernel, both schedules, a zero band."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fibsem_optflow_b200 import _native as N, synth

for (h, w, fused_min) in ((200, 333, 0), (150, 260, 1e18), (97, 131, 0)):
    I0, I1 = synth.make_pair(h, w, seed=3)
    I0 = I0.copy(); I0[:20] = 0
    s = N.Solver(N.default_params(lambda_=0.15, nscales=3, warps=2))
    s.set_option("fused_min_px", fused_min)
    u, v = s.calc(I0, I1)
    m = s.sample_matches(I0, I1, u, v, scale=0.5, npoints=9, seed=1)
    print(h, w, fused_min, s.stats.total_iterations, float(np.abs(u).mean()))
    s.close()
sl = synth.make_stack(3, 120, 200, seed=5)
s = N.Solver(N.default_params(lambda_=0.15, nscales=2, warps=2))
r = s.run_stack(sl, flows=True, apply_mask=True, npoints=5, scale=0.5, seed=3, prescale=0.5)
print("stack ok", len(r["u"]))
