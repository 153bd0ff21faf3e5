"""Developer probe: solver-loop overhead.  epsilon = 0 (never stops), one level, one warp: every enqueued
iteration really runs, so stage time minus (iterations x kernel time) is launch / schedule overhead."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fibsem_optflow_b200 import _native as N, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
I0, I1 = synth.make_pair(n, n, seed=7, shear=4.0 / n)
for inner, outer in ((30, 1), (30, 4), (2, 15)):
    s = N.Solver(N.default_params(lambda_=0.15, nscales=1, warps=1, epsilon=0.0, inner_iterations=inner, outer_iterations=outer))
    s.set_timing(True)
    for rep in range(2):
        s.calc(I0, I1)
    st = s.stats
    print(f"{n}^2 inner {inner} outer {outer}: iterations {st.total_iterations} launches {st.launches} iterate {st.ms_iterate:.3f} ms "
          f"-> {st.ms_iterate / st.total_iterations * 1e3:.1f} us/iteration; median {st.ms_median:.3f} warp {st.ms_warp:.3f} total {st.ms_total:.3f}")
    s.close()
