"""Developer probe: the production shape of gen_cross_file_list.py jobs -- top / bottom ROI strips of
100 rows at scale 0.5, wrapper defaults (lambda .05, 10 scales, 300 iterations)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fibsem_optflow_b200 import _native as N, synth

for (h, w) in ((100, 4096), (200, 4096), (100, 8192)):
    I0, I1 = synth.make_pair(h, w, seed=5)
    s = N.Solver(N.default_params())          # reference-wrapper defaults
    if os.environ.get("FUSED_MIN_PX"):
        s.set_option("fused_min_px", float(os.environ["FUSED_MIN_PX"]))
    s.set_timing(True)
    for rep in range(3):
        t = time.time(); u, v = s.calc(I0, I1); dt = time.time() - t
    st = s.stats
    print(f"{w}x{h}: wall {dt*1e3:.2f} ms gpu {st.ms_total:.2f} ms levels {st.levels} iterations {st.total_iterations} launches {st.launches} "
          f"-> {st.ms_iterate / max(st.total_iterations, 1) * 1e3:.1f} us/iteration  ({w*h/st.ms_total/1e3:.1f} Mpx/s)")
    s.close()
