"""Developer timing probe (not the contract bench): one pair per size, stats from the C ABI."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fibsem_optflow_b200 import _native as N, synth

def run(h, w, nscales, reps=2):
    dx, dy = float(os.environ.get("DX", 1.3)), float(os.environ.get("DY", -0.7))
    t = time.time(); I0, I1 = synth.make_pair(h, w, seed=7, dx=dx, dy=dy, shear=4.0/h, margin=int(os.environ.get("MARGIN", 32))); tg = time.time() - t
    if os.environ.get("MASK_FRAC"):   # a zero band like the padding of aligned FIB-SEM frames
        m = int(h * float(os.environ["MASK_FRAC"]))
        I0 = I0.copy(); I1 = I1.copy()
        I0[:m] = 0; I1[:m] = 0
        I0[:, :m // 2] = 0; I1[:, :m // 2] = 0
    s = N.Solver(N.default_params(lambda_=0.15, nscales=nscales, warps=int(os.environ.get('WARPS', 5)), inner_iterations=30, outer_iterations=10))
    if os.environ.get("FUSED_MIN_PX"):
        s.set_option("fused_min_px", float(os.environ["FUSED_MIN_PX"]))
    s.set_timing(True)
    for r in range(reps):
        t = time.time(); u, v = s.calc(I0, I1); dt = time.time() - t
        st = s.stats
        print(f"{w}x{h} S={nscales} rep{r}: wall {dt*1e3:.1f} ms  gpu {st.ms_total:.2f} ms  pyr {st.ms_pyramid:.2f} warp {st.ms_warp:.2f} "
              f"iter {st.ms_iterate:.2f} med {st.ms_median:.2f} other {st.ms_other:.2f}  iters {st.total_iterations} launches {st.launches} "
              f"Mpx/s(gpu) {w*h/st.ms_total/1e3:.1f}  alg GB {st.algorithmic_bytes/1e9:.2f} -> {st.algorithmic_bytes/st.ms_total/1e6:.0f} GB/s (gen {tg:.1f}s)")
        its = st.iters_array()
        for l in range(st.levels):
            px = st.width[l]*st.height[l]; n = int(its[l].sum())
            ms = st.ms_iterate_level[l]
            if ms > 0:
                print(f"   L{l} {st.width[l]}x{st.height[l]} iters {its[l].tolist()} iter-ms {ms:.2f}  {64.0*px*n/ms/1e6:.0f} GB/s  ({ms/n*1e3:.1f} us/iter)")
    ut, vt = synth.true_flow(h, w, dx=dx, dy=dy, shear=4.0/h, margin=int(os.environ.get("MARGIN", 32)))
    epe = np.hypot(u-ut, v-vt)
    print(f"   EPE vs truth mean {epe.mean():.4f} interior max {epe[16:-16,16:-16].max():.4f}")
    s.close()

if __name__ == "__main__":
    sizes = sys.argv[1:] or ["2048:5", "8192:6"]
    for a in sizes:
        n, sc = a.split(":")
        run(int(n), int(n), int(sc))
