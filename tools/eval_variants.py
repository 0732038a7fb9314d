"""A/B of the k_eval software-prefetch variants (EMBA_EVAL_PF, csrc/eval.cu) in one process on one workload.
usage: python tools/eval_variants.py C4"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from emba_b200 import synth
from emba_b200.legm import Engine, spline_base_ns
name = sys.argv[1] if len(sys.argv) > 1 else "C4"
sc = synth.make_config(name, device="cuda")
eng = Engine(sc.sensor_w, sc.sensor_h, sc.bearing_lut(), sc.C_th, sc.pano_w, sc.pano_h)
eng.set_events(sc.x, sc.y, sc.t_ns, sc.pol)
t0, dt = spline_base_ns(sc.t_beg, sc.dt_knots)
eng.set_state(0, t0, dt, sc.quat_init, sc.Gx_init, sc.Gy_init)
ref = None
for pf, gm in ((4, 32), (4, 8), (0, 32), (5, 32), (6, 32), (1, 32), (4, 32)):
    os.environ["EMBA_EVAL_PF"] = str(pf)
    os.environ["EMBA_EVAL_GRID"] = str(gm)
    ts = []
    for i in range(10):
        cd, cr, M = eng.evaluate(0, 0, 1.0, 5.0)
        if i >= 3:
            ts.append(eng.timings_ms()["eval_kernel"])
    _, num = eng.get_evaluation(0, None, False, True)
    if ref is None:
        ref = (cd, M, num.copy())
    same = cd == ref[0] and M == ref[1] and np.array_equal(num, ref[2])
    print(f"{name} EMBA_EVAL_PF={pf} EMBA_EVAL_GRID={gm}: k_eval {np.mean(ts):.3f} ms (min {np.min(ts):.3f}); cost/M/num_ev_map identical to PF=0: {same}", flush=True)
eng.close()
