"""Occupancy sweep of the Schur tile kernel (EMBA_SCHUR_OCC; read per call). usage: python tools/schur_variants.py C4"""
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
eng.evaluate(0, 0, 1.0, 5.0)
eng.form_normal_eq(5, 0, 1.0, 5.0)
ref = None
for occ in (4, 5, 6, 7, 4, 5, 6, 7):
    os.environ["EMBA_SCHUR_OCC"] = str(occ)
    ts = []
    for i in range(6):
        x1, x2, _, _ = eng.solve(1e-3, False, True)
        if i >= 2:
            ts.append(eng.timings_ms()["solve"])
    if ref is None:
        ref = (x1.copy(), x2.copy())
    print(f"{name} EMBA_SCHUR_OCC={occ}: solve {np.mean(ts):.3f} ms (min {np.min(ts):.3f}); x1/x2 bit-identical: "
          f"{np.array_equal(x1, ref[0]) and np.array_equal(x2, ref[1])}", flush=True)
eng.close()
