"""Runs a few passes (evaluate + form) on a workload; meant to be wrapped by ncu for a launch list. usage: quick_form.py C4 [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emba_b200 import synth
from emba_b200.legm import Engine, spline_base_ns
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sc = synth.make_config(name, device="cuda")
eng = Engine(sc.sensor_w, sc.sensor_h, sc.bearing_lut(), sc.C_th, sc.pano_w, sc.pano_h)
eng.set_events(sc.x, sc.y, sc.t_ns, sc.pol)
t0, dt = spline_base_ns(sc.t_beg, sc.dt_knots)
eng.set_state(0, t0, dt, sc.quat_init, sc.Gx_init, sc.Gy_init)
for _ in range(reps):
    eng.evaluate(0, 0, 1.0, 5.0)
    eng.form_normal_eq(5, 0, 1.0, 5.0)
print(name, eng.counters(), eng.timings_ms())
import numpy as np
_, num = eng.get_evaluation(0, None, False, True)
h = num.reshape(-1); h = h[h >= 5]
print("segment length percentiles 50/90/99/99.9/max:", [int(np.percentile(h, q)) for q in (50, 90, 99, 99.9)], int(h.max()),
      "frac of pixels > 1024:", float((h > 1024).mean()), "frac of rows in them:", float(h[h > 1024].sum() / h.sum()))
eng.solve(1e-3, False, True, want=False)
