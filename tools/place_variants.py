"""A/B of the k_place variants (EMBA_PLACE_V, csrc/assemble.cu) in one process; wrap with
ncu --metrics gpu__time_duration.sum -k regex:"k_place|k_seg_sort" for standalone kernel times.
usage: python tools/place_variants.py C4"""
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
for v in (0, 1, 2, 3, 4, 0):
    os.environ["EMBA_PLACE_V"] = str(v)
    ts, tf = [], []
    for i in range(5):
        eng.evaluate(0, 0, 1.0, 5.0)
        eng.form_normal_eq(5, 0, 1.0, 5.0)
        if i >= 2:
            t = eng.timings_ms(); ts.append(t["sort"]); tf.append(t["form"])
    A11, _, A22, b1, b2, _ = eng.get_normal_eq(False)
    if ref is None:
        ref = (A22.copy(), b2.copy())
    same = np.array_equal(A22, ref[0]) and np.array_equal(b2, ref[1])
    print(f"{name} EMBA_PLACE_V={v}: side stream {np.mean(ts):.3f} ms, form {np.mean(tf):.3f} ms; A22/b2 bit-identical: {same}", flush=True)
eng.close()
