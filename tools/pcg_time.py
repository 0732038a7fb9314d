"""PCG timing: iterations, error and wall time per iteration, persistent cooperative kernel against the multi-launch
chain (EMBA_CG_PERSIST; read per call). usage: python tools/pcg_time.py C2"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from emba_b200 import synth
from emba_b200.legm import Engine, spline_base_ns
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
sc = synth.make_config(name, device="cuda")
eng = Engine(sc.sensor_w, sc.sensor_h, sc.bearing_lut(), sc.C_th, sc.pano_w, sc.pano_h)
eng.set_events(sc.x, sc.y, sc.t_ns, sc.pol)
t0, dt = spline_base_ns(sc.t_beg, sc.dt_knots)
eng.set_state(0, t0, dt, sc.quat_init, sc.Gx_init, sc.Gy_init)
eng.evaluate(0, 0, 1.0, 5.0)
eng.form_normal_eq(5, 0, 1.0, 5.0)
ref = None
for mode in ("1", "0", "1", "0"):
    os.environ["EMBA_CG_PERSIST"] = mode
    ws = []
    for i in range(5):
        w = time.perf_counter(); x1, x2, it, err = eng.solve(1e-3, True, True); w = time.perf_counter() - w
        if i: ws.append(w)
    if ref is None:
        ref = (x1.copy(), x2.copy(), it)
    same = np.array_equal(x1, ref[0]) and np.array_equal(x2, ref[1]) and it == ref[2]
    print(f"{name} EMBA_CG_PERSIST={mode}: {it} iterations, err {err:.3e}, {np.mean(ws)*1e3:.3f} ms per solve = "
          f"{np.mean(ws)*1e6/max(it,1):.1f} us per iteration (incl. setup and result copies); bit-identical to the first: {same}", flush=True)
eng.close()
