"""Per-phase device timings of the hot path on a synthetic workload (GPU box). usage: python tools/time_phases.py C2"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from emba_b200 import synth
from emba_b200.legm import Engine, spline_base_ns
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
over = {}
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    over[k] = float(v) if "." in v or "e" in v else int(v)
sc = synth.make_config(name, device="cuda", **over)
eng = Engine(sc.sensor_w, sc.sensor_h, sc.bearing_lut(), sc.C_th, sc.pano_w, sc.pano_h)
t = time.perf_counter(); eng.set_events(sc.x, sc.y, sc.t_ns, sc.pol); t_ev = time.perf_counter() - t
t0, dt = spline_base_ns(sc.t_beg, sc.dt_knots)
t = time.perf_counter(); eng.set_state(0, t0, dt, sc.quat_init, sc.Gx_init, sc.Gy_init); t_st = time.perf_counter() - t
print(f"{name}: N={sc.n_events} n={sc.n_poses} set_events {t_ev*1e3:.1f} ms, first set_state (static rebuild) {t_st*1e3:.1f} ms; device ms {eng.setup_ms()}")
t = time.perf_counter(); eng.set_events(sc.x, sc.y, sc.t_ns, sc.pol); t_ev = time.perf_counter() - t
t = time.perf_counter(); eng.set_state(0, t0, dt, sc.quat_init, sc.Gx_init, sc.Gy_init); t_st = time.perf_counter() - t
print(f"  second window (arenas warm, pageable source): set_events {t_ev*1e3:.1f} ms, set_state {t_st*1e3:.1f} ms; device ms {eng.setup_ms()}")
import torch
xp, yp, tp, pp = (torch.from_numpy(a).pin_memory().numpy() for a in (sc.x, sc.y, sc.t_ns, sc.pol))
t = time.perf_counter(); eng.set_events(xp, yp, tp, pp); t_ev = time.perf_counter() - t
t = time.perf_counter(); eng.set_state(0, t0, dt, sc.quat_init, sc.Gx_init, sc.Gy_init); t_st = time.perf_counter() - t
print(f"  pinned source: set_events {t_ev*1e3:.1f} ms, set_state {t_st*1e3:.1f} ms; device ms {eng.setup_ms()}")
from emba_b200.legm import EventSequence
seq = EventSequence(xp, yp, tp, pp)
t = time.perf_counter(); eng.set_events_dev(seq, 0, sc.n_events); t_ev = time.perf_counter() - t
t = time.perf_counter(); eng.set_state(0, t0, dt, sc.quat_init, sc.Gx_init, sc.Gy_init); t_st = time.perf_counter() - t
print(f"  device-resident sequence: set_events_dev {t_ev*1e3:.1f} ms, set_state {t_st*1e3:.1f} ms; device ms {eng.setup_ms()}")
seq.close()
for rep in range(3):
    w = time.perf_counter(); cd, cr, M = eng.evaluate(0, 0, 1.0, 5.0); w_e = time.perf_counter() - w
    te = eng.timings_ms()
    w = time.perf_counter(); Np = eng.form_normal_eq(5, 0, 1.0, 5.0); w_f = time.perf_counter() - w
    tf = eng.timings_ms()
    w = time.perf_counter(); eng.solve(1e-3, False, True, want=False); w_s = time.perf_counter() - w
    ts = eng.timings_ms()
    w = time.perf_counter(); eng.make_candidate(1.0, True); w_c = time.perf_counter() - w
    w = time.perf_counter(); eng.evaluate(1, 0, 1.0, 5.0); w_e2 = time.perf_counter() - w
    print(f"rep{rep}: M={M} Np={Np} a12={eng.a12_entries()} | evaluate dev {te['evaluate']:.3f} (k_eval {te['eval_kernel']:.3f}) wall {w_e*1e3:.3f} | "
          f"form dev {tf['form']:.3f} (asm {tf['asm_pose_kernel']:.3f}, map {tf['map_side']:.3f}, pix {tf['pix_kernel']:.3f}, place+segsort {tf['sort']:.3f}) wall {w_f*1e3:.3f} | "
          f"solve dev {ts['solve']:.3f} wall {w_s*1e3:.3f} | candidate wall {w_c*1e3:.3f} | eval cand wall {w_e2*1e3:.3f}")
w = time.perf_counter(); log, fc = eng.solve_time_window(max_num_iter=9, alpha=5.0, thres=5); w = time.perf_counter() - w
print(f"LM: {log.shape[0]} solves, {int(log[:,4].sum())} accepted, {w*1e3/log.shape[0]:.3f} ms/iteration, cost {log[0,2]:.1f} -> {fc:.1f}")
eng.evaluate(0, 0, 1.0, 5.0); eng.form_normal_eq(5, 0, 1.0, 5.0)
w = time.perf_counter(); x = eng.solve(1e-3, True, True, want=False); w = time.perf_counter() - w
print(f"PCG: iters {x[2]} err {x[3]:.3e} wall {w*1e3:.2f} ms")
