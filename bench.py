#!/usr/bin/env python
"""Benchmark of the EMBA LM hot path on B200 (contract: see the task statement / DESIGN.md section "Measurement").

A "step" is one PASS of the hot path over the whole event window of the workload, on a fixed state:
    LEGM::evaluateDataError(eval_deriv=true) + formNormalEq + applyL2Reg
    (residuals, num_ev_map, cost, per-measurement Jacobian rows, A11/A12/A22/b1/b2).
metric = events/s for that pass (whole job, all GPUs). The LM-iteration time of a full LM run (to convergence, with
its accept/reject mix) is reported beside it, next to the reference's CPU LM iteration.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C4|C3|C2|C1|small|tiny] [--weak] [--impl reference]

Workload: C4 = BASELINE.json configs[3], the configuration the metric is quoted on (240x180 sensor, 2048x1024
panorama, ~100 M synthetic events, n = 201 control poses); it fits one GPU. N > 1 (torchrun, one rank per GPU): the
SAME window is time-sharded over the ranks (contiguous control-pose slices), partial systems combined over
NVLink inside the library -> strong scaling. --weak keeps the round-1 curve (window grows with N) for comparison.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALPHA, THRES = 5.0, 5
LM_KW = dict(max_num_iter=50, tol_fun=1e-3, num_times_tol_fun_sat=2)  # launch/playroom.launch:23-60
METRIC = "events/s for residual+Jacobian+H assembly"


def scene_kwargs(workload, n_gpus, weak):
    from emba_b200 import synth

    kw = dict(synth.CONFIGS[workload])
    if weak and n_gpus > 1:
        span = kw["t_end"] - kw["t_beg"]
        kw["t_end"] = kw["t_beg"] + span * n_gpus
        kw["periodic"] = True  # bounded yaw for long spans (keeps the view away from the poles)
    return kw


def config_of(sc, workload, world, weak):
    """the `config` object: identical keys and values in both arms"""
    name = workload if not (weak and world > 1) else f"{workload} x{world} time span"
    return {"workload": name, "events": int(sc.n_events), "sensor": [sc.sensor_w, sc.sensor_h],
            "panorama": [sc.pano_w, sc.pano_h], "control_poses": int(sc.n_poses),
            "parallelism": f"time-sharded x{world}" if world > 1 else "single GPU",
            "l2": "inputs_exceed_l2" if 44 * sc.n_events > 126e6 else "inputs_fit_l2"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms",
                                          "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def mark(self):
        """start of the timed region: only samples taken after this point are reported"""
        self.i0 = len(self.rows)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        rows = self.rows[getattr(self, "i0", 0):] or self.rows[-1:]
        for r in rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(a)
    return float(np.linalg.norm(a - b) / (den if den > 0 else 1.0))


# ------------------------------------------------------------------------------------------------------------------
# CPU side: the reference's own implementation (oracle/_ref = the unmodified reference sources compiled here), or the
# numpy port when that library is missing. Only this file's cpu_baseline / --impl reference legs use it.
# ------------------------------------------------------------------------------------------------------------------
def _ref_model(sc, n_events):
    from oracle import ref_binding as RB

    os.environ.setdefault("OMP_NUM_THREADS", "1")
    ref = RB.RefLEGM(sc.sensor_w, sc.sensor_h, sc.fx, sc.fy, sc.cx, sc.cy, sc.C_th, sc.pano_w, sc.pano_h)
    ref.set_events(sc.x[:n_events], sc.y[:n_events], sc.t_ns[:n_events], sc.pol[:n_events])
    return RB, ref


def cpu_reference_pass(sc, n_events, want_outputs=False):
    """One pass (evaluateDataError + formNormalEq + applyL2Reg) of the reference on the first n_events events.
    Returns dict(value events/s, kind, cores, seconds, [outputs])."""
    from oracle import ref_binding as RB

    n_events = (min(n_events, sc.n_events) // 100) * 100
    if RB.available():
        RB, ref = _ref_model(sc, n_events)
        tr = RB.RefTraj(sc.t_beg, sc.dt_knots, sc.quat_init)
        t = time.perf_counter()
        ep, num = ref.evaluate(tr, sc.Gx_init, sc.Gy_init, True)
        t_ev = time.perf_counter() - t
        t = time.perf_counter()
        A11, _, A22, b1, b2, act = ref.form(sc.n_poses, THRES, sc.Gx_init, sc.Gy_init, ALPHA, want_A12=False)
        t_form = time.perf_counter() - t
        out = dict(value=n_events / (t_ev + t_form), kind="reference", cores=1, seconds=t_ev + t_form, n_events=n_events,
                   t_evaluate=t_ev, t_form=t_form, ref=ref)
        if want_outputs:
            out.update(ep=ep, num=num, A11=A11, A22=A22, b1=b1, b2=b2, act=act,
                       cost=0.5 * float(ep @ ep) + ref.reg_cost(sc.Gx_init, sc.Gy_init, ALPHA))
        return out
    from oracle import emba_oracle as O

    orc = O.Oracle(sc.sensor_w, sc.sensor_h, sc.bearing_lut(), sc.pano_w, sc.pano_h, sc.C_th)
    n_events = min(n_events, 200_000)
    orc.set_events(sc.x[:n_events], sc.y[:n_events], sc.t_ns[:n_events], sc.pol[:n_events])
    t0, dtn = O.spline_base_ns(sc.t_beg, sc.dt_knots)
    t = time.perf_counter()
    ep, num = orc.evaluate(sc.quat_init, t0, dtn, sc.Gx_init, sc.Gy_init, True)
    t_ev = time.perf_counter() - t
    t = time.perf_counter()
    A = orc.form_normal_eq(sc.n_poses, THRES)
    A22, b2 = orc.apply_l2_reg(A[2], A[4], A[5], ALPHA, sc.Gx_init, sc.Gy_init)
    t_form = time.perf_counter() - t
    out = dict(value=n_events / (t_ev + t_form), kind="port", cores=1, seconds=t_ev + t_form, n_events=n_events,
               t_evaluate=t_ev, t_form=t_form, ref=None)
    if want_outputs:
        out.update(ep=ep, num=num, A11=A[0], A22=A22, b1=A[3], b2=b2, act=A[5],
                   cost=0.5 * float(ep @ ep) + orc.reg_cost(sc.Gx_init, sc.Gy_init, ALPHA))
    return out


def host_info():
    """nproc and CPU model of the box the CPU legs ran on"""
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.lower().startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return {"nproc": os.cpu_count(), "cpu_model": model}


def _omp_set_threads(k):
    """the compiled reference is built with -fopenmp (only Eigen's GEMM inside solveNormalEq has a parallel region)"""
    import ctypes
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(k))
        return True
    except OSError:
        return False


def cpu_lm_iteration_estimate(sc, cpu, Np_window, accepted_frac):
    """CPU LM-iteration time of the reference ON THIS WINDOW from the three regions it instruments itself
    (solver.cpp:105-151 form, :181-222 solve, :242-294 objective), measured on the prefix `cpu` was run on (with all
    n control poses) and scaled: the per-event regions (objective, form) linearly in N, the Schur solve linearly in
    the number of active pixels (its dense product W*A12^T costs (3n)^2 * 2Np * 2 flop, model.cpp:721-792)."""
    ref = cpu.get("ref")
    if ref is None:
        return None
    t = time.perf_counter()
    ref.solve(1e-3, False, True)
    t_solve = time.perf_counter() - t
    # the same solve with all host threads (only Eigen's GEMM in the Schur solve has a parallel region;
    # the single-thread figure stays the baseline of record)
    t_solve_all, nthr = None, min(os.cpu_count() or 1, 16)  # capped: the products are small, more threads only add overhead
    if nthr > 1 and _omp_set_threads(nthr):
        try:
            t = time.perf_counter()
            ref.solve(1e-3, False, True)
            t_solve_all = time.perf_counter() - t
        finally:
            _omp_set_threads(1)
    Np_p = max(1, ref.Np)
    scale_n = sc.n_events / cpu["n_events"]
    obj_ms = cpu["t_evaluate"] * scale_n * 1e3
    form_ms = cpu["t_form"] * scale_n * 1e3
    solve_ms = t_solve * (Np_window / Np_p) * 1e3
    return {"lm_iteration_ms": solve_ms + obj_ms + accepted_frac * form_ms,
            "regions_ms_scaled_to_window": {"objective": obj_ms, "form": form_ms, "solve": solve_ms},
            "measured_on": {"events": cpu["n_events"], "control_poses": sc.n_poses, "active_pixels": int(Np_p),
                            "objective_s": cpu["t_evaluate"], "form_s": cpu["t_form"], "solve_s": t_solve,
                            "solve_s_all_threads": t_solve_all, "threads_all": nthr},
            "host": host_info(),
            "extrapolated": True,
            "how": "objective and form scaled linearly in events, solve linearly in active pixels; iteration = solve + "
                   "objective + accepted_fraction * form (rejected steps reuse the equations, solver.cpp:66-130)"}


def cpu_full_lm(workload="C1", device="cuda"):
    """The reference's full LM loop (restated solveTimeWindow calling the reference's own LEGM methods) on a window it
    can finish, next to the device LM loop on the same window: a MEASURED LM-iteration ratio, nothing extrapolated."""
    from emba_b200 import synth
    from emba_b200.legm import Engine, spline_base_ns
    from oracle import ref_binding as RB

    if not RB.available():
        return None
    sc = synth.make_scene(**dict(synth.CONFIGS[workload]), device=device)
    RB, ref = _ref_model(sc, sc.n_events)
    tr = RB.RefTraj(sc.t_beg, sc.dt_knots, sc.quat_init)
    t = time.perf_counter()
    q_c, gx_c, gy_c, log_c, fc_c = ref.solve_time_window(tr, sc.Gx_init, sc.Gy_init, alpha=ALPHA, thres=THRES, **LM_KW)
    t_cpu = time.perf_counter() - t
    tm = ref.timers()
    eng = Engine(sc.sensor_w, sc.sensor_h, sc.bearing_lut(), sc.C_th, sc.pano_w, sc.pano_h)
    eng.set_events(sc.x, sc.y, sc.t_ns, sc.pol)
    t0, dt = spline_base_ns(sc.t_beg, sc.dt_knots)
    for _ in range(2):
        eng.set_state(0, t0, dt, sc.quat_init, sc.Gx_init, sc.Gy_init)
        eng.synchronize()
        t = time.perf_counter()
        log_g, fc_g = eng.solve_time_window(alpha=ALPHA, thres=THRES, **LM_KW)
        eng.synchronize()
        t_gpu = time.perf_counter() - t
    q_g, gx_g, gy_g = eng.get_state(0)
    eng.close()
    ang = 2 * np.arccos(np.abs(np.sum(q_g * q_c, -1)).clip(0, 1))
    return {"workload": workload, "events": int(sc.n_events), "control_poses": int(sc.n_poses),
            "cpu": {"solves": int(log_c.shape[0]), "accepted": int(log_c[:, 4].sum()), "total_s": t_cpu,
                    "lm_iteration_ms": t_cpu * 1e3 / max(1, log_c.shape[0]),
                    "regions_s": {"form": tm["form_s"], "solve": tm["solve_s"], "objective": tm["obj_s"]},
                    "cores": 1, "kind": "reference"},
            "gpu": {"solves": int(log_g.shape[0]), "accepted": int(log_g[:, 4].sum()), "total_ms": t_gpu * 1e3,
                    "lm_iteration_ms": t_gpu * 1e3 / max(1, log_g.shape[0]), "n_gpus": 1},
            "lm_iteration_ratio_cpu_over_gpu": (t_cpu / max(1, log_c.shape[0])) / (t_gpu / max(1, log_g.shape[0])),
            "parity": {"same_accept_reject_sequence": bool(log_c.shape[0] == log_g.shape[0] and
                                                            np.array_equal(log_c[:, 4], log_g[:, 4])),
                       "final_cost_rel": abs(fc_c - fc_g) / abs(fc_c), "max_rotation_error_rad": float(np.max(ang)),
                       "map_rel": rel(gx_c, gx_g)}}


def run_reference(args, workload, rank, world):
    """--impl reference: the reference's own CPU implementation of the pass, timed on the host (rank 0 only)."""
    if rank != 0:
        return
    from emba_b200 import synth

    kw = scene_kwargs(workload, args.gpus, args.weak)
    try:
        import torch
        dev = "cuda" if torch.cuda.is_available() else "cpu"
    except Exception:
        dev = "cpu"
    sc = synth.make_scene(**kw, device=dev, sort_on_device=False)  # the product library stays out of this process
    # bounded sample per step: a prefix of the window sized so that the whole run stays within a few minutes
    n_sample = min(sc.n_events, max(200_000, min(2_000_000, int(6e7 / max(1, args.steps + args.warmup)))))
    times, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_reference_pass(sc, n_sample)
        if i >= args.warmup:
            times.append(last["seconds"])
    sec = float(np.mean(times))
    n_used = last["n_events"]
    val = n_used / sec
    lm = None
    try:  # the three regions of an LM iteration on a larger prefix, scaled to the window (see cpu_lm_iteration_estimate)
        big = cpu_reference_pass(sc, min(sc.n_events, 4_000_000))
        Np_window = None
        # active pixels of the whole window are not known without a full pass: scale the prefix's count by events
        Np_window = big["ref"].Np * sc.n_events / big["n_events"] if big.get("ref") is not None else 0
        Np_window = min(Np_window, sc.pano_w * sc.pano_h)  # ... but never more than the panorama has
        lm = cpu_lm_iteration_estimate(sc, big, Np_window, 0.5)
        if lm is not None:
            lm["note"] = ("active pixels of the window estimated as prefix count x events ratio, capped at the panorama size "
                          "(an upper bound: revisits reuse pixels; the emba_b200 arm's cpu_baseline.lm_iteration_ms "
                          "uses the true count); accepted fraction 0.5 assumed")
    except Exception as ex:
        lm = {"error": str(ex)}
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "events/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3 * (sc.n_events / n_used), "higher_is_better": True,
        "scaling": "weak" if args.weak else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(sc, workload, world, args.weak),
        "cpu_baseline": {"value": val, "unit": "events/s", "cores": last["cores"], "kind": last["kind"], "host": host_info(),
                         "sample": f"first {n_used} events of the window per step, OMP_NUM_THREADS=1 "
                                   f"(the reference is single-threaded); ms_per_step is scaled linearly to the window"},
        "e2e": {"value": val, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "lm": lm,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="C4")
    ap.add_argument("--weak", action="store_true", help="grow the window with the GPU count (round-1 curve)")
    ap.add_argument("--impl", default="emba_b200", choices=["emba_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the atomic path / Poisson / adapter-protocol legs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = args.workload
    if args.impl == "reference":
        run_reference(args, workload, rank, world)
        return

    import ctypes as C

    import torch

    from emba_b200 import capi, synth
    from emba_b200.legm import Engine, EventPacket, LEGM, Trajectory, spline_base_ns

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- synthetic workload (seeded; every rank generates the same window, on its GPU)
    kw = scene_kwargs(workload, world, args.weak)
    t_gen = time.perf_counter()
    sc = synth.make_scene(**kw, device=f"cuda:{local_rank}")
    t_gen = time.perf_counter() - t_gen
    N = sc.n_events
    n = sc.n_poses
    P = sc.pano_w * sc.pano_h
    t0_ns, dt_ns = spline_base_ns(sc.t_beg, sc.dt_knots)
    # the caller's event buffers are page-locked (what a driver that wants the DMA path would do); state too
    ev_pin = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (sc.x, sc.y, sc.t_ns, sc.pol)]
    xp, yp, tp, pp_ = (t.numpy() for t in ev_pin)
    q_pin = torch.from_numpy(np.ascontiguousarray(sc.quat_init)).pin_memory()
    gx_pin = torch.from_numpy(np.ascontiguousarray(sc.Gx_init)).pin_memory()
    gy_pin = torch.from_numpy(np.ascontiguousarray(sc.Gy_init)).pin_memory()

    eng = Engine(sc.sensor_w, sc.sensor_h, sc.bearing_lut(), sc.C_th, sc.pano_w, sc.pano_h, device=local_rank)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.frombuffer(bytearray(eng.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        eng.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)

    def sync_all():
        eng.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def setup_window():
        eng.set_events(xp, yp, tp, pp_)
        eng.set_state(0, t0_ns, dt_ns, q_pin.numpy(), gx_pin.numpy(), gy_pin.numpy())

    sync_all()
    t_setup_cold = time.perf_counter()
    setup_window()  # first window: includes the one-off arena allocations
    sync_all()
    t_setup_cold = time.perf_counter() - t_setup_cold
    t_setup = time.perf_counter()
    setup_window()  # the same window again: what every further window of a sliding-window run costs
    sync_all()
    t_setup = time.perf_counter() - t_setup
    setup_dev_ms = eng.setup_ms()

    def one_pass():
        cd, cr, M = eng.evaluate(0, 0, 1.0, ALPHA)
        tm_e = eng.timings_ms()
        Np = eng.form_normal_eq(THRES, 0, 1.0, ALPHA)
        tm_f = eng.timings_ms()
        return cd + cr, M, Np, tm_e, tm_f

    # ---- device-resident pass: W warm-up + K timed steps
    clocks = ClockSampler(local_rank)
    clocks.start()  # started before the warm-up so that nvidia-smi is already streaming when the timed region begins
    for _ in range(args.warmup):
        one_pass()
    sync_all()
    clocks.mark()
    l0 = eng.launch_count()
    dev_ms, k_eval_ms, k_asm_ms, k_map_ms, ev_ms, form_ms, k_pix_ms, k_sort_ms = [], [], [], [], [], [], [], []
    wall = time.perf_counter()
    for _ in range(args.steps):
        cost, M, Np, tm_e, tm_f = one_pass()
        dev_ms.append(tm_e["evaluate"] + tm_f["form"])
        ev_ms.append(tm_e["evaluate"]); form_ms.append(tm_f["form"])
        k_eval_ms.append(tm_e["eval_kernel"]); k_asm_ms.append(tm_f["asm_pose_kernel"]); k_map_ms.append(tm_f["map_side"])
        k_pix_ms.append(tm_f["pix_kernel"]); k_sort_ms.append(tm_f["sort"])
    sync_all()
    wall = (time.perf_counter() - wall) / args.steps * 1e3
    launches = eng.launch_count() - l0
    clk = clocks.stop()
    cnt = eng.counters()
    comm = eng.comm_ms() if world > 1 else None
    ms_dev = float(np.mean(dev_ms))
    t_red = torch.tensor([ms_dev, wall], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_red, op=dist.ReduceOp.MAX)
    ms_step, wall_ms = float(t_red[0]), float(t_red[1])
    value = N / (ms_step * 1e-3)
    # per-rank kernel times and sizes (for the roofline block) gathered on rank 0
    mine = torch.tensor([float(np.mean(k_eval_ms)), float(np.mean(k_asm_ms)), float(np.mean(k_pix_ms)),
                         float(cnt["measurements_rank"]), float(cnt["a12_entries_local"]), float(cnt["a12_entries_solve"]),
                         float(np.mean(k_sort_ms)), float(np.mean(k_map_ms))], dtype=torch.float64, device="cuda")
    allr = [torch.zeros_like(mine) for _ in range(world)]
    if dist is not None:
        dist.all_gather(allr, mine)
    else:
        allr = [mine]
    allr = torch.stack(allr).cpu().numpy()

    # ---- per-kernel standalone durations: the same pass with the row placement / segment sort queued BEHIND the
    # pose-side kernel instead of beside it (EMBA_SIDE_SERIAL, csrc/assemble.cu), so that every kernel's event-timed
    # duration is its own; every rank runs the same three passes (the pass contains collectives)
    os.environ["EMBA_SIDE_SERIAL"] = "1"
    ser = []
    for i in range(4):
        _, _, _, tm_e, tm_f = one_pass()
        if i:
            ser.append([tm_e["eval_kernel"], tm_f["asm_pose_kernel"], tm_f["pix_kernel"], tm_f["sort"], tm_e["evaluate"] + tm_f["form"]])
    os.environ["EMBA_SIDE_SERIAL"] = "0"
    one_pass()
    sync_all()
    ser = np.mean(ser, 0)

    # ---- the fp64-atomic map-block path, reported beside the deterministic one (north star: "also reported")
    atomic = None
    if not args.no_extras:
        try:
            eng.set_map_path(1)
            one_pass()
            sync_all()
            ms_a, pix_a = [], []
            for _ in range(max(3, args.steps // 8)):
                _, _, _, tm_e, tm_f = one_pass()
                ms_a.append(tm_e["evaluate"] + tm_f["form"]); pix_a.append(tm_f["pix_kernel"])
            sync_all()
            atomic = {"ms_per_step": float(np.mean(ms_a)), "map_kernel_ms": float(np.mean(pix_a)),
                      "value": N / (float(np.mean(ms_a)) * 1e-3), "note": "same values up to summation order; not bit-reproducible"}
        except Exception as ex:
            atomic = {"error": str(ex)}
        eng.set_map_path(0)
        one_pass()

    # ---- end to end through the C ABI with host buffers: H2D state, pass, D2H (A11, b1, A22, b2); with several GPUs
    # only rank 0 receives the combined pose block (the call itself is collective)
    dp = C.POINTER(C.c_double)

    def pp(t):
        return C.cast(t.data_ptr(), dp)

    A11_h = torch.empty(9 * n * n, dtype=torch.float64).pin_memory()
    b1_h = torch.empty(3 * n, dtype=torch.float64).pin_memory()
    A22_h = torch.empty(4 * max(Np, 1), dtype=torch.float64).pin_memory()
    b2_h = torch.empty(2 * max(Np, 1), dtype=torch.float64).pin_memory()
    L = eng.L

    def one_e2e():
        rc = L.emba_set_state(eng.h, 0, t0_ns, dt_ns, n, pp(q_pin), pp(gx_pin), pp(gy_pin))
        assert rc == 0
        eng.evaluate(0, 0, 1.0, ALPHA)
        Np2 = eng.form_normal_eq(THRES, 0, 1.0, ALPHA)
        assert Np2 == Np
        if rank == 0:
            rc = L.emba_get_normal_eq(eng.h, pp(A11_h), pp(b1_h), pp(A22_h), pp(b2_h), None, None)
        else:
            rc = L.emba_get_normal_eq(eng.h, None, None, None, None, None, None)
        assert rc == 0

    for _ in range(max(1, args.warmup // 2)):
        one_e2e()
    sync_all()
    te = time.perf_counter()
    for _ in range(args.steps):
        one_e2e()
    sync_all()
    e2e_ms = (time.perf_counter() - te) / args.steps * 1e3
    t_red = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_red, op=dist.ReduceOp.MAX)
    e2e_ms = float(t_red[0])
    h2d = world * (32 * n + 16 * P)
    d2h = 8 * (9 * n * n + 3 * n + 4 * Np + 2 * Np)

    # ---- the same pass through the reference-shaped protocol an unchanged solver.cpp drives (N = 1): evaluateDataError
    # returns ep (M doubles, reference order) and num_ev_map, formNormalEq returns A11 / b1 / A22 / b2 / active set
    adapter = None
    if world == 1 and not args.no_extras:
        try:
            legm = LEGM.__new__(LEGM)
            legm.eng, legm._events_id, legm._pending, legm._M, legm._fix = eng, None, False, 0, 1
            evp = EventPacket(xp, yp, tp, pp_)
            legm._events_id = (id(evp.t_ns), evp.size())  # the window is already resident
            traj = Trajectory(sc.t_beg, sc.dt_knots, sc.quat_init)
            num_map = np.zeros((sc.pano_h, sc.pano_w), dtype=np.int32)

            def one_adapter():
                ep = legm.evaluateDataError(traj, sc.Gx_init, sc.Gy_init, evp, True, num_map)
                legm.formNormalEq(n, THRES, want_A12=False)
                legm.applyL2Reg(ALPHA)
                return ep.size

            one_adapter()
            sync_all()
            ta = time.perf_counter()
            reps = max(3, args.steps // 6)
            for _ in range(reps):
                Mad = one_adapter()
            sync_all()
            ad_ms = (time.perf_counter() - ta) / reps * 1e3
            adapter = {"ms_per_step": ad_ms, "value": N / (ad_ms * 1e-3), "unit": "events/s",
                       "d2h_bytes_per_step": int(8 * Mad + 4 * P + 8 * (9 * n * n + 3 * n + 6 * Np) + 8 * Np + 48 * Np),
                       "h2d_bytes_per_step": int(32 * n + 16 * P),
                       "note": "LEGM.evaluateDataError (ep + num_ev_map downloaded into pageable arrays) + formNormalEq + "
                               "applyL2Reg as solver.cpp:75-130 calls them; dense A12 is not transferred (the solve runs on the device)"}
            # leave the engine in the benchmark state
            eng.set_state(0, t0_ns, dt_ns, q_pin.numpy(), gx_pin.numpy(), gy_pin.numpy())
            one_pass()
        except Exception as ex:
            adapter = {"error": str(ex)}

    # ---- LM iteration: the full LM run of solveTimeWindow (solve + candidate + evaluate [+ assembly when accepted])
    lm = None
    try:
        for timed in (False, True):  # one untimed run first (first-use allocations of the solver's scratch)
            eng.set_state(0, t0_ns, dt_ns, q_pin.numpy(), gx_pin.numpy(), gy_pin.numpy())
            sync_all()
            tl = time.perf_counter()
            log, fcost = eng.solve_time_window(alpha=ALPHA, thres=THRES, **LM_KW)
            sync_all()
            tl = time.perf_counter() - tl
        t_red = torch.tensor([tl], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t_red, op=dist.ReduceOp.MAX)
        tl = float(t_red[0])
        nsolve = int(log.shape[0])
        nacc = int(log[:, 4].sum())
        lam_last = float(log[-1, 1]) * (0.1 if log[-1, 4] else 10.0)
        why = ("lambda > 1e3" if lam_last > 1e3 else "tol_fun satisfied" if nsolve <= LM_KW["max_num_iter"] else "max_num_iter")
        lm = {"solves": nsolve, "accepted": nacc, "rejected": nsolve - nacc, "ms_per_iteration": tl * 1e3 / max(1, nsolve),
              "total_ms": tl * 1e3, "terminated_by": why,
              "regions_device_ms_mean": {"form_when_accepted": float(log[log[:, 9] > 0, 9].mean()) if np.any(log[:, 9] > 0) else 0.0,
                                         "solve_and_update": float(log[:, 10].mean()), "evaluate": float(log[:, 11].mean())},
              "cost_first": float(log[0, 2]), "cost_last": float(fcost), "settings": dict(LM_KW, alpha=ALPHA, thres=THRES),
              "warmup_runs": 1}
    except Exception as ex:  # keep the headline even if the LM run fails
        lm = {"error": str(ex)}

    # ---- e2e of a whole LM window: event upload + pre-pass + state upload + LM run + state download
    lm_window = None
    try:
        sync_all()
        tw = time.perf_counter()
        setup_window()
        log_w, fc_w = eng.solve_time_window(alpha=ALPHA, thres=THRES, **LM_KW)
        q_out, gx_out, gy_out = eng.get_state(0)
        sync_all()
        tw = time.perf_counter() - tw
        t_red = torch.tensor([tw], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t_red, op=dist.ReduceOp.MAX)
        lm_window = {"ms": float(t_red[0]) * 1e3, "solves": int(log_w.shape[0]),
                     "h2d_bytes": int(5 * N + world * (16 * (N // 100) + 32 * n + 16 * P)), "d2h_bytes": int(world * (32 * n + 16 * P)),
                     "note": "emba_set_events (pinned x, y, polarity of the rank's own time slice + 2 timestamps per batch) + emba_set_state + "
                             "emba_solve_time_window + emba_get_state"}
    except Exception as ex:
        lm_window = {"error": str(ex)}

    # ---- multi-GPU parity, visible to the driver: the sharded pass against a single-GPU pass of the same window
    parity_mgpu = None
    if world > 1:
        try:
            eng.set_state(0, t0_ns, dt_ns, q_pin.numpy(), gx_pin.numpy(), gy_pin.numpy())
            cd, cr, Ms = eng.evaluate(0, 0, 1.0, ALPHA)
            Nps = eng.form_normal_eq(THRES, 0, 1.0, ALPHA)
            A11s, _, A22s, b1s, b2s, acts = eng.get_normal_eq(False)  # collective: combines the partial pose blocks
            x1s, x2s, _, _ = eng.solve(1e-3, False, True)
            if rank == 0:
                e1 = Engine(sc.sensor_w, sc.sensor_h, sc.bearing_lut(), sc.C_th, sc.pano_w, sc.pano_h, device=local_rank)
                e1.set_events(xp, yp, tp, pp_)
                e1.set_state(0, t0_ns, dt_ns, q_pin.numpy(), gx_pin.numpy(), gy_pin.numpy())
                cd1, cr1, M1 = e1.evaluate(0, 0, 1.0, ALPHA)
                Np1 = e1.form_normal_eq(THRES, 0, 1.0, ALPHA)
                A111, _, A221, b11, b21, act1 = e1.get_normal_eq(False)
                x11, x21, _, _ = e1.solve(1e-3, False, True)
                e1.close()
                parity_mgpu = {"against": "single-GPU pass of the same window on rank 0", "M_identical": bool(M1 == Ms),
                               "Np_identical": bool(Np1 == Nps), "active_set_identical": bool(np.array_equal(act1, acts)),
                               "cost_rel": abs((cd + cr) - (cd1 + cr1)) / abs(cd1 + cr1), "A11_rel": rel(A111, A11s),
                               "b1_rel": rel(b11, b1s), "A22_rel": rel(A221, A22s), "b2_rel": rel(b21, b2s),
                               "x1_rel": rel(x11, x1s), "x2_rel": rel(x21, x2s)}
                parity_mgpu["ok"] = bool(parity_mgpu["M_identical"] and parity_mgpu["Np_identical"] and
                                         parity_mgpu["active_set_identical"] and parity_mgpu["cost_rel"] < 1e-12 and
                                         max(parity_mgpu["A11_rel"], parity_mgpu["b1_rel"], parity_mgpu["A22_rel"],
                                             parity_mgpu["b2_rel"]) < 1e-9 and parity_mgpu["x1_rel"] < 1e-7)
        except Exception as ex:
            parity_mgpu = {"error": str(ex)}
        sync_all()

    # ---- "next" row N3: Poisson reconstruction of the intensity map from the device-resident gradient maps
    poisson = None
    if rank == 0 and not args.no_extras:
        try:
            eng.reconstruct_map(0)  # plan creation (sine matrices) + warm-up
            tp_ = []
            for _ in range(5):
                t0p = time.perf_counter(); eng.reconstruct_map(0); tp_.append((time.perf_counter() - t0p) * 1e3)
            flop = 4.0 * (sc.pano_h * sc.pano_w * (sc.pano_w + sc.pano_h))  # 2 x (S_H X S_W), 2 flop per MAC
            poisson = {"panorama": [sc.pano_w, sc.pano_h], "wall_ms_incl_d2h": float(np.median(tp_)),
                       "fp64_gflop": flop / 1e9,
                       "note": "DST-I as fp64 tensor-core GEMMs (csrc/poisson.cu); wall time includes the D2H copy of the image"}
        except Exception as ex:
            poisson = {"error": str(ex)}

    if rank == 0:
        peak, peak_src = measured_peak()
        # algorithmic bytes of ONE launch on ONE rank: per-measurement figure x that rank's rows
        kern_ranks = []
        for r in range(world):
            k_ev, k_as, k_px, Mr, nnz_loc, nnz_solve, k_so, k_mp = allr[r]
            Nr = N / world
            kern_ranks.append({
                "k_eval": (k_ev, 64.0 * Mr + 20.0 * P + 0.64 * Nr),
                "k_asm_pose": (k_as, 188.0 * Mr + 48.0 * P + 0.64 * Nr),
                "k_pix": (k_px, 132.0 * Mr + 8.0 * nnz_loc + 40.0 * Np)})
        # the slowest rank decides the step: report its kernels
        slow = int(np.argmax([sum(v[0] for v in kr.values()) for kr in kern_ranks]))
        kern = kern_ranks[slow]
        dom = max(kern, key=lambda k: kern[k][0])
        ach = kern[dom][1] / (kern[dom][0] * 1e-3) / 1e9
        traffic = None
        try:  # DRAM bytes per launch of the same kernel from the committed ncu --set full capture (profiles/)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            if world == 1 and workload in tj:
                traffic = tj[workload].get(dom)
        except Exception:
            traffic = None
        nnz12 = float(allr[:, 5].sum())  # every rank holds the merged strips of the pixels it owns
        pass_bytes = 28.0 * N + 44.0 * P + 40.0 * Np + 8.0 * nnz12 + 72.0 * n * n + 2.2 * N
        out = {
            "metric": METRIC, "value": value, "unit": "events/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": config_of(sc, workload, world, args.weak),
            "workload_stats": {"measurements": int(M), "active_pixels": int(Np), "a12_entries": int(nnz12),
                               "long_pixel_segments": cnt["long_segments"], "work_items_rank0": cnt["work_items"]},
            "e2e": {"value": N / (e2e_ms * 1e-3), "unit": "events/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms,
                    "window_setup_ms": t_setup * 1e3, "window_setup_first_ms": t_setup_cold * 1e3,
                    "window_setup_device_ms": {"set_events": setup_dev_ms[0], "static_rebuild": setup_dev_ms[1]},
                    "adapter_protocol": adapter, "lm_window": lm_window,
                    "note": "per step: pinned H2D of control poses + both maps (every rank), pass, D2H of A11,b1,A22,b2 "
                            "(rank 0); the event window is uploaded once per window (window_setup_ms: emba_set_events from "
                            "pinned arrays + first emba_set_state), as the reference passes it by const ref"},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": traffic, "peak_source": peak_src, "rank": slow,
                         "kernels_ms": dict({k: v[0] for k, v in kern.items()},
                                            place_and_segment_sort_side_stream=float(allr[slow, 6]),
                                            map_side_total=float(allr[slow, 7])),
                         "kernels_alg_gbs": {k: v[1] / (v[0] * 1e-3) / 1e9 for k, v in kern.items() if v[0] > 0},
                         "standalone_rank0": {
                             "note": "3 passes with the placement / segment sort queued behind the pose-side kernel "
                                     "instead of beside it: each kernel's own duration (same pass time)",
                             "kernels_ms": {"k_eval": float(ser[0]), "k_asm_pose": float(ser[1]), "k_pix": float(ser[2]),
                                            "place_and_segment_sort": float(ser[3])},
                             "ms_per_step": float(ser[4]),
                             "kernels_frac": {k: kern_ranks[0][k][1] / (t * 1e-3) / 1e9 / peak
                                              for k, t in (("k_eval", ser[0]), ("k_asm_pose", ser[1]), ("k_pix", ser[2])) if t > 0}},
                         "kernels_frac": {k: v[1] / (v[0] * 1e-3) / 1e9 / peak for k, v in kern.items() if v[0] > 0},
                         "pass_alg_bytes": pass_bytes,
                         "pass_frac": pass_bytes / (ms_step * 1e-3) / 1e9 / (peak * world),
                         "note": "achieved = algorithmic bytes of one launch on one rank / its launch time; pass_frac = "
                                 "SURVEY 8(d) bytes of the whole window / step time / (peak x GPUs)"},
            "breakdown_ms": {"evaluate": float(np.mean(ev_ms)), "form": float(np.mean(form_ms)), "wall_per_step": wall_ms,
                             "scene_generation_s": t_gen},
            "lm": lm,
            "map_path_atomic": atomic,
            "poisson_reconstruction": poisson,
        }
        if parity_mgpu is not None:
            out["parity_multi_gpu"] = parity_mgpu
        if comm is not None:
            out["comm_ms_rank0_last_step"] = comm
        if world == 1 and not args.no_cpu_baseline:
            # bounded sample: ~10-30 s of single-thread CPU work on a prefix of the same window; its outputs double as
            # a parity check of the device path on the same prefix
            n_cpu = min(N, 5_000_000)
            cpu = cpu_reference_pass(sc, n_cpu, want_outputs=True)
            out["cpu_baseline"] = {"value": cpu["value"], "unit": "events/s", "cores": cpu["cores"], "kind": cpu["kind"],
                                   "host": host_info(),
                                   "sample": f"one pass over the first {cpu['n_events']} events of the same window "
                                             f"({cpu['seconds']:.1f} s, OMP_NUM_THREADS=1: the reference is single-threaded)"}
            try:
                e2 = Engine(sc.sensor_w, sc.sensor_h, sc.bearing_lut(), sc.C_th, sc.pano_w, sc.pano_h, device=local_rank)
                ne = cpu["n_events"]
                e2.set_events(sc.x[:ne], sc.y[:ne], sc.t_ns[:ne], sc.pol[:ne])
                e2.set_state(0, t0_ns, dt_ns, sc.quat_init, sc.Gx_init, sc.Gy_init)
                cd2, cr2, M2 = e2.evaluate(0, 0, 1.0, ALPHA)
                ep2, num2 = e2.get_evaluation(0, M2)
                Np2 = e2.form_normal_eq(THRES, 0, 1.0, ALPHA)
                A11g, _, A22g, b1g, b2g, actg = e2.get_normal_eq(False)
                e2.close()
                par = {"against": f"{cpu['kind']} CPU pass on the first {ne} events of the window",
                       "M_identical": bool(M2 == cpu["ep"].size), "num_ev_map_identical": bool(np.array_equal(num2, cpu["num"])),
                       "Np_identical": bool(Np2 == cpu["act"].size), "active_set_identical": bool(np.array_equal(actg, cpu["act"])),
                       "ep_rel": rel(cpu["ep"], ep2), "cost_rel": abs((cd2 + cr2) - cpu["cost"]) / abs(cpu["cost"]),
                       "A11_rel": rel(cpu["A11"], A11g), "b1_rel": rel(cpu["b1"], b1g), "A22_rel": rel(cpu["A22"], A22g),
                       "b2_rel": rel(cpu["b2"], b2g)}
                par["ok"] = bool(par["M_identical"] and par["num_ev_map_identical"] and par["active_set_identical"] and
                                 par["ep_rel"] < 1e-6 and max(par["A11_rel"], par["b1_rel"], par["A22_rel"], par["b2_rel"]) < 1e-9)
                out["parity"] = par
            except Exception as ex:
                out["parity"] = {"error": str(ex)}
            try:
                acc_frac = (lm["accepted"] / max(1, lm["solves"])) if lm and "solves" in lm else 0.5
                est = cpu_lm_iteration_estimate(sc, cpu, Np, acc_frac)
                if est is not None:
                    out["cpu_baseline"].update(est)
                    if lm and "ms_per_iteration" in lm:
                        out["cpu_baseline"]["lm_iteration_ratio_cpu_over_gpu"] = est["lm_iteration_ms"] / lm["ms_per_iteration"]
            except Exception as ex:
                out["cpu_baseline"]["lm_iteration_error"] = str(ex)
            del cpu
            try:  # a window the CPU can finish: measured end to end, nothing extrapolated
                out["cpu_baseline"]["full_lm_run"] = cpu_full_lm("C1", f"cuda:{local_rank}")
            except Exception as ex:
                out["cpu_baseline"]["full_lm_run"] = {"error": str(ex)}
        print(json.dumps(out), flush=True)
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
