#!/usr/bin/env python
"""Benchmark of the EMBA LM hot path on B200 (contract: see the task statement / DESIGN.md section "Measurement").

A "step" is one PASS of the hot path over the whole event window of the workload, on a fixed state:
    LEGM::evaluateDataError(eval_deriv=true) + formNormalEq + applyL2Reg
    (residuals, num_ev_map, cost, per-measurement Jacobian rows, A11/A12/A22/b1/b2).
metric = events/s for that pass (whole job, all GPUs). The full LM iteration time is reported beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C1|C2|C3|C4|small|tiny] [--impl reference]

N = 1: workload C2 (BASELINE.json configs[1]: 240x180 sensor, 1024x512 panorama, ~10M synthetic events, n=97).
N > 1: launched by torchrun, one rank per GPU; the window is time-sharded (contiguous control-pose slices) and the
partial (H, g) are combined with NCCL all-reduce inside the library. Weak scaling: the window grows with N
(N x the C2 time span, ~N x 10M events) so per-GPU work stays fixed.
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


def scene_kwargs(workload, n_gpus):
    from emba_b200 import synth

    kw = dict(synth.CONFIGS[workload])
    if n_gpus > 1:
        span = kw["t_end"] - kw["t_beg"]
        kw["t_end"] = kw["t_beg"] + span * n_gpus
        kw["periodic"] = True  # bounded yaw for long spans (keeps the view away from the seam)
    return kw


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


def cpu_reference_pass(sc, n_events, repeats=1):
    """Reference CPU implementation (oracle/_ref, the unmodified reference sources) of one pass on the first
    n_events events. Returns (events/s, kind, cores, seconds)."""
    from oracle import ref_binding as RB

    n_events = (min(n_events, sc.n_events) // 100) * 100
    if RB.available():
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        ref = RB.RefLEGM(sc.sensor_w, sc.sensor_h, sc.fx, sc.fy, sc.cx, sc.cy, sc.C_th, sc.pano_w, sc.pano_h)
        ref.set_events(sc.x[:n_events], sc.y[:n_events], sc.t_ns[:n_events], sc.pol[:n_events])
        tr = RB.RefTraj(sc.t_beg, sc.dt_knots, sc.quat_init)
        best = None
        for _ in range(repeats):
            t = time.perf_counter()
            ref.evaluate(tr, sc.Gx_init, sc.Gy_init, True)
            ref.form(sc.n_poses, THRES, sc.Gx_init, sc.Gy_init, ALPHA, want_A12=False)
            dt = time.perf_counter() - t
            best = dt if best is None else min(best, dt)
        return n_events / best, "reference", 1, best
    from oracle import emba_oracle as O

    orc = O.Oracle(sc.sensor_w, sc.sensor_h, sc.bearing_lut(), sc.pano_w, sc.pano_h, sc.C_th)
    n_events = min(n_events, 200_000)
    orc.set_events(sc.x[:n_events], sc.y[:n_events], sc.t_ns[:n_events], sc.pol[:n_events])
    t0, dtn = O.spline_base_ns(sc.t_beg, sc.dt_knots)
    t = time.perf_counter()
    orc.evaluate(sc.quat_init, t0, dtn, sc.Gx_init, sc.Gy_init, True)
    A = orc.form_normal_eq(sc.n_poses, THRES)
    orc.apply_l2_reg(A[2], A[4], A[5], ALPHA, sc.Gx_init, sc.Gy_init)
    dt = time.perf_counter() - t
    return n_events / dt, "port", 1, dt


def run_reference(args, workload, rank, world):
    """--impl reference: the reference's own CPU implementation of the pass, timed on the host (rank 0 only)."""
    if rank != 0:
        return
    from emba_b200 import synth

    kw = scene_kwargs(workload, args.gpus)
    try:
        import torch
        dev = "cuda" if torch.cuda.is_available() else "cpu"
    except Exception:
        dev = "cpu"
    sc = synth.make_scene(**kw, device=dev)
    # bounded sample per step: a prefix of the window sized for ~3 s of CPU work
    n_sample = min(sc.n_events, max(200_000, min(2_000_000, int(1.4e8 / max(1, args.steps + args.warmup)))))
    times = []
    for i in range(args.warmup + args.steps):
        v, kind, cores, sec = cpu_reference_pass(sc, n_sample)
        if i >= args.warmup:
            times.append(sec)
    sec = float(np.mean(times))
    n_used = (n_sample // 100) * 100
    val = n_used / sec
    out = {
        "impl": "reference", "metric": "events/s for residual+Jacobian+H assembly", "value": val, "unit": "events/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3 * (sc.n_events / n_used), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "events": sc.n_events, "sensor": [sc.sensor_w, sc.sensor_h],
                   "panorama": [sc.pano_w, sc.pano_h], "control_poses": sc.n_poses},
        "cpu_baseline": {"value": val, "unit": "events/s", "cores": cores, "kind": kind,
                         "sample": f"first {n_used} events of the window per step, OMP_NUM_THREADS=1 "
                                   f"(the reference is single-threaded); ms_per_step is scaled linearly to the window"},
        "e2e": {"value": val, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--impl", default="emba_b200", choices=["emba_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lm-iters", type=int, default=6)
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
    from emba_b200.legm import Engine, spline_base_ns

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- synthetic workload (seeded; every rank generates the same window, on its GPU)
    kw = scene_kwargs(workload, world)
    t_gen = time.perf_counter()
    sc = synth.make_scene(**kw, device=f"cuda:{local_rank}")
    t_gen = time.perf_counter() - t_gen
    N = sc.n_events
    n = sc.n_poses
    P = sc.pano_w * sc.pano_h
    t0_ns, dt_ns = spline_base_ns(sc.t_beg, sc.dt_knots)

    eng = Engine(sc.sensor_w, sc.sensor_h, sc.bearing_lut(), sc.C_th, sc.pano_w, sc.pano_h, device=local_rank)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.frombuffer(bytearray(eng.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        eng.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
    t_setup = time.perf_counter()
    eng.set_events(sc.x, sc.y, sc.t_ns, sc.pol)
    # pinned host buffers for the state (the step's inputs)
    q_pin = torch.from_numpy(np.ascontiguousarray(sc.quat_init)).pin_memory()
    gx_pin = torch.from_numpy(np.ascontiguousarray(sc.Gx_init)).pin_memory()
    gy_pin = torch.from_numpy(np.ascontiguousarray(sc.Gy_init)).pin_memory()
    eng.set_state(0, t0_ns, dt_ns, q_pin.numpy(), gx_pin.numpy(), gy_pin.numpy())
    t_setup = time.perf_counter() - t_setup

    def sync_all():
        eng.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

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
    nnz12_local = eng.a12_entries()
    ms_dev = float(np.mean(dev_ms))
    t_red = torch.tensor([ms_dev, wall], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_red, op=dist.ReduceOp.MAX)
    ms_step, wall_ms = float(t_red[0]), float(t_red[1])
    value = N / (ms_step * 1e-3)

    # ---- the fp64-atomic map-block path, reported beside the deterministic one (north star: "also reported")
    atomic = None
    try:
        eng.set_map_path(1)
        for _ in range(2):
            one_pass()
        sync_all()
        ms_a, pix_a = [], []
        for _ in range(max(3, args.steps // 5)):
            _, _, _, tm_e, tm_f = one_pass()
            ms_a.append(tm_e["evaluate"] + tm_f["form"]); pix_a.append(tm_f["pix_kernel"])
        sync_all()
        atomic = {"ms_per_step": float(np.mean(ms_a)), "map_kernel_ms": float(np.mean(pix_a)),
                  "value": N / (float(np.mean(ms_a)) * 1e-3), "note": "same values up to summation order; not bit-reproducible"}
    except Exception as ex:
        atomic = {"error": str(ex)}
    eng.set_map_path(0)
    one_pass()

    # ---- end to end through the C ABI with host buffers: H2D state, pass, D2H (cost, A11, b1, A22, b2)
    A11_h = torch.empty(9 * n * n, dtype=torch.float64).pin_memory()
    b1_h = torch.empty(3 * n, dtype=torch.float64).pin_memory()
    A22_h = torch.empty(4 * max(Np, 1), dtype=torch.float64).pin_memory()
    b2_h = torch.empty(2 * max(Np, 1), dtype=torch.float64).pin_memory()
    L = eng.L
    dp = C.POINTER(C.c_double)

    def pp(t):
        return C.cast(t.data_ptr(), dp)

    def one_e2e():
        rc = L.emba_set_state(eng.h, 0, t0_ns, dt_ns, n, pp(q_pin), pp(gx_pin), pp(gy_pin))
        assert rc == 0
        eng.evaluate(0, 0, 1.0, ALPHA)
        Np2 = eng.form_normal_eq(THRES, 0, 1.0, ALPHA)
        assert Np2 == Np
        rc = L.emba_get_normal_eq(eng.h, pp(A11_h), pp(b1_h), pp(A22_h), pp(b2_h), None, None)
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
    h2d = 32 * n + 16 * P
    d2h = 8 * (9 * n * n + 3 * n + 4 * Np + 2 * Np) + 24

    # ---- LM iteration time (solve + candidate + evaluate [+ assembly when accepted]), short run
    lm = None
    try:
        # one untimed run first (first-use allocations of the solver's scratch), then the same run timed
        for timed in (False, True):
            eng.set_state(0, t0_ns, dt_ns, sc.quat_init, sc.Gx_init, sc.Gy_init)
            sync_all()
            tl = time.perf_counter()
            log, fcost = eng.solve_time_window(max_num_iter=args.lm_iters - 1, alpha=ALPHA, thres=THRES)
            sync_all()
            tl = time.perf_counter() - tl
        lm = {"iterations": int(log.shape[0]), "accepted": int(log[:, 4].sum()), "ms_per_iteration": tl * 1e3 / max(1, log.shape[0]),
              "cost_first": float(log[0, 2]), "cost_last": float(fcost), "warmup_runs": 1}
    except Exception as ex:  # keep the headline even if the short LM run fails
        lm = {"error": str(ex)}

    # ---- "next" row N3: Poisson reconstruction of the intensity map from the device-resident gradient maps
    poisson = None
    if rank == 0:
        try:
            eng.reconstruct_map(0)  # plan creation (sine matrices) + warm-up
            tp = []
            for _ in range(5):
                t0p = time.perf_counter(); eng.reconstruct_map(0); tp.append((time.perf_counter() - t0p) * 1e3)
            flop = 4.0 * (sc.pano_h * sc.pano_w * (sc.pano_w + sc.pano_h))  # 2 x (S_H X S_W), 2 flop per MAC
            poisson = {"panorama": [sc.pano_w, sc.pano_h], "wall_ms_incl_d2h": float(np.median(tp)),
                       "fp64_gflop": flop / 1e9,
                       "note": "DST-I as fp64 tensor-core GEMMs (csrc/poisson.cu); wall time includes the D2H copy of the image"}
        except Exception as ex:
            poisson = {"error": str(ex)}

    if rank == 0:
        peak, peak_src = measured_peak()
        nnz12 = nnz12_local * world  # every rank owns 1/world of the pixels
        M_all = M
        # algorithmic bytes: per-kernel figure x the units one launch processes on one rank
        b_eval = 60.0 * eng.num_pairs() / world + 20.0 * P + 0.64 * N
        b_asm = 188.0 * M_all / world + 48.0 * P + 0.64 * N
        b_pix = 132.0 * M_all / world + 8.0 * nnz12 + 40.0 * Np
        kern = {"k_eval": (float(np.mean(k_eval_ms)), b_eval), "k_asm_pose": (float(np.mean(k_asm_ms)), b_asm),
                "k_pix": (float(np.mean(k_pix_ms)), b_pix)}
        dom = max(kern, key=lambda k: kern[k][0])
        ach = kern[dom][1] / (kern[dom][0] * 1e-3) / 1e9
        traffic = None
        try:  # DRAM bytes per launch of the same kernel from the committed ncu --set full capture (profiles/)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if world == 1 and workload in tj:
                traffic = tj[workload].get(dom)
        except Exception:
            traffic = None
        pass_bytes = 28.0 * N + 44.0 * P + 40.0 * Np + 8.0 * nnz12 + 72.0 * n * n + 2.2 * N
        out = {
            "metric": "events/s for residual+Jacobian+H assembly", "value": value, "unit": "events/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload if world == 1 else f"{workload} x{world} time span", "events": N,
                       "measurements": int(M_all), "sensor": [sc.sensor_w, sc.sensor_h],
                       "panorama": [sc.pano_w, sc.pano_h], "control_poses": n, "active_pixels": int(Np),
                       "a12_entries": int(nnz12), "l2": "inputs_exceed_l2" if 44 * M_all > 126e6 else "inputs_fit_l2",
                       "parallelism": f"time-sharded x{world}" if world > 1 else "single GPU"},
            "e2e": {"value": N / (e2e_ms * 1e-3), "unit": "events/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms,
                    "window_setup_ms": t_setup * 1e3,
                    "note": "per step: pinned H2D of control poses + both maps, pass, D2H of A11,b1,A22,b2; the event "
                            "window is uploaded once per window (window_setup_ms), as the reference passes it by const ref"},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernels_ms": dict({k: v[0] for k, v in kern.items()}, row_sort_side_stream=float(np.mean(k_sort_ms)),
                                            map_side_total=float(np.mean(k_map_ms))),
                         "kernels_alg_gbs": {k: v[1] / (v[0] * 1e-3) / 1e9 for k, v in kern.items() if v[0] > 0},
                         "pass_alg_bytes": pass_bytes, "pass_frac": pass_bytes / (ms_step * 1e-3) / 1e9 / peak},
            "breakdown_ms": {"evaluate": float(np.mean(ev_ms)), "form": float(np.mean(form_ms)), "wall_per_step": wall_ms,
                             "scene_generation_s": t_gen},
            "lm": lm,
            "map_path_atomic": atomic,
            "poisson_reconstruction": poisson,
        }
        if world == 1 and not args.no_cpu_baseline:
            # bounded sample: ~10-30 s of single-thread CPU work
            n_cpu = min(N, 10_000_000)
            v, kind, cores, sec = cpu_reference_pass(sc, n_cpu)
            out["cpu_baseline"] = {"value": v, "unit": "events/s", "cores": cores, "kind": kind,
                                   "sample": f"one pass over the first {(n_cpu // 100) * 100} events of the same window "
                                             f"({sec:.1f} s, OMP_NUM_THREADS=1: the reference is single-threaded)"}
        print(json.dumps(out), flush=True)
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
