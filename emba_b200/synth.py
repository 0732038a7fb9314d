"""Seeded synthetic rotational event data for the EMBA hot path.

Panorama log-intensity texture -> ground-truth gradient maps; smooth SO(3) ground-truth
trajectory; per-sensor-pixel threshold-crossing event simulator; perturbed initial control
poses and initial map. Everything the benchmark and the parity tests feed through the C ABI
comes from here (there is no network for datasets, and no ROS for rosbags).

The simulator is written with torch tensor ops only so it runs on the CPU (tests, here) and
on the GPU (bench.py, large N). torch is plumbing for data generation; it is not on the
measured path.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np
import torch


@dataclasses.dataclass
class Scene:
    # sensor
    sensor_w: int
    sensor_h: int
    fx: float
    fy: float
    cx: float
    cy: float
    # panorama
    pano_w: int
    pano_h: int
    # model
    C_th: float
    # spline
    t_beg: float
    dt_knots: float
    n_poses: int
    # data (numpy)
    x: np.ndarray = None  # uint16 [N]
    y: np.ndarray = None  # uint16 [N]
    t_ns: np.ndarray = None  # int64 [N]
    pol: np.ndarray = None  # uint8 [N]
    quat_gt: np.ndarray = None  # [n,4] xyzw
    quat_init: np.ndarray = None  # [n,4] xyzw
    Gx_gt: np.ndarray = None  # [H,W]
    Gy_gt: np.ndarray = None
    Gx_init: np.ndarray = None
    Gy_init: np.ndarray = None

    @property
    def n_events(self):
        return int(self.t_ns.size)

    def bearing_lut(self):
        """Host-side bearing LUT for a zero-distortion pinhole camera, as the reference's
        EventWarper::precomputeBearingVectors (src/utils/event_pano_warper.cpp:27-41) builds
        it with image_geometry: ((x-cx)/fx, (y-cy)/fy, 1), row-major over the sensor."""
        xs = np.arange(self.sensor_w, dtype=np.float64)
        ys = np.arange(self.sensor_h, dtype=np.float64)
        X, Y = np.meshgrid(xs, ys)
        return np.stack([(X - self.cx) / self.fx, (Y - self.cy) / self.fy, np.ones_like(X)], -1).reshape(-1, 3)


def _sobel8(L):
    """0.125 * 3x3 Sobel with reflect-101 borders (same operator the reference applies to the
    gradient maps, src/emba/model.cpp:87-97); returns d/dx, d/dy."""
    P = np.pad(L, 1, mode="reflect")
    rx = P[:, 2:] - P[:, :-2]
    gx = rx[:-2, :] + 2.0 * rx[1:-1, :] + rx[2:, :]
    ry = P[:, :-2] + 2.0 * P[:, 1:-1] + P[:, 2:]
    gy = ry[2:, :] - ry[:-2, :]
    return 0.125 * gx, 0.125 * gy


def make_texture(pano_w, pano_h, seed=1, std=1.5):
    """Gaussian-filtered white noise (sigma = 4 px at 1024x512, scaled with resolution),
    periodic in x, normalised to the given std."""
    rng = np.random.default_rng(seed)
    noise = rng.standard_normal((pano_h, pano_w))
    sigma = 4.0 * pano_w / 1024.0
    fy = np.fft.fftfreq(pano_h)[:, None]
    fx = np.fft.fftfreq(pano_w)[None, :]
    filt = np.exp(-2.0 * (np.pi * sigma) ** 2 * (fx * fx + fy * fy))
    L = np.real(np.fft.ifft2(np.fft.fft2(noise) * filt))
    L = (L - L.mean()) / L.std() * std
    return L


YAW0 = 0.0  # yaw offset of the ground-truth trajectory (make_scene sets it; pi-ish values look at the panorama seam)


def gt_rotvec(t, yaw_rate=0.35, periodic=False):
    """Ground-truth rotation vector phi(t) = (0.25 sin 1.3t, 0.35 t, 0.08 cos 0.7t) rad; with
    periodic=True the yaw is a bounded triangle wave with the same rate (for spans > 10 s)."""
    t = np.asarray(t, dtype=np.float64)
    if YAW0 != 0.0:
        return np.stack([0.25 * np.sin(1.3 * t), YAW0 + yaw_rate * t, 0.08 * np.cos(0.7 * t)], -1)
    if periodic:
        # triangle wave: constant |yaw rate|, bounded to +-1.6 rad (keeps the view away from the seam for any span)
        A = 1.6
        ph = (yaw_rate * t + A) % (4 * A)
        yaw = np.where(ph < 2 * A, ph - A, 3 * A - ph)
    else:
        yaw = yaw_rate * t
    return np.stack([0.25 * np.sin(1.3 * t), yaw, 0.08 * np.cos(0.7 * t)], -1)


def _exp_so3_torch(phi):
    """Rodrigues, batched: phi (...,3) -> R (...,3,3)."""
    th = torch.linalg.norm(phi, dim=-1, keepdim=True).clamp_min(1e-300)
    k = phi / th
    K = torch.zeros(phi.shape[:-1] + (3, 3), dtype=phi.dtype, device=phi.device)
    K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
    K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
    s = torch.sin(th)[..., None]
    c = torch.cos(th)[..., None]
    I = torch.eye(3, dtype=phi.dtype, device=phi.device).expand(K.shape)
    return I + s * K + (1 - c) * (K @ K)


def _rotvec_to_quat(phi):
    phi = np.asarray(phi, dtype=np.float64)
    th = np.linalg.norm(phi, axis=-1, keepdims=True)
    half = 0.5 * th
    with np.errstate(invalid="ignore", divide="ignore"):
        f = np.where(th > 1e-12, np.sin(half) / np.where(th > 1e-12, th, 1.0), 0.5)
    return np.concatenate([f * phi, np.cos(half)], -1)


def _quat_mul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz], -1)


def simulate_events(L, scene: Scene, t_start, t_stop, *, dt_sim=0.25e-3, yaw_rate=0.35, periodic=False,
                    device="cpu", max_events=None):
    """Per-pixel threshold-crossing simulator. At every simulation step the log intensity seen
    by each sensor pixel (bilinear sample of L at the warped location) is compared with the
    pixel's reference level; every crossing of a multiple of C_th emits an event whose
    timestamp is linearly interpolated inside the step. Returns time-sorted numpy arrays."""
    dev = torch.device(device)
    f64 = torch.float64
    W, H = scene.pano_w, scene.pano_h
    Lt = torch.as_tensor(L, dtype=f64, device=dev)
    lut = torch.as_tensor(scene.bearing_lut(), dtype=f64, device=dev)  # [S,3]
    S = lut.shape[0]
    fxp, fyp = W / (2 * math.pi), H / math.pi
    C = scene.C_th
    n_steps = int(round((t_stop - t_start) / dt_sim))
    ts = t_start + dt_sim * np.arange(n_steps + 1)
    phis = torch.as_tensor(gt_rotvec(ts, yaw_rate, periodic), dtype=f64, device=dev)
    Rs = _exp_so3_torch(phis)  # [T,3,3]

    def sample(R):
        rb = lut @ R.T
        X, Y, Z = rb[:, 0], rb[:, 1], rb[:, 2]
        px = W / 2 + fxp * torch.atan2(X, Z)
        py = H / 2 + fyp * torch.asin(Y / torch.linalg.norm(rb, dim=1))
        x0 = torch.floor(px)
        y0 = torch.floor(py)
        ax, ay = px - x0, py - y0
        x0i = x0.long() % W
        x1i = (x0i + 1) % W
        y0i = y0.long().clamp(0, H - 1)
        y1i = (y0i + 1).clamp(0, H - 1)
        return ((1 - ax) * (1 - ay) * Lt[y0i, x0i] + ax * (1 - ay) * Lt[y0i, x1i]
                + (1 - ax) * ay * Lt[y1i, x0i] + ax * ay * Lt[y1i, x1i])

    pix = torch.arange(S, device=dev)
    L_prev = sample(Rs[0])
    ref = L_prev.clone()
    out_t, out_p, out_s = [], [], []
    total = 0
    for k in range(1, n_steps + 1):
        Lk = sample(Rs[k])
        d = Lk - ref
        ncross = torch.floor(d.abs() / C)
        nmax = int(ncross.max().item()) if ncross.numel() else 0
        if nmax > 0:
            sgn = torch.sign(d)
            dL = Lk - L_prev
            for j in range(1, nmax + 1):
                m = ncross >= j
                lvl = ref[m] + sgn[m] * (j * C)
                frac = ((lvl - L_prev[m]) / dL[m]).clamp(0.0, 1.0)
                out_t.append(ts[k - 1] + frac * dt_sim)
                out_p.append((sgn[m] > 0).to(torch.uint8))
                out_s.append(pix[m])
                total += int(m.sum().item())
            ref = ref + sgn * ncross * C
        L_prev = Lk
        if max_events is not None and total >= max_events:
            break
    if not out_t:
        z = np.zeros(0)
        return z.astype(np.uint16), z.astype(np.uint16), z.astype(np.int64), z.astype(np.uint8)
    t = torch.cat(out_t)
    p = torch.cat(out_p)
    s = torch.cat(out_s)
    t_ns = torch.round(t * 1e9).to(torch.int64)
    order = torch.argsort(t_ns, stable=True)
    t_ns, p, s = t_ns[order], p[order], s[order]
    x = (s % scene.sensor_w).to(torch.int32)
    y = (s // scene.sensor_w).to(torch.int32)
    return (x.cpu().numpy().astype(np.uint16), y.cpu().numpy().astype(np.uint16), t_ns.cpu().numpy(),
            p.cpu().numpy().astype(np.uint8))


def _exp_so3_np(phi):
    th = np.linalg.norm(phi, axis=-1, keepdims=True).clip(1e-300)
    k = phi / th
    K = np.zeros(phi.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
    K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
    s = np.sin(th)[..., None]
    c = np.cos(th)[..., None]
    return np.eye(3) + s * K + (1 - c) * (K @ K)


def simulate_events_cuda(L, scene: Scene, t_start, t_stop, *, dt_sim=0.25e-3, yaw_rate=0.35, periodic=False,
                         device_index=0, sort_on_device=True):
    """Same event model as simulate_events, one CUDA thread per sensor pixel (emba_b200/csrc/synth.cu through
    include/emba_synth.h). Used for the large benchmark configurations."""
    import ctypes as C
    import os

    lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libemba_synth.so"))
    dp = C.POINTER(C.c_double)
    lib.emba_synth_simulate.argtypes = [C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_int, dp, C.c_double, C.c_int, dp,
                                        C.c_double, C.c_double, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    lib.emba_synth_fetch.argtypes = [C.c_void_p, C.POINTER(C.c_uint16), C.POINTER(C.c_uint16), C.POINTER(C.c_int64),
                                     C.POINTER(C.c_uint8)]
    lib.emba_synth_free.argtypes = [C.c_void_p]
    n_steps = int(round((t_stop - t_start) / dt_sim))
    ts = t_start + dt_sim * np.arange(n_steps + 1)
    Rs = np.ascontiguousarray(_exp_so3_np(gt_rotvec(ts, yaw_rate, periodic)).reshape(-1, 9))
    lut = np.ascontiguousarray(scene.bearing_lut())
    Lc = np.ascontiguousarray(L, dtype=np.float64)
    h = C.c_void_p()
    n = C.c_int64(0)
    rc = lib.emba_synth_simulate(device_index, scene.sensor_w, scene.sensor_h, lut.ctypes.data_as(dp), scene.pano_w,
                                 scene.pano_h, Lc.ctypes.data_as(dp), float(scene.C_th), n_steps,
                                 Rs.ctypes.data_as(dp), float(t_start), float(dt_sim), C.byref(h), C.byref(n))
    if rc != 0:
        raise RuntimeError(f"emba_synth_simulate failed ({rc})")
    N = n.value
    x = np.empty(N, dtype=np.uint16)
    y = np.empty(N, dtype=np.uint16)
    t = np.empty(N, dtype=np.int64)
    p = np.empty(N, dtype=np.uint8)
    rc = lib.emba_synth_fetch(h, x.ctypes.data_as(C.POINTER(C.c_uint16)), y.ctypes.data_as(C.POINTER(C.c_uint16)),
                              t.ctypes.data_as(C.POINTER(C.c_int64)), p.ctypes.data_as(C.POINTER(C.c_uint8)))
    lib.emba_synth_free(h)
    if rc != 0:
        raise RuntimeError(f"emba_synth_fetch failed ({rc})")
    # the simulator emits sensor-pixel-major streams: sort by time with the library's device-resident event sequence
    # (stable, so equal stamps keep the simulator's pixel order), like the reference sorts a parsed bag
    if not sort_on_device:  # bench.py's reference arm: nothing of the product library runs in that process
        order = np.argsort(t, kind="stable")
        return x[order], y[order], t[order], p[order]
    from .legm import EventSequence

    seq = EventSequence(x, y, t, p, device=device_index)
    seq.sort_by_time()
    x, y, t, p = seq.download()
    seq.close()
    return x, y, t, p


def make_scene(sensor_w=128, sensor_h=128, fx=91.4014729896821, fy=None, cx=None, cy=None, pano_w=1024, pano_h=512,
               C_th=0.45, t_beg=0.1, t_end=2.4, dt_knots=0.05, texture_std=1.5, seed=1, dt_sim=0.25e-3,
               yaw_rate=0.35, periodic=False, device="cpu", pose_noise_deg=0.3, map_scale=0.7, map_noise=0.02,
               max_events=None, guard=1e-3, yaw0=0.0, sort_on_device=True) -> Scene:
    """Build a full seeded problem instance. Defaults are config C1 of SURVEY.md section 8(d)
    (calib/DVS-playroom.yaml intrinsics, launch/playroom.launch parameters)."""
    global YAW0
    YAW0 = float(yaw0)
    fy = fx if fy is None else fy
    cx = sensor_w / 2.0 if cx is None else cx
    cy = sensor_h / 2.0 if cy is None else cy
    n_poses = int(round((t_end - t_beg) / dt_knots)) + 1
    sc = Scene(sensor_w, sensor_h, fx, fy, cx, cy, pano_w, pano_h, C_th, t_beg, dt_knots, n_poses)
    L = make_texture(pano_w, pano_h, seed=seed, std=texture_std)
    sc.Gx_gt, sc.Gy_gt = _sobel8(L)
    # events strictly inside the spline support, like EMBA::getEventSubset's 1 ms guard
    # (src/emba/emba.cpp:476-478)
    if str(device).startswith("cuda"):
        dev_idx = int(str(device).split(":")[1]) if ":" in str(device) else 0
        x, y, t_ns, pol = simulate_events_cuda(L, sc, t_beg + guard, t_end - guard, dt_sim=dt_sim, yaw_rate=yaw_rate,
                                               periodic=periodic, device_index=dev_idx, sort_on_device=sort_on_device)
    else:
        x, y, t_ns, pol = simulate_events(L, sc, t_beg + guard, t_end - guard, dt_sim=dt_sim, yaw_rate=yaw_rate,
                                          periodic=periodic, device=device, max_events=max_events)
    if max_events is not None and t_ns.size > max_events:
        x, y, t_ns, pol = x[:max_events], y[:max_events], t_ns[:max_events], pol[:max_events]
    n_keep = (t_ns.size // 100) * 100  # the reference drops the tail batch (model.cpp:78-79)
    sc.x, sc.y, sc.t_ns, sc.pol = x[:n_keep], y[:n_keep], t_ns[:n_keep], pol[:n_keep]
    tk = t_beg + dt_knots * np.arange(n_poses)
    sc.quat_gt = _rotvec_to_quat(gt_rotvec(tk, yaw_rate, periodic))
    rng = np.random.default_rng(seed)
    dphi = rng.standard_normal((n_poses, 3)) * np.deg2rad(pose_noise_deg)
    dphi[0] = 0.0
    q = _quat_mul(_rotvec_to_quat(dphi), sc.quat_gt)
    sc.quat_init = q / np.linalg.norm(q, axis=-1, keepdims=True)
    rng7 = np.random.default_rng(seed + 6)
    sc.Gx_init = map_scale * sc.Gx_gt + map_noise * rng7.standard_normal(sc.Gx_gt.shape)
    sc.Gy_init = map_scale * sc.Gy_gt + map_noise * rng7.standard_normal(sc.Gy_gt.shape)
    YAW0 = 0.0
    return sc


# Named configurations; texture_std tunes the event count.
CONFIGS = {
    # tiny: CPU-second test case
    "tiny": dict(sensor_w=32, sensor_h=24, fx=30.0, pano_w=256, pano_h=128, C_th=0.3, t_beg=0.1, t_end=0.6,
                 dt_knots=0.05, texture_std=1.5, dt_sim=0.5e-3),
    # seam: the tiny scene looking at the phi = +-pi seam of the panorama (warped events round to column pano_w)
    "seam": dict(sensor_w=32, sensor_h=24, fx=30.0, pano_w=256, pano_h=128, C_th=0.3, t_beg=0.1, t_end=0.6,
                 dt_knots=0.05, texture_std=1.5, dt_sim=0.5e-3, yaw0=3.0),
    # small: a few 10k events
    "small": dict(sensor_w=64, sensor_h=48, fx=60.0, pano_w=512, pano_h=256, C_th=0.3, t_beg=0.1, t_end=1.1,
                  dt_knots=0.05, texture_std=1.5, dt_sim=0.5e-3),
    # C1: playroom.launch (128x128 DVS-playroom.yaml, 1024x512, n=47, C_th=0.45), ~1M events
    "C1": dict(sensor_w=128, sensor_h=128, fx=91.4014729896821, pano_w=1024, pano_h=512, C_th=0.45, t_beg=0.1,
               t_end=2.4, dt_knots=0.05, texture_std=1.5),
    # C2: bay.launch-style 240x180, f=200, 1024x512, t in [0.1,4.9], n=97, C_th=0.2, ~10M events
    "C2": dict(sensor_w=240, sensor_h=180, fx=200.0, pano_w=1024, pano_h=512, C_th=0.2, t_beg=0.1, t_end=4.9,
               dt_knots=0.05, texture_std=1.25),
    # C3: shapes.launch-style, 2048x1024, t in [1,11], n=201, ~30M events
    "C3": dict(sensor_w=240, sensor_h=180, fx=200.0, pano_w=2048, pano_h=1024, C_th=0.2, t_beg=1.0, t_end=11.0,
               dt_knots=0.05, texture_std=1.9, periodic=True),
    # C4: as C3 with ~100M events
    "C4": dict(sensor_w=240, sensor_h=180, fx=200.0, pano_w=2048, pano_h=1024, C_th=0.2, t_beg=1.0, t_end=11.0,
               dt_knots=0.05, texture_std=6.3, periodic=True),
}


def make_config(name, **over) -> Scene:
    kw = dict(CONFIGS[name])
    kw.update(over)
    return make_scene(**kw)
