"""Flat binary event-window format (the on-disk / wire format next to the hot path; ROS
bags stay on the host side of the reference).

Layout (little endian): 32-byte header  b"EMBAEV01", int64 N, int32 sensor_w, int32 sensor_h, int64 reserved
followed by the SoA arrays in this order: x uint16[N], y uint16[N], polarity uint8[N], padding to 8 bytes,
t_ns int64[N]. Events are time-sorted (the reference sorts them after reading the bag,
src/utils/rosbag_loading.cpp:61-65); the loader checks it. The arrays are exactly what `emba_set_events` takes.
"""
from __future__ import annotations

import numpy as np

MAGIC = b"EMBAEV01"


def save_events(path, x, y, t_ns, pol, sensor_w, sensor_h):
    x = np.ascontiguousarray(x, dtype="<u2")
    y = np.ascontiguousarray(y, dtype="<u2")
    p = np.ascontiguousarray(pol, dtype=np.uint8)
    t = np.ascontiguousarray(t_ns, dtype="<i8")
    n = x.size
    if not (y.size == n and p.size == n and t.size == n):
        raise ValueError("event arrays differ in length")
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(np.array([n], dtype="<i8").tobytes())
        f.write(np.array([sensor_w, sensor_h], dtype="<i4").tobytes())
        f.write(np.array([0], dtype="<i8").tobytes())
        f.write(x.tobytes())
        f.write(y.tobytes())
        f.write(p.tobytes())
        f.write(b"\0" * ((-(5 * n)) % 8))
        f.write(t.tobytes())


def load_events(path):
    """Returns x, y, t_ns, pol, (sensor_w, sensor_h); memory-maps the file (the arrays can be handed to the C ABI
    without a copy)."""
    raw = np.memmap(path, dtype=np.uint8, mode="r")
    if raw.size < 32 or bytes(raw[:8]) != MAGIC:
        raise ValueError("not an EMBAEV01 file")
    n = int(raw[8:16].view("<i8")[0])
    w, h = (int(v) for v in raw[16:24].view("<i4"))
    off = 32
    need = off + 5 * n + ((-(5 * n)) % 8) + 8 * n
    if n < 0 or raw.size != need:
        raise ValueError("truncated or oversized EMBAEV01 file")
    x = raw[off:off + 2 * n].view("<u2")
    y = raw[off + 2 * n:off + 4 * n].view("<u2")
    p = raw[off + 4 * n:off + 5 * n]
    toff = off + 5 * n + ((-(5 * n)) % 8)
    t = raw[toff:toff + 8 * n].view("<i8")
    if n > 1 and np.any(np.diff(t) < 0):
        raise ValueError("events are not time-sorted")
    if n and (x.max() >= w or y.max() >= h):
        raise ValueError("event coordinates outside the sensor")
    return x, y, t, p, (w, h)
