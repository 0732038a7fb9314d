"""Flat binary event-window format (emba_b200/eventio.py): round trip, ragged sizes, corruption checks."""
import numpy as np
import pytest

from emba_b200 import eventio


@pytest.mark.parametrize("n", [0, 1, 7, 100, 12345])
def test_roundtrip(tmp_path, n):
    rng = np.random.default_rng(n)
    x = rng.integers(0, 240, n).astype(np.uint16)
    y = rng.integers(0, 180, n).astype(np.uint16)
    t = np.sort(rng.integers(0, 10**10, n)).astype(np.int64)
    p = rng.integers(0, 2, n).astype(np.uint8)
    f = tmp_path / "ev.bin"
    eventio.save_events(f, x, y, t, p, 240, 180)
    x2, y2, t2, p2, (w, h) = eventio.load_events(f)
    assert (w, h) == (240, 180)
    assert np.array_equal(x, x2) and np.array_equal(y, y2) and np.array_equal(t, t2) and np.array_equal(p, p2)


def test_golden_scene_roundtrip(tmp_path, tiny):
    f = tmp_path / "tiny.bin"
    eventio.save_events(f, tiny.x, tiny.y, tiny.t_ns, tiny.pol, tiny.sensor_w, tiny.sensor_h)
    x, y, t, p, _ = eventio.load_events(f)
    assert np.array_equal(t, tiny.t_ns) and np.array_equal(x, tiny.x) and np.array_equal(p, tiny.pol)


def test_rejects_bad_files(tmp_path):
    f = tmp_path / "bad.bin"
    f.write_bytes(b"nonsense")
    with pytest.raises(ValueError):
        eventio.load_events(f)
    x = np.array([1, 2], dtype=np.uint16)
    t = np.array([5, 3], dtype=np.int64)  # not sorted
    eventio.save_events(f, x, x, t, np.zeros(2, np.uint8), 8, 8)
    with pytest.raises(ValueError):
        eventio.load_events(f)
    eventio.save_events(f, x, x, np.sort(t), np.zeros(2, np.uint8), 2, 8)  # x == 2 outside a 2-wide sensor
    with pytest.raises(ValueError):
        eventio.load_events(f)
    with pytest.raises(ValueError):
        eventio.save_events(f, x, x[:1], t, np.zeros(2, np.uint8), 8, 8)
