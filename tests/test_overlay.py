"""draw_axis / frame annotation against the line-by-line restatement of reference utils.py:13-43 and
demo_video.py:11-35, including the reference's draw-before-next-crop order."""
import numpy as np
import pytest

import overlay_oracle as oo
from whenet_b200 import overlay


class _FakeModel:
    """Angles that depend on the crop's content (so a crop cut from an annotated frame gives different angles)."""
    def get_angle(self, img):
        img = np.asarray(img, np.float64)
        m = img.reshape(img.shape[0], -1).mean(axis=1)
        return ((m * 1.7) % 360 - 180).astype(np.float32), ((m * 0.9) % 180 - 90).astype(np.float32), ((m * 0.5) % 180 - 90).astype(np.float32)


def test_axis_endpoints_known_answers():
    (x1, y1), (x2, y2), (x3, y3) = overlay.axis_endpoints(0.0, 0.0, 0.0, 50.0, 60.0, 10)
    assert (x1, y1) == (60.0, 60.0) and (x2, y2) == (50.0, 70.0) and (x3, y3) == (50.0, 60.0)
    (_a, _b, (x3, y3)) = overlay.axis_endpoints(90.0, 0.0, 0.0, 0.0, 0.0, 10)     # yaw +90 deg: the Z axis points to -x (yaw is negated)
    assert abs(x3 + 10) < 1e-9 and abs(y3) < 1e-9


@pytest.mark.parametrize("seed", range(5))
def test_draw_axis_pixel_identical(seed):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (240, 320, 3), dtype=np.uint8)
    yaw, pitch, roll = rng.uniform(-180, 180), rng.uniform(-99, 96), rng.uniform(-99, 96)
    for kw in ({}, {"tdx": 100.5, "tdy": 77.25, "size": 43.0}):
        a = overlay.draw_axis(img.copy(), yaw, pitch, roll, **kw)
        b = oo.draw_axis_ref(img.copy(), yaw, pitch, roll, **kw)
        assert np.array_equal(a, b)


@pytest.mark.parametrize("display", ["simple", "full"])
def test_process_frame_reference_order(display):
    rng = np.random.default_rng(7)
    frame = rng.integers(0, 256, (360, 480, 3), dtype=np.uint8)
    boxes = [(40.0, 60.0, 160.0, 170.0), (60.0, 120.0, 200.0, 260.0), (200.0, 300.0, 340.0, 470.0)]     # overlapping heads
    model = _FakeModel()
    ref = frame.copy()
    want = []
    for b in boxes:
        ref, ang = oo.process_detection_ref(model, ref, b, display)
        want.append(ang)
    got, yaw, pitch, roll = overlay.process_frame(model, frame.copy(), boxes, display=display, reference_order=True)
    assert np.array_equal(got, ref)
    assert np.allclose(np.stack([yaw, pitch, roll], axis=1), np.array(want, np.float32))
