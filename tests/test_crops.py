"""Crop front-end: the oracle against OpenCV itself (pinned: cv2 runs here), the host
margin arithmetic against the oracle's restatement of demo_video.py:13-21, and - on the GPU - the CUDA
kernel bit-for-bit against cv2.resize."""
import numpy as np
import pytest

import crop_oracle as co
from conftest import SNAP

cv2 = pytest.importorskip("cv2")


def _sizes():
    rng = np.random.default_rng(0)
    fixed = [(83, 64), (187, 164), (224, 224), (448, 448), (50, 400), (400, 48), (1, 1), (2, 3), (225, 223), (449, 447), (17, 1080)]
    return fixed + [(int(rng.integers(3, 600)), int(rng.integers(3, 600))) for _ in range(20)]


def test_resize_oracle_equals_cv2_bit_for_bit():
    rng = np.random.default_rng(1)
    for h, w in _sizes():
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(co.resize_linear_u8(img), cv2.resize(img, (224, 224))), (h, w)


def test_demo_crop_recipe_matches_committed_fixture(sample_crops):
    """The committed Sample crops were produced by cv2 through demo.py:7-12; a re-crop of a synthetic frame with the
    oracle must agree with cv2 as well (guards the BGR->RGB order and the slice convention)."""
    rng = np.random.default_rng(2)
    frame = rng.integers(0, 256, (360, 640, 3), dtype=np.uint8)
    boxes = np.array([[10.2, 20.7, 200.9, 180.1], [0, 0, 360, 640], [300.5, 600.5, 359.9, 639.9]], np.float32)
    got = co.crop_batch(frame, boxes, margin=True)
    for i, b in enumerate(boxes):
        y0, y1, x0, x1 = co.enlarge_box(b, 360, 640)
        ref = cv2.resize(cv2.cvtColor(frame[y0:y1, x0:x1], cv2.COLOR_BGR2RGB), (224, 224))
        assert np.array_equal(got[i], ref)
    assert sample_crops.shape == (2, 224, 224, 3)


def test_margin_arithmetic_host_equals_oracle():
    from whenet_b200 import crops
    rng = np.random.default_rng(3)
    for _ in range(500):
        h, w = int(rng.integers(100, 1200)), int(rng.integers(100, 2000))
        y0, x0 = rng.uniform(0, h - 10), rng.uniform(0, w - 10)
        box = np.array([y0, x0, y0 + rng.uniform(5, h), x0 + rng.uniform(5, w)], np.float32)
        assert crops.enlarge_box(box, h, w) == co.enlarge_box(box, h, w)
    # the reference's quirk: the far side grows by a fraction of the ALREADY enlarged extent (demo_video.py:15-16)
    assert crops.enlarge_box((100, 100, 200, 200), 1000, 1000) == (90, 211, 80, 224)


@pytest.mark.gpu
def test_gpu_crop_resize_equals_cv2():
    import torch
    import whenet_b200
    from whenet_b200._lib import check
    from whenet_b200.whenet import _ptr
    m = whenet_b200.WHENet(None, device=0, precision="bf16", max_batch=64)
    rng = np.random.default_rng(4)
    frame = rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
    rects = [(0, 1080, 0, 1920), (0, 448, 0, 448), (100, 101, 200, 201), (5, 229, 7, 231), (1000, 1080, 1800, 1920), (0, 2, 0, 3)]
    for _ in range(40):
        y0, x0 = int(rng.integers(0, 1070)), int(rng.integers(0, 1910))
        rects.append((y0, int(rng.integers(y0 + 1, 1081)), x0, int(rng.integers(x0 + 1, 1921))))
    r = np.array(rects, dtype=np.int32)
    out = torch.empty((len(rects), 224, 224, 3), dtype=torch.uint8, device="cuda")
    check(m._L.whenet_crop_resize_u8(m._h, _ptr(frame), 1080, 1920, 0, _ptr(r), len(rects), 1, _ptr(out)))
    m.synchronize()
    got = out.cpu().numpy()
    for i, (y0, y1, x0, x1) in enumerate(rects):
        ref = cv2.resize(cv2.cvtColor(frame[y0:y1, x0:x1], cv2.COLOR_BGR2RGB), (224, 224))
        assert np.array_equal(got[i], ref), (i, rects[i])
    # invalid slices are refused like cv2.resize refuses an empty image
    bad = np.array([[10, 10, 0, 5]], dtype=np.int32)
    assert m._L.whenet_crop_resize_u8(m._h, _ptr(frame), 1080, 1920, 0, _ptr(bad), 1, 1, _ptr(out)) == -1
    m.close()


@pytest.mark.gpu
def test_stream_path_equals_per_head_reference_path():
    """get_angle_from_frame (all heads batched, crops made on the GPU) == get_angle on crops made with cv2 one by
    one as demo_video.py:21-27 does - bitwise, because the crops are bit-identical and the kernels batch invariant."""
    import whenet_b200
    m = whenet_b200.WHENet(SNAP, device=0, precision="fp32", max_batch=16)
    rng = np.random.default_rng(5)
    frame = rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8)
    boxes = np.array([[50.3, 60.1, 250.2, 300.9], [400, 900, 719, 1279], [0, 0, 100, 80], [300.7, 500.2, 460.1, 640.6]], np.float32)
    y, p, r, crops = m.get_angle_from_frame(frame, boxes, return_crops=True)
    ref_crops = co.crop_batch(frame, boxes, margin=True)
    assert np.array_equal(crops, ref_crops)
    y2, p2, r2 = m.get_angle(ref_crops)
    assert np.array_equal(y, y2) and np.array_equal(p, p2) and np.array_equal(r, r2)
    e = m.get_angle_from_frame(frame, np.zeros((0, 4), np.float32))
    assert e[0].shape == (0,)
    m.close()
