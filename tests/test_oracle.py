"""The CPU oracle against its committed golden vectors and against itself (float64 vs float32 vs the
torch-CPU port), plus the sensitivity guards of SURVEY.md section 8c: wrong structural constants must
move the angles, so a silent change of eps / padding cannot pass."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import whenet_oracle as wo
from conftest import SNAP


def test_golden_angles_fp64(oracle64, sample_crops, golden):
    yaw, pitch, roll = oracle64.get_angle(sample_crops)
    assert yaw.dtype == np.float32
    for i, s in enumerate(golden["samples"]):
        assert abs(yaw[i] - s["yaw"]) < 1e-4 and abs(pitch[i] - s["pitch"]) < 1e-4 and abs(roll[i] - s["roll"]) < 1e-4
    # the values SURVEY.md 8c recorded independently at survey time
    assert np.allclose(yaw, [15.6566, -136.6856], atol=1e-3)
    assert np.allclose(pitch, [3.8123, -25.9486], atol=1e-3)
    assert np.allclose(roll, [10.8108, 5.4601], atol=1e-3)


def test_golden_logits(oracle64, sample_crops):
    import os
    from conftest import GOLD
    ref = np.load(os.path.join(GOLD, "sample_logits_f64.npy"))
    _a, logits = oracle64.get_angle(sample_crops, return_logits=True)
    assert np.abs(np.concatenate(logits, axis=1) - ref).max() < 1e-9
    assert [int(np.argmax(l[0])) for l in logits] == [69, 27, 44]
    assert [int(np.argmax(l[1])) for l in logits] == [23, 18, 37]


def test_fp32_vs_fp64(oracle32, oracle64, sample_crops):
    a, b = oracle32.get_angle(sample_crops), oracle64.get_angle(sample_crops)
    assert max(np.abs(x - y).max() for x, y in zip(a, b)) < 1e-3


def test_block_boundary_shapes(oracle32, sample_crops):
    taps = {}
    oracle32.get_angle(sample_crops[:1], taps)
    assert taps["stem"].shape == (1, 112, 112, 32)
    want = [16, 24, 24, 40, 40, 80, 80, 80, 112, 112, 112, 192, 192, 192, 192, 320]
    hw = [112, 56, 56, 28, 28, 14, 14, 14, 14, 14, 14, 7, 7, 7, 7, 7]
    for i in range(16):
        assert taps["block%d" % (i + 1)].shape == (1, hw[i], hw[i], want[i])
    assert taps["head"].shape == (1, 7, 7, 1280) and taps["pooled"].shape == (1, 1280)


def test_wrong_bn_eps_moves_angles(sample_crops, oracle64):
    bad = wo.load_oracle(SNAP, np.float64, bn_eps=1e-5)
    d = max(np.abs(x - y).max() for x, y in zip(bad.get_angle(sample_crops), oracle64.get_angle(sample_crops)))
    assert d > 0.3


def test_wrong_padding_moves_angles(sample_crops, oracle64):
    """PyTorch-style symmetric padding instead of TF 'SAME' (SURVEY.md 8c measured 16.6 deg): must be far outside any tolerance."""
    bad = wo.load_oracle(SNAP, np.float64, symmetric_pad=True)
    d = max(np.abs(x - y).max() for x, y in zip(bad.get_angle(sample_crops), oracle64.get_angle(sample_crops)))
    assert d > 1.0


def test_batch_composition_irrelevant(oracle64, sample_crops):
    both = oracle64.get_angle(sample_crops)
    one = oracle64.get_angle(sample_crops[1:2])
    assert max(abs(float(b[1]) - float(o[0])) for b, o in zip(both, one)) < 1e-4


def test_torch_port_matches_numpy_oracle(oracle32, jitter_crops):
    from whenet_b200 import weights
    names, w = weights.load_snapshot(SNAP)
    port = wo.TorchCpuPort(names, w)
    a = port.get_angle(jitter_crops)
    b = oracle32.get_angle(jitter_crops)
    assert max(np.abs(x - y).max() for x, y in zip(a, b)) < 5e-3


def test_input_shape_checked(oracle64):
    with pytest.raises(ValueError):
        oracle64.forward_normalised(np.zeros((1, 200, 224, 3)))


def test_preprocess_is_reference_formula():
    img = np.arange(2 * 224 * 224 * 3, dtype=np.uint8).reshape(2, 224, 224, 3)
    x = wo.preprocess(img)
    assert x.dtype == np.float64
    assert x[0, 0, 0, 1] == (1 / 255 - 0.456) / 0.224


@settings(max_examples=50, deadline=None)
@given(st.lists(st.floats(-80, 80, allow_nan=False, width=32), min_size=66, max_size=66), st.floats(-1e4, 1e4, width=32))
def test_softmax_decode_properties(logit_row, shift):
    x = np.array([logit_row], dtype=np.float32)
    p = wo.softmax(x.copy())
    assert abs(p.sum() - 1) < 1e-5 and (p >= 0).all()
    # shift invariance (the reference subtracts the row max, utils.py:8)
    p2 = wo.softmax(x + np.float32(shift))
    assert np.abs(p - p2).max() < 1e-3
    yaw_logits = np.zeros((1, 120), np.float32)
    y, pi, r = wo.decode(yaw_logits, x, x)
    assert -99 - 1e-3 <= pi[0] <= 96 + 1e-3 and pi[0] == r[0]
    assert abs(y[0] - (59.5 * 3 - 180)) < 1e-3        # uniform yaw distribution -> mean bin index 59.5


def test_decode_onehot_bins():
    for b in (0, 17, 65):
        x = np.full((1, 66), -1e4, np.float32)
        x[0, b] = 0
        yaw = np.full((1, 120), -1e4, np.float32)
        yaw[0, b] = 0
        y, p, r = wo.decode(yaw, x, x)
        assert abs(y[0] - (b * 3 - 180)) < 1e-4 and abs(p[0] - (b * 3 - 99)) < 1e-4 and abs(r[0] - (b * 3 - 99)) < 1e-4


def test_softmax_does_not_mutate():
    x = np.array([[1.0, 2.0, 3.0]], np.float32)
    y = x.copy()
    wo.softmax(x)
    assert np.array_equal(x, y)
