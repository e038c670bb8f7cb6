"""Second, independent anchor for the oracle's STRUCTURE (last row): torchvision's ``efficientnet_b0``.

The reference's Keras stack cannot run here, so the oracle's golden angles are its own ("parity unpinned").  What CAN be
checked against code the builder did not write is the graph: torchvision ships an EfficientNet-B0 whose block table, strides,
squeeze widths, skip rule, SE ordering and SiLU were written by the PyTorch maintainers.  The 315 tensors of ``WHENet.h5`` are
poured into that module tree in traversal order (65 convs, 16 depthwise convs, 49 BatchNorms - counts and every shape must
match, or the load raises), BatchNorm eps is set to the efficientnet==0.0.4 value (1e-3) and the stride-2 convolutions get
TensorFlow's asymmetric 'SAME' padding.  Then:

  * torchvision-graph + WHENet weights + TF padding   == oracle (float32)       to 2e-3 deg
  * torchvision's NATIVE symmetric padding             moves the angles > 1 deg  (the wrong-padding guard, from both sides:
    torchvision's own graph and the oracle's ``symmetric_pad`` knob agree with each other on HOW wrong it is)
  * eps = 1e-5 (torchvision's default)                 moves the angles > 0.3 deg
"""
import numpy as np
import pytest

import whenet_oracle as wo
from conftest import SNAP

torch = pytest.importorskip("torch")
tv = pytest.importorskip("torchvision")


def _tv_whenet(weights, tf_same_pad=True, eps=1e-3):
    nn = torch.nn
    F = torch.nn.functional
    m = tv.models.efficientnet_b0(weights=None).eval()
    convs = [c for c in m.features.modules() if isinstance(c, nn.Conv2d)]
    bns = [b for b in m.features.modules() if isinstance(b, nn.BatchNorm2d)]
    assert len(convs) == 81 and len(bns) == 49
    ci = di = 0
    with torch.no_grad():
        for c in convs:
            if c.groups > 1:
                di += 1
                w = weights["depthwise_conv2d_%d/depthwise_kernel:0" % di]            # [kh,kw,C,1]
                t = torch.from_numpy(np.ascontiguousarray(w.transpose(2, 3, 0, 1)))   # [C,1,kh,kw]
            else:
                ci += 1
                w = weights["conv2d_%d/kernel:0" % ci]                                # HWIO
                t = torch.from_numpy(np.ascontiguousarray(w.transpose(3, 2, 0, 1)))   # OIHW
                if c.bias is not None:
                    b = torch.from_numpy(weights["conv2d_%d/bias:0" % ci])
                    assert c.bias.shape == b.shape
                    c.bias.copy_(b)
                else:
                    assert ("conv2d_%d/bias:0" % ci) not in weights
            assert tuple(c.weight.shape) == tuple(t.shape), (ci, di, c.weight.shape, t.shape)
            c.weight.copy_(t)
        assert ci == 65 and di == 16
        for i, b in enumerate(bns, 1):
            p = "batch_normalization_%d/" % i
            for dst, src in ((b.weight, "gamma:0"), (b.bias, "beta:0"), (b.running_mean, "moving_mean:0"), (b.running_var, "moving_variance:0")):
                t = torch.from_numpy(weights[p + src])
                assert dst.shape == t.shape, (i, src)
                dst.copy_(t)
            b.eps = eps
    if tf_same_pad:
        for c in convs:
            if c.stride[0] == 2:
                k = c.kernel_size[0]
                total = k - 2                       # even inputs: (out-1)*2 + k - in = k - 2
                before, after = total // 2, total - total // 2
                c.padding = (0, 0)
                c.register_forward_pre_hook(lambda mod, args, b=before, a=after: (F.pad(args[0], (b, a, b, a)),))
    heads = [(torch.from_numpy(weights[n + "/kernel:0"]), torch.from_numpy(weights[n + "/bias:0"])) for n in ("yaw_new", "pitch_new", "roll_new")]

    def get_angle(img):
        x = torch.from_numpy(wo.preprocess(img).astype(np.float32)).permute(0, 3, 1, 2).contiguous()
        with torch.no_grad():
            f = m.avgpool(m.features(x)).flatten(1)
            logits = [(f @ k + b).numpy() for k, b in heads]
        return np.stack(wo.decode(*logits), axis=1)
    return get_angle


@pytest.fixture(scope="module")
def raw_weights():
    from whenet_b200 import weights
    return weights.load_snapshot(SNAP)[1]


def test_torchvision_graph_reproduces_oracle(raw_weights, oracle32, sample_crops, jitter_crops):
    crops = np.concatenate([sample_crops, jitter_crops[:2]])
    got = _tv_whenet(raw_weights)(crops)
    ref = np.stack(oracle32.get_angle(crops), axis=1)
    assert np.abs(got - ref).max() < 2e-3, np.abs(got - ref).max()


def test_wrong_padding_moves_angles_both_implementations(raw_weights, oracle64, sample_crops):
    ref = np.stack(oracle64.get_angle(sample_crops), axis=1)
    tv_sym = _tv_whenet(raw_weights, tf_same_pad=False)(sample_crops)                       # torchvision's native padding
    or_sym = np.stack(wo.load_oracle(SNAP, np.float32, symmetric_pad=True).get_angle(sample_crops), axis=1)
    assert np.abs(tv_sym - ref).max() > 1.0 and np.abs(or_sym - ref).max() > 1.0            # SURVEY.md 8c measured 16.6 deg
    assert np.abs(tv_sym - or_sym).max() < 5e-3                                              # ... and they are wrong the same way


def test_wrong_eps_moves_angles_torchvision(raw_weights, oracle64, sample_crops):
    ref = np.stack(oracle64.get_angle(sample_crops), axis=1)
    got = _tv_whenet(raw_weights, eps=1e-5)(sample_crops)
    assert np.abs(got - ref).max() > 0.3
