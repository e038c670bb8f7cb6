"""WHENet network description shared by the weight packer, the bench and the tests.

The reference assembles the graph as ``efn.EfficientNetB0(include_top=False,
input_shape=(224,224,3))`` -> ``GlobalAveragePooling2D`` -> three ``Dense``
heads (reference ``whenet.py:8-14``).  The EfficientNet-B0 builder itself lives
in the un-vendored pip package ``efficientnet==0.0.4`` (reference
``requirements.txt:5``); its block table below is the published B0 table and
is cross-checked against every tensor shape in ``WHENet.h5`` by
``assign_weights`` (65 conv2d / 16 depthwise / 49 BN consumed in the order of
the file's ``layer_names`` attribute).

Nothing in this file computes: it is the single place that says which tensor
feeds which layer, so the oracle and the CUDA packer cannot disagree silently.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

IMG = 224
BN_EPS = 1e-3          # efficientnet==0.0.4 BatchNormalization epsilon (not in the .h5)
MEAN = (0.485, 0.456, 0.406)   # reference whenet.py:23
STD = (0.229, 0.224, 0.225)    # reference whenet.py:24
N_YAW, N_PITCH, N_ROLL = 120, 66, 66   # reference whenet.py:11-13
N_LOGITS = N_YAW + N_PITCH + N_ROLL    # 252
STEM_C = 32
HEAD_C = 1280

# (kernel, stride, expand_ratio, Cin, Cout, repeats) - EfficientNet-B0
_STAGES = [
    (3, 1, 1, 32, 16, 1),
    (3, 2, 6, 16, 24, 2),
    (5, 2, 6, 24, 40, 2),
    (3, 2, 6, 40, 80, 3),
    (5, 1, 6, 80, 112, 3),
    (5, 2, 6, 112, 192, 4),
    (3, 1, 6, 192, 320, 1),
]


@dataclass
class Block:
    idx: int          # 1..16
    hin: int
    hout: int
    cin: int
    cexp: int
    cout: int
    k: int
    s: int
    cse: int
    skip: bool
    has_expand: bool
    # filled by assign_weights: names inside the weights dict
    w: Dict[str, str] = field(default_factory=dict)

    @property
    def pad_before(self) -> int:
        """TF 'SAME' padding placed before (top/left); the rest goes after."""
        total = max((self.hout - 1) * self.s + self.k - self.hin, 0)
        return total // 2


def same_pad(n_in: int, k: int, s: int):
    """TensorFlow 'SAME': out=ceil(in/s); pad_total=max((out-1)s+k-in,0); before=floor(total/2)."""
    n_out = -(-n_in // s)
    total = max((n_out - 1) * s + k - n_in, 0)
    return n_out, total // 2, total - total // 2


def blocks() -> List[Block]:
    out: List[Block] = []
    h = IMG // 2  # after the stride-2 stem
    idx = 0
    for (k, s, e, cin, cout, rep) in _STAGES:
        for r in range(rep):
            idx += 1
            b_cin = cin if r == 0 else cout
            b_s = s if r == 0 else 1
            hout = -(-h // b_s)
            out.append(Block(idx=idx, hin=h, hout=hout, cin=b_cin, cexp=b_cin * e, cout=cout,
                             k=k, s=b_s, cse=max(1, int(b_cin * 0.25)),
                             skip=(b_s == 1 and b_cin == cout), has_expand=(e != 1)))
            h = hout
    return out


def assign_weights(weights: Dict[str, np.ndarray]) -> List[Block]:
    """Bind tensors to layers in Keras topological order and shape-check all of them.

    Mirrors what ``Model.load_weights`` does (reference ``whenet.py:16``): the
    i-th conv2d / depthwise_conv2d / batch_normalization layer in graph order
    receives the i-th such group of the file.
    """
    conv = dw = bn = 0

    def take_conv(shape, bias=False):
        nonlocal conv
        conv += 1
        name = "conv2d_%d" % conv
        got = weights[name + "/kernel:0"].shape
        if tuple(got) != tuple(shape):
            raise ValueError("%s: kernel shape %s, expected %s" % (name, got, shape))
        if bias != ((name + "/bias:0") in weights):
            raise ValueError("%s: unexpected bias presence" % name)
        return name

    def take_dw(k, c):
        nonlocal dw
        dw += 1
        name = "depthwise_conv2d_%d" % dw
        got = weights[name + "/depthwise_kernel:0"].shape
        if tuple(got) != (k, k, c, 1):
            raise ValueError("%s: shape %s, expected %s" % (name, got, (k, k, c, 1)))
        return name

    def take_bn(c):
        nonlocal bn
        bn += 1
        name = "batch_normalization_%d" % bn
        for part in ("gamma", "beta", "moving_mean", "moving_variance"):
            got = weights["%s/%s:0" % (name, part)].shape
            if tuple(got) != (c,):
                raise ValueError("%s/%s: shape %s, expected (%d,)" % (name, part, got, c))
        return name

    stem = {"conv": take_conv((3, 3, 3, STEM_C)), "bn": take_bn(STEM_C)}
    blks = blocks()
    for b in blks:
        if b.has_expand:
            b.w["expand"] = take_conv((1, 1, b.cin, b.cexp))
            b.w["expand_bn"] = take_bn(b.cexp)
        b.w["dw"] = take_dw(b.k, b.cexp)
        b.w["dw_bn"] = take_bn(b.cexp)
        b.w["se_reduce"] = take_conv((1, 1, b.cexp, b.cse), bias=True)
        b.w["se_expand"] = take_conv((1, 1, b.cse, b.cexp), bias=True)
        b.w["project"] = take_conv((1, 1, b.cexp, b.cout))
        b.w["project_bn"] = take_bn(b.cout)
    head = {"conv": take_conv((1, 1, blks[-1].cout, HEAD_C)), "bn": take_bn(HEAD_C)}
    for nm, units in (("yaw_new", N_YAW), ("pitch_new", N_PITCH), ("roll_new", N_ROLL)):
        if tuple(weights[nm + "/kernel:0"].shape) != (HEAD_C, units):
            raise ValueError("%s kernel shape %s" % (nm, weights[nm + "/kernel:0"].shape))
        if tuple(weights[nm + "/bias:0"].shape) != (units,):
            raise ValueError("%s bias shape" % nm)
    if (conv, dw, bn) != (65, 16, 49):
        raise ValueError("consumed %d/%d/%d conv/dw/bn, expected 65/16/49" % (conv, dw, bn))
    return stem, blks, head


def random_weights(seed: int = 0) -> Dict[str, np.ndarray]:
    """Randomly-initialised tensors with the WHENet.h5 names/shapes.

    ``WHENet(snapshot=None)`` in the reference leaves Keras' default
    initialisers in place (reference ``whenet.py:15``); the exact values are
    unspecified there, so this only has to be *a* valid initialisation.  Used
    for synthetic-weight benches and tests that must not depend on the
    reference artefact.
    """
    rng = np.random.default_rng(seed)
    w: Dict[str, np.ndarray] = {}
    conv = dw = bn = 0

    def add_conv(shape, bias=False):
        nonlocal conv
        conv += 1
        fan_in = shape[0] * shape[1] * shape[2]
        w["conv2d_%d/kernel:0" % conv] = (rng.standard_normal(shape) * np.sqrt(1.0 / fan_in)).astype(np.float32)
        if bias:
            w["conv2d_%d/bias:0" % conv] = (rng.standard_normal(shape[3]) * 0.1).astype(np.float32)

    def add_dw(k, c):
        nonlocal dw
        dw += 1
        w["depthwise_conv2d_%d/depthwise_kernel:0" % dw] = (
            rng.standard_normal((k, k, c, 1)) * np.sqrt(1.0 / (k * k))).astype(np.float32)

    def add_bn(c):
        nonlocal bn
        bn += 1
        p = "batch_normalization_%d/" % bn
        w[p + "gamma:0"] = rng.uniform(0.8, 1.6, c).astype(np.float32)
        w[p + "beta:0"] = (rng.standard_normal(c) * 0.2).astype(np.float32)
        w[p + "moving_mean:0"] = (rng.standard_normal(c) * 0.2).astype(np.float32)
        w[p + "moving_variance:0"] = rng.uniform(0.5, 1.5, c).astype(np.float32)

    add_conv((3, 3, 3, STEM_C)); add_bn(STEM_C)
    for b in blocks():
        if b.has_expand:
            add_conv((1, 1, b.cin, b.cexp)); add_bn(b.cexp)
        add_dw(b.k, b.cexp); add_bn(b.cexp)
        add_conv((1, 1, b.cexp, b.cse), bias=True)
        add_conv((1, 1, b.cse, b.cexp), bias=True)
        add_conv((1, 1, b.cexp, b.cout)); add_bn(b.cout)
    add_conv((1, 1, 320, HEAD_C)); add_bn(HEAD_C)
    for nm, units in (("yaw_new", N_YAW), ("pitch_new", N_PITCH), ("roll_new", N_ROLL)):
        w[nm + "/kernel:0"] = (rng.standard_normal((HEAD_C, units)) * np.sqrt(1.0 / HEAD_C)).astype(np.float32)
        w[nm + "/bias:0"] = (rng.standard_normal(units) * 0.05).astype(np.float32)
    return w


def macs_per_crop() -> Dict[str, int]:
    """Multiply-accumulates per 224x224 crop, by op class."""
    m = {"stem": 112 * 112 * 27 * 32, "expand": 0, "dw": 0, "se": 0, "project": 0}
    for b in blocks():
        if b.has_expand:
            m["expand"] += b.hin * b.hin * b.cin * b.cexp
        m["dw"] += b.hout * b.hout * b.k * b.k * b.cexp
        # squeeze (mean over H,W) + two FCs + the gate multiply over H,W
        m["se"] += 2 * b.cexp * b.cse + 2 * b.hout * b.hout * b.cexp
        m["project"] += b.hout * b.hout * b.cexp * b.cout
    m["head"] = 49 * 320 * HEAD_C + 49 * HEAD_C     # 1x1 conv + global average pool
    m["fc"] = HEAD_C * N_LOGITS
    m["total"] = sum(m.values())
    return m
