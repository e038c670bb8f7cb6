"""Host side of the stream-path crop front-end.

``enlarge_box`` mirrors the margin / clamp / truncate arithmetic of reference
``demo_video.py:13-21`` in the same float32 scalar arithmetic numpy performs there; the pixel work
(slice, BGR->RGB, cv2-exact bilinear resize to 224x224 for every head of the frame at once) runs in
``crop_resize_kernel`` through ``whenet_crop_resize_u8``.
"""
from __future__ import annotations

import numpy as np


def enlarge_bounds(bbox, height: int, width: int):
    """(y_min, x_min, y_max, x_max) from the detector -> the margin-enlarged FLOAT bounds (y_min, y_max, x_min, x_max).

    The reference updates ``y_min`` / ``x_min`` first and then uses the UPDATED value for the far side
    (demo_video.py:15-18); reproduced as is.  The overlay (demo_video.py:25-29) uses these before truncation."""
    y_min, x_min, y_max, x_max = [np.float32(v) for v in bbox]
    y_min = max(0, y_min - abs(y_min - y_max) / 10)
    y_max = min(height, y_max + abs(y_min - y_max) / 10)
    x_min = max(0, x_min - abs(x_min - x_max) / 5)
    x_max = min(width, x_max + abs(x_min - x_max) / 5)
    x_max = min(x_max, width)
    return y_min, y_max, x_min, x_max


def enlarge_box(bbox, height: int, width: int):
    """... truncated to the integer slice bounds (y0, y1, x0, x1) of demo_video.py:21."""
    y_min, y_max, x_min, x_max = enlarge_bounds(bbox, height, width)
    return int(y_min), int(y_max), int(x_min), int(x_max)


def rects_from_boxes(boxes, height: int, width: int, margin: bool = True) -> np.ndarray:
    """(M,4) detector boxes -> (M,4) int32 slice bounds (y0, y1, x0, x1)."""
    out = np.empty((len(boxes), 4), dtype=np.int32)
    for i, b in enumerate(boxes):
        if margin:
            out[i] = enlarge_box(b, height, width)
        else:
            y0, x0, y1, x1 = [int(v) for v in b]
            out[i] = (y0, y1, x0, x1)
    return out
