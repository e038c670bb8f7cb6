"""CPU oracle for the crop front-end of the stream path.  TEST INFRASTRUCTURE ONLY.

Restates, for uint8 3-channel images:
  * the margin / clamp / truncate arithmetic of reference demo_video.py:13-21
  * cv2.resize(img, (224, 224)) with the default INTER_LINEAR as OpenCV's own 8-bit
    fixed-point kernel computes it (reference demo_video.py:23, demo.py:11): half-pixel centres,
    11-bit coefficients rounded to nearest-even, horizontal pass in int32, vertical pass
    ((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2 >> 2; exact 2x down-scaling takes the 2x2 box path.
Unlike the network, this one IS pinned: cv2 runs in this image, tests compare bit-for-bit against it.
"""
import numpy as np


def enlarge_box(bbox, height, width):
    """demo_video.py:13-21.  bbox = (y_min, x_min, y_max, x_max) as the detector returns it (float32).
    Returns the integer slice bounds (y0, y1, x0, x1).  Note the reference updates y_min / x_min first
    and then uses the UPDATED value for the max side (demo_video.py:15-18)."""
    y_min, x_min, y_max, x_max = [np.float32(v) for v in bbox]
    y_min = max(0, y_min - abs(y_min - y_max) / 10)
    y_max = min(height, y_max + abs(y_min - y_max) / 10)
    x_min = max(0, x_min - abs(x_min - x_max) / 5)
    x_max = min(width, x_max + abs(x_min - x_max) / 5)
    x_max = min(x_max, width)
    return int(y_min), int(y_max), int(x_min), int(x_max)


def _axis_tables(src, dst):
    """OpenCV resize(): per destination index the left source index and the two 11-bit weights."""
    scale = 1.0 / (float(dst) / float(src))                      # double, as inv_scale -> scale
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    return s, f


def resize_linear_u8(img, dst_h=224, dst_w=224):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    if h == 2 * dst_h and w == 2 * dst_w:                         # resize(): INTER_LINEAR -> INTER_AREA fast path
        x = img.astype(np.int32)
        return ((x[0::2, 0::2] + x[0::2, 1::2] + x[1::2, 0::2] + x[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    sx, fx = _axis_tables(w, dst_w)
    sy, fy = _axis_tables(h, dst_h)
    # x: clamp with weight reset (resize.cpp: "if( sx < 0 ) fx = 0, sx = 0; if( sx >= ssize.width-1 ) fx = 0, sx = width-1")
    lo = sx < 0
    fx = np.where(lo, np.float32(0), fx); sx = np.where(lo, 0, sx)
    hi = sx >= w - 1
    fx = np.where(hi, np.float32(0), fx); sx = np.where(hi, w - 1, sx)
    a1 = np.rint(fx * np.float32(2048)).astype(np.int32)          # saturate_cast<short>: round half to even
    a0 = np.rint((np.float32(1) - fx) * np.float32(2048)).astype(np.int32)
    sx1 = np.minimum(sx + 1, w - 1)
    # y: rows are clipped, weights are NOT reset
    b1 = np.rint(fy * np.float32(2048)).astype(np.int32)
    b0 = np.rint((np.float32(1) - fy) * np.float32(2048)).astype(np.int32)
    y0 = np.clip(sy, 0, h - 1); y1 = np.clip(sy + 1, 0, h - 1)
    src = img.astype(np.int32)
    hrow = src[:, sx] * a0[None, :, None] + src[:, sx1] * a1[None, :, None]      # [h, dst_w, c]
    r0 = hrow[y0]; r1 = hrow[y1]
    out = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def crop_batch(frame_bgr, boxes, margin=True):
    """All heads of one frame -> (M,224,224,3) RGB uint8, demo_video.py:13-24 without the drawing side effects."""
    h, w = frame_bgr.shape[:2]
    out = []
    for b in boxes:
        if margin:
            y0, y1, x0, x1 = enlarge_box(b, h, w)
        else:
            y0, x0, y1, x1 = [int(v) for v in b]
        crop = frame_bgr[y0:y1, x0:x1][:, :, ::-1]               # BGR -> RGB (demo_video.py:22)
        out.append(resize_linear_u8(crop))
    return np.stack(out) if out else np.zeros((0, 224, 224, 3), np.uint8)
