"""Host side of the stream path's output: the pose-axis overlay and frame annotation of the reference demos
(SURVEY.md section 8f-4).  Pure host code on top of OpenCV's drawing primitives, exactly the calls the reference makes,
so annotated frames are pixel-identical to the reference's for the same angles.

  ``axis_endpoints`` / ``draw_axis``   reference utils.py:13-43
  ``annotate_head``                    reference demo_video.py:25-34  (rectangle, axes, optional yaw/pitch/roll text)
  ``process_frame``                    reference demo_video.py:11-35,57-58 for all heads of one frame

``process_frame(reference_order=True)`` reproduces the reference's order of operations bit for bit: it annotates the frame
after EACH head and cuts the next head's crop from the already annotated frame (demo_video.py:57-58 calls
process_detection head by head on the same array), one forward per head.  ``reference_order=False`` (the fast path) cuts
every crop from the CLEAN frame on the GPU in one batch (``WHENet.get_angle_from_frame``) and annotates afterwards; crops that
overlap an earlier head's rectangle / axes / text then differ from the reference's by those drawn pixels - a deliberate
deviation (the drawn overlay is not image content), documented in DESIGN.md.
"""
from __future__ import annotations

from math import cos, sin

import numpy as np

from . import crops as _crops


def axis_endpoints(yaw, pitch, roll, tdx, tdy, size):
    """End points (x1,y1) red X axis, (x2,y2) green Y axis, (x3,y3) blue Z axis; angles in degrees (utils.py:15-38)."""
    pitch = pitch * np.pi / 180
    yaw = -(yaw * np.pi / 180)
    roll = roll * np.pi / 180
    x1 = size * (cos(yaw) * cos(roll)) + tdx
    y1 = size * (cos(pitch) * sin(roll) + cos(roll) * sin(pitch) * sin(yaw)) + tdy
    x2 = size * (-cos(yaw) * sin(roll)) + tdx
    y2 = size * (cos(pitch) * cos(roll) - sin(pitch) * sin(yaw) * sin(roll)) + tdy
    x3 = size * (sin(yaw)) + tdx
    y3 = size * (-cos(yaw) * sin(pitch)) + tdy
    return (x1, y1), (x2, y2), (x3, y3)


def draw_axis(img, yaw, pitch, roll, tdx=None, tdy=None, size=100):
    """reference utils.py:13-43: draws on ``img`` in place (BGR) and returns it."""
    import cv2
    if tdx is None or tdy is None:
        height, width = img.shape[:2]
        tdx, tdy = width / 2, height / 2
    (x1, y1), (x2, y2), (x3, y3) = axis_endpoints(float(yaw), float(pitch), float(roll), tdx, tdy, size)
    cv2.line(img, (int(tdx), int(tdy)), (int(x1), int(y1)), (0, 0, 255), 2)
    cv2.line(img, (int(tdx), int(tdy)), (int(x2), int(y2)), (0, 255, 0), 2)
    cv2.line(img, (int(tdx), int(tdy)), (int(x3), int(y3)), (255, 0, 0), 2)
    return img


def annotate_head(img, bounds, yaw, pitch, roll, display: str = "simple"):
    """What demo_video.py:25-34 draws for one head.  ``bounds`` = the margin-enlarged FLOAT bounds
    (y_min, y_max, x_min, x_max) of demo_video.py:15-19 (their fractional parts enter tdx/tdy/size before truncation)."""
    import cv2
    y_min, y_max, x_min, x_max = bounds
    cv2.rectangle(img, (int(x_min), int(y_min)), (int(x_max), int(y_max)), (0, 0, 0), 2)
    draw_axis(img, yaw, pitch, roll, tdx=(x_min + x_max) / 2, tdy=(y_min + y_max) / 2, size=abs(x_max - x_min) // 2)
    if display == "full":
        f, c = cv2.FONT_HERSHEY_SIMPLEX, (100, 255, 0)
        cv2.putText(img, "yaw: {}".format(np.round(yaw)), (int(x_min), int(y_min)), f, 0.4, c, 1)
        cv2.putText(img, "pitch: {}".format(np.round(pitch)), (int(x_min), int(y_min) - 15), f, 0.4, c, 1)
        cv2.putText(img, "roll: {}".format(np.round(roll)), (int(x_min), int(y_min) - 30), f, 0.4, c, 1)
    return img


def process_frame(model, frame, boxes, display: str = "simple", reference_order: bool = True):
    """All detections of one BGR frame: angles + annotation, in place.  Returns ``(frame, yaw, pitch, roll)``.
    ``model`` is anything with ``get_angle`` (and ``get_angle_from_frame`` for the batched path)."""
    import cv2
    n = len(boxes)
    yaw, pitch, roll = (np.zeros((n,), np.float32) for _ in range(3))
    if n == 0:
        return frame, yaw, pitch, roll
    bounds = [_crops.enlarge_bounds(b, frame.shape[0], frame.shape[1]) for b in boxes]
    if reference_order:
        for i, (y0, y1, x0, x1) in enumerate(bounds):
            crop = frame[int(y0):int(y1), int(x0):int(x1)]
            crop = cv2.resize(cv2.cvtColor(crop, cv2.COLOR_BGR2RGB), (224, 224))
            y, p, r = model.get_angle(np.expand_dims(crop, axis=0))
            yaw[i], pitch[i], roll[i] = np.squeeze([y, p, r])
            annotate_head(frame, bounds[i], yaw[i], pitch[i], roll[i], display)
    else:
        y, p, r = model.get_angle_from_frame(frame, boxes, margin=True)
        yaw[:], pitch[:], roll[:] = y, p, r
        for i in range(n):
            annotate_head(frame, bounds[i], yaw[i], pitch[i], roll[i], display)
    return frame, yaw, pitch, roll
