"""CPU restatement of the reference's overlay for the tests (TEST INFRASTRUCTURE ONLY, like whenet_oracle.py).

``draw_axis_ref`` restates reference utils.py:13-43 line by line; ``process_detection_ref`` restates demo_video.py:11-35
(including the draw-before-the-next-crop order of demo_video.py:57-58).  Golden check: the axis end points for
(yaw, pitch, roll) = (0, 0, 0) are the image X axis, the image Y axis and the centre itself (utils.py:28-38 by hand)."""
from math import cos, sin

import cv2
import numpy as np


def draw_axis_ref(img, yaw, pitch, roll, tdx=None, tdy=None, size=100):
    pitch = pitch * np.pi / 180                     # utils.py:15
    yaw = -(yaw * np.pi / 180)                      # utils.py:16
    roll = roll * np.pi / 180                       # utils.py:17
    if tdx != None and tdy != None:                 # utils.py:19-25  (noqa: E711 - the reference's own comparison)
        tdx = tdx
        tdy = tdy
    else:
        height, width = img.shape[:2]
        tdx = width / 2
        tdy = height / 2
    x1 = size * (cos(yaw) * cos(roll)) + tdx                                              # utils.py:28
    y1 = size * (cos(pitch) * sin(roll) + cos(roll) * sin(pitch) * sin(yaw)) + tdy       # utils.py:29
    x2 = size * (-cos(yaw) * sin(roll)) + tdx                                             # utils.py:33
    y2 = size * (cos(pitch) * cos(roll) - sin(pitch) * sin(yaw) * sin(roll)) + tdy       # utils.py:34
    x3 = size * (sin(yaw)) + tdx                                                          # utils.py:37
    y3 = size * (-cos(yaw) * sin(pitch)) + tdy                                            # utils.py:38
    cv2.line(img, (int(tdx), int(tdy)), (int(x1), int(y1)), (0, 0, 255), 2)             # utils.py:40-42
    cv2.line(img, (int(tdx), int(tdy)), (int(x2), int(y2)), (0, 255, 0), 2)
    cv2.line(img, (int(tdx), int(tdy)), (int(x3), int(y3)), (255, 0, 0), 2)
    return img


def process_detection_ref(model, img, bbox, display="simple"):
    y_min, x_min, y_max, x_max = bbox                                      # demo_video.py:13
    y_min = max(0, y_min - abs(y_min - y_max) / 10)                        # demo_video.py:15-19
    y_max = min(img.shape[0], y_max + abs(y_min - y_max) / 10)
    x_min = max(0, x_min - abs(x_min - x_max) / 5)
    x_max = min(img.shape[1], x_max + abs(x_min - x_max) / 5)
    x_max = min(x_max, img.shape[1])
    img_rgb = img[int(y_min):int(y_max), int(x_min):int(x_max)]            # demo_video.py:21-24
    img_rgb = cv2.cvtColor(img_rgb, cv2.COLOR_BGR2RGB)
    img_rgb = cv2.resize(img_rgb, (224, 224))
    img_rgb = np.expand_dims(img_rgb, axis=0)
    cv2.rectangle(img, (int(x_min), int(y_min)), (int(x_max), int(y_max)), (0, 0, 0), 2)     # demo_video.py:26
    yaw, pitch, roll = model.get_angle(img_rgb)                            # demo_video.py:27-28
    yaw, pitch, roll = np.squeeze([yaw, pitch, roll])
    draw_axis_ref(img, yaw, pitch, roll, tdx=(x_min + x_max) / 2, tdy=(y_min + y_max) / 2, size=abs(x_max - x_min) // 2)
    if display == "full":                                                  # demo_video.py:31-34
        cv2.putText(img, "yaw: {}".format(np.round(yaw)), (int(x_min), int(y_min)), cv2.FONT_HERSHEY_SIMPLEX, 0.4, (100, 255, 0), 1)
        cv2.putText(img, "pitch: {}".format(np.round(pitch)), (int(x_min), int(y_min) - 15), cv2.FONT_HERSHEY_SIMPLEX, 0.4, (100, 255, 0), 1)
        cv2.putText(img, "roll: {}".format(np.round(roll)), (int(x_min), int(y_min) - 30), cv2.FONT_HERSHEY_SIMPLEX, 0.4, (100, 255, 0), 1)
    return img, (yaw, pitch, roll)
