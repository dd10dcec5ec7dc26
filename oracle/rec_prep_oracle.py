"""TEST INFRASTRUCTURE ONLY (oracle). The recogniser pre-processing of the reference's deploy loop as its own sequence of
cv2 / numpy calls: R/deploy/pytorch/run_ocr.py:212-220 (cvtColor by rec_img_mode) and
R/pytocr/data/imaug/rec_img_aug.py:108-134 (`resize_norm_img` as RecResizeImg calls it: padding=True, resized_w=None).
The checker of csrc/prep.cuh (tests/test_geometry_host.py, bit for bit) and of deploy/run_ocr.py (tests/test_run_ocr_gpu.py)."""
import math

import cv2
import numpy as np


def rec_preprocess(part_img, img_mode, image_shape):
    if img_mode == "GRAY":
        img = cv2.cvtColor(part_img, cv2.COLOR_BGR2GRAY)
    elif img_mode == "RGB":
        img = cv2.cvtColor(part_img, cv2.COLOR_BGR2RGB)
    else:
        img = part_img.copy()
    imgC, imgH, imgW = image_shape
    h, w = img.shape[:2]
    ratio = w / float(h)
    if math.ceil(imgH * ratio) > imgW:
        resized_w = imgW
    else:
        resized_w = int(math.ceil(imgH * ratio))
    resized_image = cv2.resize(img, (resized_w, imgH))
    resized_image = resized_image.astype("float32")
    if image_shape[0] == 1 and len(img.shape) == 2:
        resized_image = resized_image / 255
        resized_image = resized_image[np.newaxis, :]
    else:
        resized_image = resized_image.transpose((2, 0, 1)) / 255
    resized_image -= 0.5
    resized_image /= 0.5
    padding_im = np.zeros((imgC, imgH, imgW), dtype=np.float32)
    padding_im[:, :, 0:resized_w] = resized_image
    return padding_im
