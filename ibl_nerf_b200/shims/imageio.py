"""Minimal `imageio` stand-in (imwrite / imread through OpenCV) for the reference's PNG export."""
import cv2
import numpy as np


def imwrite(path, img):
    img = np.asarray(img)
    if img.ndim == 3 and img.shape[-1] == 3:
        img = img[..., ::-1]
    elif img.ndim == 3 and img.shape[-1] == 4:
        img = img[..., [2, 1, 0, 3]]
    cv2.imwrite(str(path), img)


def imread(path, pilmode=None, **kwargs):
    """`pilmode="RGB"` (dataset_mitsuba.py:38) forces three channels."""
    img = cv2.imread(str(path), cv2.IMREAD_COLOR if pilmode == "RGB" else cv2.IMREAD_UNCHANGED)
    if img is not None and img.ndim == 3:
        img = img[..., ::-1] if img.shape[-1] == 3 else img[..., [2, 1, 0, 3]]
    return img
