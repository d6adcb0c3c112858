"""Detections of one image as a lazily materialised list of ``LabeledBoundingBox``.

The reference returns ``List[LabeledBoundingBox]`` per image (`app/inference_cascade_app.py:319-334`,
`app/inference_app.py:203-212`).  The device path gets its detections as arrays; building thousands of Python
objects per batch costs more than the D2H copy of the arrays, so the objects are created on FIRST USE:
``BoxList`` is a ``Sequence`` that behaves like that list (len / iteration / indexing / ``==`` / ``+`` / any
other list method through the materialised list) and additionally exposes the arrays (``rows()``) for
consumers that never need objects (``dist.run_sharded``, the FDDB writer).

Value types follow the reference: python ints + the float32 softmax before NMS (cascade `:319-334`), numpy
ints + ``float(weight)`` after ``cv2.groupRectangles`` (`app/inference_app.py:203-212`).
"""
from collections.abc import Sequence
from typing import List

import numpy as np

from .rectangles import LabeledBoundingBox


class BoxList(Sequence):
    __slots__ = ("_boxes", "_conf", "_images", "_label", "_post_nms", "_list")

    def __init__(self, boxes: np.ndarray, conf: np.ndarray, images, label, post_nms: bool):
        """``boxes`` [k,4] int32 (xmin, ymin, xmax, ymax), ``conf`` [k], ``images``: ONE image object (all boxes
        of one image) or a list of k objects (flat results over several images)."""
        self._boxes, self._conf, self._images, self._label, self._post_nms = boxes, conf, images, label, post_nms
        self._list = None

    # ------------------------------------------------------------------ arrays
    def rows(self) -> np.ndarray:
        """``[k,5]`` float64 rows ``(xmin, ymin, xmax, ymax, confidence)`` without creating any object."""
        if self._list is not None:
            return np.asarray([[b.xmin, b.ymin, b.xmax, b.ymax, b.confidence] for b in self._list], np.float64).reshape(-1, 5)
        out = np.empty((self._boxes.shape[0], 5), np.float64)
        out[:, :4] = self._boxes
        out[:, 4] = self._conf
        return out

    # ------------------------------------------------------------------ list behaviour
    def _mat(self) -> List[LabeledBoundingBox]:
        if self._list is None:
            new, cls, fg = LabeledBoundingBox.__new__, LabeledBoundingBox, self._label
            imgs = self._images
            per_box = isinstance(imgs, list)
            out = []
            bx = self._boxes
            if self._post_nms:        # numpy ints and float(weight) like `app/inference_app.py:203-212`
                cols = (list(bx[:, 0]), list(bx[:, 1]), list(bx[:, 2]), list(bx[:, 3]), np.asarray(self._conf, np.float64).tolist())
            else:                     # python ints and the float32 softmax like cascade `:319-334`
                cols = (bx[:, 0].tolist(), bx[:, 1].tolist(), bx[:, 2].tolist(), bx[:, 3].tolist(), list(self._conf))
            for j, (x0, y0, x1, y1, c) in enumerate(zip(*cols)):
                o = new(cls)
                o._xmin = x0; o._ymin = y0; o._xmax = x1; o._ymax = y1
                o._label = fg; o._confidence = c; o._image = imgs[j] if per_box else imgs
                out.append(o)
            self._list = out
        return self._list

    def __len__(self):
        return len(self._list) if self._list is not None else int(self._boxes.shape[0])

    def __getitem__(self, i):
        return self._mat()[i]

    def __iter__(self):
        return iter(self._mat())

    def __eq__(self, other):
        return self._mat() == (other._mat() if isinstance(other, BoxList) else other)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __add__(self, other):
        return self._mat() + list(other)

    def __radd__(self, other):
        return list(other) + self._mat()

    def __repr__(self):
        return repr(self._mat())

    def __getattr__(self, name):            # append / extend / sort / ... act on the materialised list
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(self._mat(), name)
