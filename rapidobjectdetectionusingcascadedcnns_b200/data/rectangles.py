"""Box semantics of the hot path (mirror of `data/rectangles.py:12-111,161-220,223-307`).

Half-open pixel rectangles with the origin in the top-left corner; ``Window`` coordinates live
in a scaled pyramid level, the ``*_norm`` properties map them back to the original frame with
``int(coord / scale)`` (each edge truncated on its own, so norm boxes may be non-square by one
pixel).  Window *extraction* itself happens on the GPU (librodc planner + K1); these classes
exist so that results and the ``run_inference_on_windows`` entry point look exactly like the
reference's.
"""
from typing import List


class Rectangle:
    __slots__ = ("_xmin", "_ymin", "_xmax", "_ymax")

    def __init__(self, xmin, ymin, xmax, ymax):
        self._xmin, self._ymin, self._xmax, self._ymax = xmin, ymin, xmax, ymax

    @property
    def xmin(self):
        return self._xmin

    @property
    def ymin(self):
        return self._ymin

    @property
    def xmax(self):
        return self._xmax

    @property
    def ymax(self):
        return self._ymax

    @property
    def width(self):
        return self.xmax - self.xmin

    @property
    def height(self):
        return self.ymax - self.ymin

    @property
    def is_valid(self):
        return None not in [self.xmin, self.ymin, self.xmax, self.ymax]

    def intersects(self, other: "Rectangle") -> bool:
        """`data/rectangles.py:61-80` (touching edges count as intersecting)."""
        return not (other.xmin > self.xmax or other.ymin > self.ymax or
                    self.xmin > other.xmax or self.ymin > other.ymax)

    def intersection_over_union(self, other: "Rectangle") -> float:
        """`data/rectangles.py:82-111` (inclusive-pixel convention, +1 on every extent)."""
        x_a, y_a = max(self.xmin, other.xmin), max(self.ymin, other.ymin)
        x_b, y_b = min(self.xmax, other.xmax), min(self.ymax, other.ymax)
        inter = (x_b - x_a + 1) * (y_b - y_a + 1)
        own = (self.xmax - self.xmin + 1) * (self.ymax - self.ymin + 1)
        oth = (other.xmax - other.xmin + 1) * (other.ymax - other.ymin + 1)
        return inter / float(own + oth - inter)


class BoundingBox(Rectangle):
    __slots__ = ()


class LabeledBoundingBox(BoundingBox):
    __slots__ = ("_xmin", "_ymin", "_xmax", "_ymax", "_label", "_confidence", "_image")

    def __init__(self, xmin, ymin, xmax, ymax, label, confidence=0.0, image=None):
        # flat initialisation (thousands of boxes per batch are built on the host)
        self._xmin = xmin; self._ymin = ymin; self._xmax = xmax; self._ymax = ymax
        self._label = label; self._confidence = confidence; self._image = image

    @property
    def label(self):
        return self._label

    @property
    def confidence(self) -> float:
        return self._confidence

    @property
    def image(self):
        return self._image

    @staticmethod
    def vertically_enlarge_bboxes(original_bboxes: List["LabeledBoundingBox"], enlarge_top=0.2,
                                  enlarge_bottom=None) -> List["LabeledBoundingBox"]:
        """`data/rectangles.py:185-220`: ``ymin' = max(ymin - 0.2 * height, 0)`` (float result);
        the bottom edge moves only when ``enlarge_bottom`` is given."""
        out = []
        for b in original_bboxes:
            ymin_new = max(b.ymin - (enlarge_top * b.height), 0) if enlarge_top is not None else b.ymin
            if enlarge_bottom is not None:
                ymax_new = min(b.ymax + (enlarge_bottom * b.height), b.image.img_height_original - 1)
            else:
                ymax_new = b.ymax
            out.append(LabeledBoundingBox(b.xmin, ymin_new, b.xmax, ymax_new, b.label, b.confidence, b.image))
        return out


class Window(Rectangle):
    __slots__ = ("_image", "_scale")

    def __init__(self, xmin, ymin, xmax, ymax, image, scale=1.0):
        self._image, self._scale = image, scale
        Rectangle.__init__(self, xmin, ymin, xmax, ymax)

    @property
    def scale(self):
        return self._scale

    @property
    def image(self):
        return self._image

    @property
    def xmin_norm(self):
        return int(self.xmin / self.scale)

    @property
    def ymin_norm(self):
        return int(self.ymin / self.scale)

    @property
    def xmax_norm(self):
        return int(self.xmax / self.scale)

    @property
    def ymax_norm(self):
        return int(self.ymax / self.scale)

    @property
    def width_norm(self):
        return int(self.xmax_norm - self.xmin_norm)

    @property
    def height_norm(self):
        return int(self.ymax_norm - self.ymin_norm)

    @classmethod
    def extract_windows(cls, img, convert_raw_to_np=True):
        """All sliding windows of ``img`` (`data/rectangles.py:309-435`): returns ``(windows_raw,
        windows_info)`` in the reference's order (level, x outer, y inner) with ``cf img_width`` as the
        window size.  The pyramid and the crops come from the GPU (librodc planner + K1 + window gather,
        bit-exact with the reference's Pillow path); raises ValueError like ``Window.raw`` when a window
        leaves its level."""
        import torch
        from .. import config as cf
        from .. import native
        frame = img.raw_original(True)
        win = int(cf.get("img_width"))
        plan = native.Plan(frame.shape[1], frame.shape[0], win, float(cf.get("window_scale_factor")),
                           float(cf.get("min_window_length")))
        if plan.n_windows and not plan.fits:
            raise ValueError("xmax must not leave the image boundaries")
        infos = [cls(int(x), int(y), int(x) + win, int(y) + win, img, float(plan.scale[l]))
                 for l, x, y in zip(plan.win_level.tolist(), plan.win_x.tolist(), plan.win_y.tolist())]
        if plan.n_windows:
            d = torch.from_numpy(frame[None].copy()).cuda()
            raw = native.extract_windows(plan, native.pyramid(plan, d)).cpu().numpy()
        else:
            import numpy as np
            raw = np.zeros((0, win, win, 3), np.uint8)
        return (raw if convert_raw_to_np else list(raw)), infos

    @property
    def raw_norm(self):
        """Unscaled pixels of the receptive field in the original frame (`:293-307`)."""
        full = self.image.raw_original(True)
        if self.ymax_norm > full.shape[0]:
            raise ValueError("ymax_norm must not leave the image boundaries")
        if self.xmax_norm > full.shape[1]:
            raise ValueError("xmax_norm must not leave the image boundaries")
        return full[self.ymin_norm:self.ymax_norm, self.xmin_norm:self.xmax_norm]
