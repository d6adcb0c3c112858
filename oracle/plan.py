"""Oracle (test infrastructure): pyramid plan + sliding-window grid.

Restates `data/rectangles.py:309-435` (``Window.extract_windows``) and the
``*_norm`` properties `data/rectangles.py:248-270` in plain Python float64, in
the exact operation order of the reference (repeated division, float compares,
``int()`` truncation).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np


@dataclass
class Level:
    scale: float          # python float, result of repeated ``scale /= factor``
    width: int            # int(W * scale)  (scipy.misc.imresize size rule, `data/image_info.py:240`)
    height: int           # int(H * scale)
    step_x: int
    step_y: int
    nx: int               # grid columns
    ny: int               # grid rows
    first_window: int     # index of the level's first window in the table


@dataclass
class Plan:
    img_w: int
    img_h: int
    win: int
    factor: float
    min_window_length: float
    levels: List[Level] = field(default_factory=list)
    # window table, order: level (descending size) -> x outer -> y inner
    win_level: np.ndarray = None   # int32 [N]
    win_x: np.ndarray = None       # int32 [N]  xmin in the level
    win_y: np.ndarray = None       # int32 [N]
    norm: np.ndarray = None        # int32 [N,4] xmin_norm, ymin_norm, xmax_norm, ymax_norm

    @property
    def n_windows(self) -> int:
        return int(self.win_level.shape[0])


def plan_pyramid(img_w: int, img_h: int, win: int, factor: float = 1.1,
                 min_window_length: float = 0.075) -> Plan:
    """`data/rectangles.py:338-425`; all arithmetic in float64 as written there."""
    plan = Plan(img_w, img_h, win, factor, min_window_length)
    scale = 1.0
    iw = img_w                                                 # ints from shape[1]/shape[0]; float after the first division
    ih = img_h
    max_windows_per_row = 1.0 / min_window_length              # :359
    max_img_length = max_windows_per_row * win                 # :360
    lv, xs, ys, norms = [], [], [], []
    n = 0
    while True:
        if iw < max_img_length and ih < max_img_length:        # :366
            step_x = max(min(int(0.4 * win), int(0.1 * iw)), 1)   # :386
            step_y = max(min(int(0.4 * win), int(0.1 * ih)), 1)   # :387
            first = n
            nx = 0
            ny = 0
            xmin = -1 * step_x
            while True:
                xmin += step_x
                xmax = xmin + win
                if xmax >= iw:                                  # :395 (float compare)
                    break
                col = 0
                ymin = -1 * step_y
                while True:
                    ymin += step_y
                    ymax = ymin + win
                    if ymax >= ih:                              # :403
                        break
                    lv.append(len(plan.levels)); xs.append(xmin); ys.append(ymin)
                    norms.append((int(xmin / scale), int(ymin / scale),      # :248-262
                                  int(xmax / scale), int(ymax / scale)))
                    n += 1
                    col += 1
                ny = col
                nx += 1
            if scale != 1.0:
                lw, lh = int(img_w * scale), int(img_h * scale)   # imresize: (array(im.size)*ratio).astype(int)
            else:
                lw, lh = img_w, img_h                             # `data/image_info.py:209-213`
            plan.levels.append(Level(scale, lw, lh, step_x, step_y, nx, ny, first))
        scale /= factor                                         # :418
        ih /= factor                                            # :419
        iw /= factor                                            # :420
        if ih < win or iw < win:                                # :423
            break
    plan.win_level = np.asarray(lv, np.int32)
    plan.win_x = np.asarray(xs, np.int32)
    plan.win_y = np.asarray(ys, np.int32)
    plan.norm = np.asarray(norms, np.int32).reshape(-1, 4)
    return plan


def windows_fit(plan: Plan) -> bool:
    """``Window.raw`` raises ValueError when a window leaves its level
    (`data/rectangles.py:286-289`); the image is then skipped by
    `app/abstract_inference_app.py:165-183`."""
    for li, lvl in enumerate(plan.levels):
        m = plan.win_level == li
        if not m.any():
            continue
        if (plan.win_x[m] + plan.win > lvl.width).any() or (plan.win_y[m] + plan.win > lvl.height).any():
            return False
    return True
