"""Oracle (test infrastructure): integer-exact restatement of
``cv2.resize(src, (dw, dh))`` with the default ``INTER_LINEAR`` on 8UC3.

Reference call site: `app/inference_cascade_app.py:363-373` (per-survivor
re-crop to the next stage's input size).  OpenCV is an un-vendored dependency
(pinned 3.3.1); this restates the published algorithm of ``imgproc/resize.cpp``
(``resizeGeneric_`` with ``HResizeLinear`` / ``VResizeLinear`` 8u fixed point,
``INTER_RESIZE_COEF_BITS = 11``, and the ``INTER_AREA`` 2x fast path that
``INTER_LINEAR`` falls through to).  Pinned bit-exact against the installed
OpenCV in ``tests/test_oracle_resize.py``.
"""
from __future__ import annotations

import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def linear_taps(src: int, dst: int, clamp_frac: bool):
    """Per destination index: source index ``s`` and int16 weights ``(a0, a1)``.

    ``clamp_frac=True`` is the x-axis rule (fraction forced to 0 at the borders);
    ``False`` is the y-axis rule (fraction kept; the two row indices are clamped
    individually by the caller).
    """
    scale = 1.0 / (float(dst) / float(src))          # inv_scale = dst/src ; scale = 1/inv_scale (double)
    s_idx = np.zeros(dst, np.int64)
    a0 = np.zeros(dst, np.int64)
    a1 = np.zeros(dst, np.int64)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if clamp_frac:
            if s < 0:
                f, s = np.float32(0.0), 0
            if s >= src - 1:
                f, s = np.float32(0.0), src - 1
        s_idx[d] = s
        # saturate_cast<short>(x * 2048) : cvRound (round half to even) of float32 math
        a0[d] = int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(COEF_SCALE))))
        a1[d] = int(np.rint(np.float32(f * np.float32(COEF_SCALE))))
    return s_idx, a0, a1


def cv_resize_linear(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """``cv2.resize(src, (dw, dh))`` for ``[H,W,C]`` uint8."""
    sh, sw = src.shape[:2]
    if dw == sw and dh == sh:
        return src.copy()
    if sw == 2 * dw and sh == 2 * dh:                # INTER_AREA fast path (iscale_x == iscale_y == 2)
        s = src.astype(np.int32)
        out = (s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2
        return out.astype(np.uint8)
    xs, xa0, xa1 = linear_taps(sw, dw, True)
    ys, yb0, yb1 = linear_taps(sh, dh, False)
    s = src.astype(np.int64)
    x1 = np.minimum(xs + 1, sw - 1)
    # horizontal pass: int32 rows T[y, dx, c] = S[y, xs]*a0 + S[y, xs+1]*a1
    T = s[:, xs, :] * xa0[None, :, None] + s[:, x1, :] * xa1[None, :, None]
    y0 = np.clip(ys, 0, sh - 1)
    y1 = np.clip(ys + 1, 0, sh - 1)
    T0 = T[y0] >> 4
    T1 = T[y1] >> 4
    out = (((yb0[:, None, None] * T0) >> 16) + ((yb1[:, None, None] * T1) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)
