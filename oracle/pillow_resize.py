"""Oracle (test infrastructure): integer-exact restatement of Pillow's
``Image.resize(size, BILINEAR)`` for 8-bit RGB.

The reference reaches it through ``scipy.misc.imresize(arr, ratio)``
(`data/image_info.py:240`), which for a float ``ratio`` resizes to
``(int(W*ratio), int(H*ratio))`` with ``resample=BILINEAR``.  Pillow is an
un-vendored dependency (pinned 5.1.0 in `requirements.txt`); this restates the
published algorithm of ``libImaging/Resample.c`` (``precompute_coeffs``,
``normalize_coeffs_8bpc``, ``ImagingResampleHorizontal_8bpc`` /
``Vertical_8bpc``): antialiased triangle filter whose support grows with the
downscale factor, 22-bit fixed-point coefficients, horizontal pass -> uint8 ->
vertical pass -> uint8.  Pinned bit-exact against the installed Pillow in
``tests/test_oracle_resize.py``.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2   # 22


def _ctrunc(x: float) -> int:
    """C ``(int)`` cast of a double: truncation toward zero."""
    return int(x)


def bilinear_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray, int]:
    """``precompute_coeffs`` + ``normalize_coeffs_8bpc`` for the full-image box.

    Returns ``(xmin[out], xcount[out], kk[out, ksize] int32, ksize)``.
    """
    scale = float(in_size) / float(out_size)
    filterscale = scale if scale >= 1.0 else 1.0
    support = 1.0 * filterscale                    # bilinear filter support = 1.0
    ksize = int(math.ceil(support)) * 2 + 1
    xmin_a = np.zeros(out_size, np.int32)
    xcnt_a = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = _ctrunc(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = _ctrunc(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        ws = []
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            w = 1.0 - a if a < 1.0 else 0.0
            ws.append(w)
            ww += w
        for x in range(xmax):
            w = ws[x]
            if ww != 0.0:
                w /= ww
            # normalize_coeffs_8bpc: round half away from zero via +-0.5 then (int)
            v = w * (1 << PRECISION_BITS)
            kk[xx, x] = _ctrunc(-0.5 + v) if w < 0 else _ctrunc(0.5 + v)
        xmin_a[xx] = xmin
        xcnt_a[xx] = xmax
    return xmin_a, xcnt_a, kk, ksize


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resample_axis(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One pass along ``axis`` (0 = vertical, 1 = horizontal) of ``[H,W,C]`` uint8."""
    in_size = img.shape[axis]
    xmin, xcnt, kk, ksize = bilinear_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)            # [in, other, C]
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    half = 1 << (PRECISION_BITS - 1)
    for xx in range(out_size):
        n = int(xcnt[xx])
        seg = src[xmin[xx]:xmin[xx] + n]
        acc = half + np.tensordot(kk[xx, :n].astype(np.int64), seg, axes=(0, 0))
        out[xx] = _clip8(acc)
    return np.moveaxis(out, 0, axis)


def resize_bilinear(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """``Image.fromarray(img).resize((out_w, out_h), BILINEAR)`` for ``[H,W,3]`` uint8.
    Horizontal pass first; a pass is skipped when that dimension is unchanged."""
    h, w = img.shape[:2]
    out = img
    if out_w != w:
        out = resample_axis(out, out_w, 1)
    if out_h != h:
        out = resample_axis(out, out_h, 0)
    return np.ascontiguousarray(out)


def imresize_ratio(img: np.ndarray, ratio: float) -> np.ndarray:
    """``scipy.misc.imresize(img, float_ratio)`` (SciPy 1.0.1 semantics):
    size = ``(array(im.size) * ratio).astype(int)`` = ``(int(W*r), int(H*r))``."""
    h, w = img.shape[:2]
    return resize_bilinear(img, int(w * ratio), int(h * ratio))
