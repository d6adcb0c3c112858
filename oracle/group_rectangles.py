"""Oracle (test infrastructure): restatement of
``cv2.groupRectangles(rects, groupThreshold, eps=0.2)``.

Reference call site: `app/inference_app.py:200-201` (``_run_nms``; confidence
of a merged box = cluster size, `:206-212`).  OpenCV is an un-vendored
dependency (pinned 3.3.1); this restates the published algorithm of
``objdetect/cascadedetect.cpp`` (``groupRectangles`` + ``cv::partition`` with
the ``SimilarRects`` predicate).  Pinned against the installed OpenCV,
including output order, in ``tests/test_oracle_group_rectangles.py``.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def _cv_round_f32(x: np.float32) -> int:
    """``cvRound(float)``: round half to even."""
    return int(np.rint(np.float32(x)))


def similar_matrix(r: np.ndarray, eps: float = 0.2) -> np.ndarray:
    """``SimilarRects``: all four edge deltas <= eps*(min(w)+min(h))*0.5 (double)."""
    r = r.astype(np.int64)
    x, y, w, h = r[:, 0], r[:, 1], r[:, 2], r[:, 3]
    delta = eps * (np.minimum(w[:, None], w[None, :]) + np.minimum(h[:, None], h[None, :])).astype(np.float64) * 0.5
    return ((np.abs(x[:, None] - x[None, :]) <= delta) &
            (np.abs(y[:, None] - y[None, :]) <= delta) &
            (np.abs((x + w)[:, None] - (x + w)[None, :]) <= delta) &
            (np.abs((y + h)[:, None] - (y + h)[None, :]) <= delta))


def partition_labels(sim: np.ndarray) -> Tuple[np.ndarray, int]:
    """``cv::partition``: connected components of the similarity graph; classes are
    numbered in order of their first member."""
    n = sim.shape[0]
    parent = np.arange(n)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    ii, jj = np.nonzero(np.triu(sim, 1))
    for a, b in zip(ii.tolist(), jj.tolist()):
        ra, rb = find(a), find(b)
        if ra != rb:
            if ra < rb:
                parent[rb] = ra
            else:
                parent[ra] = rb
    labels = np.empty(n, np.int64)
    cls = {}
    for i in range(n):
        root = find(i)
        if root not in cls:
            cls[root] = len(cls)
        labels[i] = cls[root]
    return labels, len(cls)


def group_rectangles(rects, group_threshold: int, eps: float = 0.2):
    """Returns ``(rects int32 [K',4] as x,y,w,h, weights int32 [K'])``."""
    r = np.asarray(rects, np.int32).reshape(-1, 4)
    n = r.shape[0]
    if group_threshold <= 0 or n == 0:
        return r.copy(), np.ones(n, np.int32)
    labels, ncls = partition_labels(similar_matrix(r, eps))
    sums = np.zeros((ncls, 4), np.int64)
    cnt = np.zeros(ncls, np.int64)
    np.add.at(sums, labels, r.astype(np.int64))
    np.add.at(cnt, labels, 1)
    rr = np.zeros((ncls, 4), np.int64)
    for i in range(ncls):
        s = np.float32(1.0) / np.float32(cnt[i])
        for k in range(4):
            # saturate_cast<int>(int * float): the int is converted to float32 first
            rr[i, k] = _cv_round_f32(np.float32(np.float32(np.int32(sums[i, k])) * s))
    out_r, out_w = [], []
    for i in range(ncls):
        n1 = int(cnt[i])
        if n1 <= group_threshold:
            continue
        r1 = rr[i]
        nested = False
        for j in range(ncls):
            n2 = int(cnt[j])
            if j == i or n2 <= group_threshold:
                continue
            r2 = rr[j]
            dx = int(np.rint(np.float64(r2[2]) * eps))               # saturate_cast<int>(double) = cvRound
            dy = int(np.rint(np.float64(r2[3]) * eps))
            if (r1[0] >= r2[0] - dx and r1[1] >= r2[1] - dy and
                    r1[0] + r1[2] <= r2[0] + r2[2] + dx and
                    r1[1] + r1[3] <= r2[1] + r2[3] + dy and
                    (n2 > max(3, n1) or n1 < 3)):
                nested = True
                break
        if not nested:
            out_r.append(r1)
            out_w.append(n1)
    return (np.asarray(out_r, np.int32).reshape(-1, 4), np.asarray(out_w, np.int32))
