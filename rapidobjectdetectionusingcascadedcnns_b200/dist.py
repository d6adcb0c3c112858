"""Multi-GPU: image-sharded replicas, one process per GPU.

Frames are independent through every step of the path, so frame ``j`` goes to rank
``j mod world_size``, every rank holds all stage weights and runs the whole cascade, and the
only exchange is the final gather of detections: ONE all-gather of a fixed-capacity block per rank --
row 0 carries ``(count, error flag)``, rows 1.. the detections ``(frame, xmin, ymin, xmax, ymax, confidence)``.
The capacity grows (on every rank alike, because every rank sees every count) only when a batch does not fit.
No collective runs inside the cascade.  Works with NCCL (GPU tensors) and gloo (CPU tests).
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

_CAPACITY = 4096          # detection rows per rank and exchange (grows when needed, same value on all ranks)


class ShardError(RuntimeError):
    """A rank failed inside ``run_sharded``; raised on EVERY rank after the exchange so that nobody hangs."""


def shard_indices(n_items: int, rank: int, world_size: int) -> List[int]:
    """Frame ``j`` -> rank ``j mod world_size``."""
    return list(range(rank, n_items, world_size))


def bind_rank_to_cores(local_rank: int, local_world: int) -> Optional[List[int]]:
    """Give this process (and the threads it starts later: copy pool, decode pool) its own slice of the host
    cores instead of letting every rank float over the same set.  Returns the slice, or None when the
    platform has no affinity API or fewer than two cores per rank."""
    if not hasattr(os, "sched_getaffinity") or local_world < 2:
        return None
    avail = sorted(os.sched_getaffinity(0))
    per = len(avail) // local_world
    if per < 2:
        return None
    mine = avail[local_rank * per:(local_rank + 1) * per]
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return None
    return mine


def gather_detections(rows: np.ndarray, device=None, error: bool = False) -> np.ndarray:
    """``rows`` ``[k,6]`` float64 with GLOBAL frame ids in column 0.  Returns the concatenation over
    ranks sorted by frame id (stable, so the per-frame class order is preserved) on every rank.
    ``error=True`` marks this rank as failed: every rank then raises ``ShardError`` after the exchange."""
    global _CAPACITY
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        if error:
            raise ShardError("the local run failed")
        return rows
    ws = dist.get_world_size()
    backend = dist.get_backend()
    dev = torch.device("cpu") if backend == "gloo" else (device or torch.device("cuda", torch.cuda.current_device()))
    rows = np.ascontiguousarray(rows, np.float64).reshape(-1, 6)
    k = rows.shape[0]
    while True:
        cap = _CAPACITY
        block = np.zeros((cap + 1, 6), np.float64)
        block[0, 0], block[0, 1] = k, 1.0 if error else 0.0
        block[1:1 + min(k, cap)] = rows[:cap]
        mine = torch.from_numpy(block)
        if dev.type == "cuda":
            mine = mine.pin_memory().to(dev, non_blocking=True)
        allbuf = torch.empty((ws * (cap + 1), 6), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allbuf, mine)                 # the one collective of the step
        host = allbuf.cpu().numpy().reshape(ws, cap + 1, 6)       # the one device -> host read
        counts = host[:, 0, 0].astype(np.int64)
        if (host[:, 0, 1] != 0).any():
            bad = [int(r) for r in np.nonzero(host[:, 0, 1])[0]]
            raise ShardError("rank(s) %s failed inside the sharded run" % bad)
        if counts.max() <= cap:
            break
        _CAPACITY = int(2 ** np.ceil(np.log2(counts.max())))      # every rank computes the same new capacity
    parts = [host[r, 1:1 + c] for r, c in enumerate(counts.tolist())]
    allrows = np.concatenate(parts) if parts else np.zeros((0, 6))
    order = np.argsort(allrows[:, 0], kind="stable")
    return allrows[order]


def _rows_of(per_image, ids: Sequence[int]) -> np.ndarray:
    """Detections of this rank's images as ``[k,6]`` rows with global frame ids (no box objects are created
    when the app returned ``BoxList`` views)."""
    parts = []
    for gid, per in zip(ids, per_image):
        if len(per) == 0:
            continue
        r5 = per.rows() if hasattr(per, "rows") else np.asarray(
            [[b.xmin, b.ymin, b.xmax, b.ymax, b.confidence] for b in per], np.float64).reshape(-1, 5)
        parts.append(np.column_stack([np.full((r5.shape[0],), float(gid)), r5]))
    return np.concatenate(parts) if parts else np.zeros((0, 6), np.float64)


def run_sharded(app, images: Sequence, merge: bool = True, as_rows: bool = False):
    """Run ``app.run_inference_on_images`` on this rank's shard (image ``j`` -> rank ``j mod world_size``) and
    gather all detections.  Returns, on every rank, ``[n_images]`` lists of ``(xmin, ymin, xmax, ymax,
    confidence)`` tuples -- or the sorted ``[K,6]`` row array with ``as_rows=True``.

    A rank whose local run raises (e.g. the reference's ``ValueError`` when none of its images yields a window,
    `app/abstract_inference_app.py:106-107`) still enters the collective with an error flag; every rank then
    raises ``ShardError`` instead of the healthy ranks waiting for ever."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    ws = dist.get_world_size() if dist.is_initialized() else 1
    mine = shard_indices(len(images), rank, ws)
    rows, failure = np.zeros((0, 6), np.float64), None
    try:
        if mine and merge and hasattr(app, "run_inference_rows"):
            rows = app.run_inference_rows([images[j] for j in mine])        # arrays end to end, no box objects
            rows[:, 0] = np.asarray(mine, np.float64)[rows[:, 0].astype(np.int64)]
        else:
            local = app.run_inference_on_images([images[j] for j in mine], merge=merge) if mine else []
            rows = _rows_of(local, mine)
    except Exception as e:           # noqa: BLE001 -- the flag must reach the collective whatever happened
        failure = e
    try:
        allrows = gather_detections(rows, error=failure is not None)
    except ShardError as e:
        if failure is not None:
            raise ShardError("%s (local error on rank %d: %r)" % (e, rank, failure)) from failure
        raise
    if as_rows:
        return allrows
    out = [[] for _ in images]
    for r in allrows:
        out[int(r[0])].append((r[1], r[2], r[3], r[4], r[5]))
    return out
