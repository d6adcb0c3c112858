"""Multi-GPU: image-sharded replicas, one process per GPU.

Frames are independent through every step of the path, so frame ``j`` goes to rank
``j mod world_size``, every rank holds all stage weights and runs the whole cascade, and the
only exchange is the final gather of detections: an all-gather of per-rank counts followed by
an all-gather of padded ``[Kmax, 6]`` rows ``(frame, xmin, ymin, xmax, ymax, confidence)``.
No collective runs inside the cascade.  Works with NCCL (GPU tensors) and gloo (CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world_size: int) -> List[int]:
    """Frame ``j`` -> rank ``j mod world_size``."""
    return list(range(rank, n_items, world_size))


def gather_detections(rows: np.ndarray, device=None) -> np.ndarray:
    """``rows`` ``[k,6]`` float64 with GLOBAL frame ids in column 0.  Returns the concatenation over
    ranks sorted by frame id (stable, so the per-frame class order is preserved) on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rows
    ws = dist.get_world_size()
    backend = dist.get_backend()
    dev = torch.device("cpu") if backend == "gloo" else (device or torch.device("cuda", torch.cuda.current_device()))
    # two collectives and two host reads in total: counts, then one padded [ws * Kmax, 6] block
    cnt = torch.tensor([rows.shape[0]], dtype=torch.int64, device=dev)
    counts_t = torch.empty((ws,), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts_t, cnt)
    counts = counts_t.cpu().tolist()
    kmax = max(max(counts), 1)
    pad = torch.zeros((kmax, 6), dtype=torch.float64, device=dev)
    if rows.shape[0]:
        pad[:rows.shape[0]] = torch.from_numpy(np.ascontiguousarray(rows, np.float64)).to(dev)
    allbuf = torch.empty((ws * kmax, 6), dtype=torch.float64, device=dev)     # concatenated along dim 0
    dist.all_gather_into_tensor(allbuf, pad)
    host = allbuf.cpu().numpy().reshape(ws, kmax, 6)
    parts = [host[r, :c] for r, c in enumerate(counts)]
    allrows = np.concatenate(parts) if parts else np.zeros((0, 6))
    order = np.argsort(allrows[:, 0], kind="stable")
    return allrows[order]


def run_sharded(app, images: Sequence, merge: bool = True):
    """Run ``app.run_inference_on_images`` on this rank's shard and gather all detections.
    Returns ``[n_images]`` lists of ``(xmin, ymin, xmax, ymax, confidence)`` tuples on every rank."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    ws = dist.get_world_size() if dist.is_initialized() else 1
    mine = shard_indices(len(images), rank, ws)
    local = app.run_inference_on_images([images[j] for j in mine], merge=merge) if mine else []
    rows = [[mine[i], b.xmin, b.ymin, b.xmax, b.ymax, b.confidence] for i, per in enumerate(local) for b in per]
    allrows = gather_detections(np.asarray(rows, np.float64).reshape(-1, 6))
    out = [[] for _ in images]
    for r in allrows:
        out[int(r[0])].append((r[1], r[2], r[3], r[4], r[5]))
    return out
