"""Torch-tensor level wrappers over the C ABI (PyTorch = device memory + streams only).

Every function takes CUDA tensors, passes raw pointers + the current stream to librodc and
returns CUDA tensors; nothing here computes on the host and nothing falls back.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import IMPL_SIMT, IMPL_TENSOR, SRC_DENSE, SRC_GATHER, SRC_PYRAMID, SRC_RECROP, WindowSource, check


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.RodcError("librodc operates on CUDA tensors only (no CPU fallback)")


def device_info() -> Tuple[int, int, int]:
    a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
    check(_lib.lib().rodc_device_info(C.byref(a), C.byref(b), C.byref(c)), "rodc_device_info")
    return a.value, b.value, c.value


# --------------------------------------------------------------------------------- plan
class Plan:
    """Pyramid plan + window grid of one frame shape (host planner of librodc).

    Mirrors the scale loop / grid of ``Window.extract_windows`` (data/rectangles.py:338-425)."""

    def __init__(self, img_w: int, img_h: int, win: int, factor: float = 1.1, min_window_length: float = 0.075):
        L = _lib.lib()
        h = C.c_void_p()
        check(L.rodc_plan_create(int(img_w), int(img_h), int(win), float(factor), float(min_window_length), C.byref(h)),
              "rodc_plan_create")
        self._h = h
        self.img_w, self.img_h, self.win = int(img_w), int(img_h), int(win)
        self.factor, self.min_window_length = float(factor), float(min_window_length)
        self._read()

    @classmethod
    def from_scales(cls, img_w: int, img_h: int, scales: Sequence[float]) -> "Plan":
        """Pyramid of explicit scales without a window grid: the frames ``image.raw_scaled(True, next_scale)``
        of the resize-then-crop branch (`app/inference_cascade_app.py:374-394`)."""
        self = cls.__new__(cls)
        sc = np.ascontiguousarray(scales, np.float64)
        h = C.c_void_p()
        check(_lib.lib().rodc_plan_create_scales(int(img_w), int(img_h), int(sc.shape[0]), sc.ctypes.data_as(_lib.c_f64p),
                                                 C.byref(h)), "rodc_plan_create_scales")
        self._h = h
        self.img_w, self.img_h, self.win = int(img_w), int(img_h), 0
        self.factor, self.min_window_length = 0.0, 0.0
        self._read()
        return self

    def _read(self):
        L, h = _lib.lib(), self._h
        self.n_levels = L.rodc_plan_n_levels(h)
        self.n_windows = int(L.rodc_plan_n_windows(h))
        self.fits = bool(L.rodc_plan_fits(h))
        self.pyramid_bytes = int(L.rodc_plan_pyramid_bytes(h))
        self.scratch_bytes = int(L.rodc_plan_scratch_bytes(h))
        nl, nw = max(self.n_levels, 1), max(self.n_windows, 1)
        self.scale = np.zeros(nl, np.float64)
        self.level_w = np.zeros(nl, np.int32); self.level_h = np.zeros(nl, np.int32)
        self.step_x = np.zeros(nl, np.int32); self.step_y = np.zeros(nl, np.int32)
        self.nx = np.zeros(nl, np.int32); self.ny = np.zeros(nl, np.int32)
        self.first_window = np.zeros(nl, np.int64); self.byte_offset = np.zeros(nl, np.int64)
        p32 = lambda a: a.ctypes.data_as(_lib.c_i32p)
        check(L.rodc_plan_levels(h, self.scale.ctypes.data_as(_lib.c_f64p), p32(self.level_w), p32(self.level_h),
                                 p32(self.step_x), p32(self.step_y), p32(self.nx), p32(self.ny),
                                 self.first_window.ctypes.data_as(_lib.c_i64p),
                                 self.byte_offset.ctypes.data_as(_lib.c_i64p)), "rodc_plan_levels")
        self.win_level = np.zeros(nw, np.int32); self.win_x = np.zeros(nw, np.int32); self.win_y = np.zeros(nw, np.int32)
        self.norm = np.zeros((nw, 4), np.int32)
        check(L.rodc_plan_windows(h, p32(self.win_level), p32(self.win_x), p32(self.win_y), p32(self.norm)),
              "rodc_plan_windows")
        for name in ("scale", "level_w", "level_h", "step_x", "step_y", "nx", "ny", "first_window", "byte_offset"):
            setattr(self, name, getattr(self, name)[:self.n_levels])
        self.win_level, self.win_x, self.win_y = (a[:self.n_windows] for a in (self.win_level, self.win_x, self.win_y))
        self.norm = self.norm[:self.n_windows]

    @property
    def handle(self):
        return self._h

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().rodc_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass


def pyramid(plan: Plan, frames: torch.Tensor, out: Optional[torch.Tensor] = None,
            scratch: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K1.  frames ``[n,H,W,3]`` u8 CUDA -> packed pyramid ``[n, plan.pyramid_bytes]`` u8."""
    _require_cuda(frames)
    assert frames.dtype == torch.uint8 and frames.is_contiguous()
    n = frames.shape[0]
    assert tuple(frames.shape[1:]) == (plan.img_h, plan.img_w, 3), (frames.shape, plan.img_h, plan.img_w)
    if out is None:
        out = torch.empty((n, plan.pyramid_bytes), dtype=torch.uint8, device=frames.device)
    if scratch is None:
        scratch = torch.empty((n, plan.scratch_bytes), dtype=torch.uint8, device=frames.device)
    check(_lib.lib().rodc_pyramid(plan.handle, _ptr(frames), n, _ptr(out), _ptr(scratch), _stream()), "rodc_pyramid")
    return out


def pyramid_level(plan: Plan, pyr: torch.Tensor, frame: int, level: int) -> torch.Tensor:
    """View of one level ``[h,w,3]`` inside a packed pyramid (for tests)."""
    off = int(plan.byte_offset[level]); h = int(plan.level_h[level]); w = int(plan.level_w[level])
    return pyr[frame, off:off + h * w * 3].view(h, w, 3)


def extract_windows(plan: Plan, pyr: torch.Tensor) -> torch.Tensor:
    """Raw stage-0 windows ``[n*N, win, win, 3]`` u8 (``Window.raw`` for every grid position)."""
    _require_cuda(pyr)
    n = pyr.shape[0]
    out = torch.empty((n * plan.n_windows, plan.win, plan.win, 3), dtype=torch.uint8, device=pyr.device)
    check(_lib.lib().rodc_extract_windows(plan.handle, _ptr(pyr), n, _ptr(out), _stream()), "rodc_extract_windows")
    return out


def window_meta(plan: Plan, n_frames: int, frame_id_base: int, device, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Rows ``(frame_id, xmin_norm, ymin_norm, xmax_norm, ymax_norm)`` of every stage-0 window (optionally into ``out``)."""
    if out is None:
        out = torch.empty((n_frames * plan.n_windows, 5), dtype=torch.int32, device=device)
    assert out.is_contiguous() and out.numel() == n_frames * plan.n_windows * 5
    with torch.cuda.device(out.device):
        check(_lib.lib().rodc_window_meta(plan.handle, n_frames, frame_id_base, _ptr(out), _stream()), "rodc_window_meta")
    return out


def window_gather_rows(plan: Plan, n_frames: int, pyramid_base_off: int, win_off: torch.Tensor, win_pitch: torch.Tensor):
    """Fill ``win_off`` (i64) / ``win_pitch`` (i32) slices for ``n_frames`` frames of ``plan`` whose packed pyramids
    start ``pyramid_base_off`` bytes into the shared gather buffer."""
    _require_cuda(win_off, win_pitch)
    check(_lib.lib().rodc_window_gather_rows(plan.handle, int(n_frames), int(pyramid_base_off), _ptr(win_off), _ptr(win_pitch),
                                             _stream()), "rodc_window_gather_rows")


def window_levels(plan: Plan, n_frames: int, slot_base: int, device, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Level index (+ ``slot_base``) of every stage-0 window of ``n_frames`` frames, ``[n_frames*N]`` i32."""
    if out is None:
        out = torch.empty((n_frames * plan.n_windows,), dtype=torch.int32, device=device)
    assert out.is_contiguous() and out.numel() == n_frames * plan.n_windows
    with torch.cuda.device(out.device):
        check(_lib.lib().rodc_window_levels(plan.handle, int(n_frames), int(slot_base), _ptr(out), _stream()), "rodc_window_levels")
    return out


def select_rows(src: torch.Tensor, idx: torch.Tensor, n: int) -> torch.Tensor:
    """``src[idx[:n]]`` for small per-window records (rows of a multiple of 4 bytes), done by librodc."""
    _require_cuda(src, idx)
    assert src.is_contiguous() and idx.dtype == torch.int32
    row_bytes = src.element_size() * (src[0].numel() if src.dim() > 1 else 1)
    out = torch.empty((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    check(_lib.lib().rodc_select_rows(_ptr(src), _ptr(idx), int(n), int(row_bytes), _ptr(out), _stream()), "rodc_select_rows")
    return out


def scaled_slots(entries: Sequence[Tuple[float, int, int, int, int, int, int]], device) -> torch.Tensor:
    """Device copy of a ``rodc_scaled_slot`` table from ``(scale, base_off, frame_stride, width, height,
    frame_base, individually)`` tuples."""
    arr = (_lib.ScaledSlot * max(len(entries), 1))()
    for i, e in enumerate(entries):
        arr[i] = _lib.ScaledSlot(float(e[0]), int(e[1]), int(e[2]), int(e[3]), int(e[4]), int(e[5]), int(e[6]))
    host = np.frombuffer(arr, dtype=np.uint8).copy()
    return torch.from_numpy(host).to(device)


def window_scaled_rows(meta: torch.Tensor, slot: torch.Tensor, n: int, slots: torch.Tensor, n_slots: int, size: int,
                       dense_rel_off: int = 0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """GATHER rows of the resize-then-crop branch; returns ``(win_off, win_pitch, err)`` (``err`` is a device
    i32: bit 0 = a crop leaves its resized frame, the reference's ``Window.raw`` ValueError)."""
    _require_cuda(meta, slot, slots)
    win_off = torch.empty((n,), dtype=torch.int64, device=meta.device)
    win_pitch = torch.empty((n,), dtype=torch.int32, device=meta.device)
    err = torch.zeros((1,), dtype=torch.int32, device=meta.device)
    check(_lib.lib().rodc_window_scaled_rows(_ptr(meta), _ptr(slot), int(n), _ptr(slots), int(n_slots), int(size),
                                             int(dense_rel_off), _ptr(win_off), _ptr(win_pitch), _ptr(err), _stream()),
          "rodc_window_scaled_rows")
    return win_off, win_pitch, err


def upload_frames(frames: Sequence[np.ndarray], flat: torch.Tensor, offs: Sequence[int]) -> None:
    """Asynchronous H2D copies of page-locked host frames into ``flat`` at byte offsets ``offs`` (one library call)."""
    _require_cuda(flat)
    n = len(frames)
    ptrs = (C.c_void_p * n)(*[f.__array_interface__["data"][0] for f in frames])
    sizes = np.asarray([f.nbytes for f in frames], np.int64)
    o = np.asarray(offs, np.int64)
    check(_lib.lib().rodc_upload_frames(ptrs, sizes.ctypes.data_as(_lib.c_i64p), o.ctypes.data_as(_lib.c_i64p), n,
                                        _ptr(flat), _stream()), "rodc_upload_frames")


class FrameTable:
    """Device table of the original frames for the RECROP source: one flat u8 buffer plus
    per-frame byte offsets and (H, W)."""

    def __init__(self, base: torch.Tensor, offsets: Sequence[int], hw: Sequence[Tuple[int, int]]):
        _require_cuda(base)
        self.base = base
        self.off = torch.tensor(list(offsets), dtype=torch.int64, device=base.device)
        self.hw = torch.tensor([list(x) for x in hw], dtype=torch.int32, device=base.device).reshape(-1, 2)
        self.n = len(offsets)

    @staticmethod
    def from_batch(frames: torch.Tensor) -> "FrameTable":
        n, h, w, _ = frames.shape
        return FrameTable(frames, [i * h * w * 3 for i in range(n)], [(h, w)] * n)

    @staticmethod
    def from_list(frames: List[torch.Tensor]) -> "FrameTable":
        flat = torch.cat([f.reshape(-1) for f in frames])
        offs, o = [], 0
        for f in frames:
            offs.append(o); o += f.numel()
        return FrameTable(flat, offs, [(f.shape[0], f.shape[1]) for f in frames])


def recrop(table: FrameTable, meta: torch.Tensor, size: int) -> torch.Tensor:
    """K4 standalone: ``cv2.resize(frame[y0:y1, x0:x1], (size,size))`` per meta row -> ``[n,size,size,3]`` u8."""
    _require_cuda(meta)
    n = meta.shape[0]
    out = torch.empty((n, size, size, 3), dtype=torch.uint8, device=meta.device)
    check(_lib.lib().rodc_recrop(_ptr(table.base), _ptr(table.off), _ptr(table.hw), _ptr(meta), n, size, _ptr(out),
                                 _stream()), "rodc_recrop")
    return out


# --------------------------------------------------------------------------------- stage
class Stage:
    """One stage net on the device (weights in the reference's layouts, see network/stage_model.py).
    ``geometry`` = dict(in_size, conv_stride, pool_size, pool_stride) for non-default architectures."""

    def __init__(self, weights: Dict[str, np.ndarray], device=None, geometry: Optional[Dict[str, int]] = None):
        if not torch.cuda.is_available():
            raise _lib.RodcError("a CUDA device is required (no CPU fallback)")
        geo = dict(conv_stride=1, pool_size=3, pool_stride=1)
        geo.update(geometry or {})
        n_conv = 0
        while "conv%d_W" % n_conv in weights:
            n_conv += 1
        cws = [np.ascontiguousarray(weights["conv%d_W" % j], np.float32) for j in range(n_conv)]
        cbs = [np.ascontiguousarray(weights["conv%d_b" % j], np.float32) for j in range(n_conv)]
        ksize = int(cws[0].shape[0])
        chans = np.asarray([c.shape[3] for c in cws], np.int32)
        self.channels = int(chans[-1])
        f1 = np.ascontiguousarray(weights["fc1_W"], np.float32)
        self.K, self.fc1_size = int(f1.shape[0]), int(f1.shape[1])
        if "in_size" in geo:
            self.in_size = int(geo["in_size"])
        else:
            self.in_size = int(round((self.K // self.channels) ** 0.5))
            assert self.in_size * self.in_size * self.channels == self.K, "pass geometry['in_size'] for strided nets"
        f2 = np.ascontiguousarray(weights["fc2_W"], np.float32)
        self.n_prev = int(f2.shape[0]) - self.fc1_size
        self.width = self.fc1_size + self.n_prev
        f1b = np.ascontiguousarray(weights["fc1_b"], np.float32)
        f2b = np.ascontiguousarray(weights["fc2_b"], np.float32)
        h = C.c_void_p()
        wptr = (C.c_void_p * n_conv)(*[c.ctypes.data for c in cws])
        bptr = (C.c_void_p * n_conv)(*[c.ctypes.data for c in cbs])
        ctx = torch.cuda.device(device) if device is not None else torch.cuda.device(torch.cuda.current_device())
        with ctx:
            torch.cuda.init()
            check(_lib.lib().rodc_stage_create_general(
                self.in_size, n_conv, chans.ctypes.data_as(_lib.c_i32p), ksize, int(geo["conv_stride"]),
                int(geo["pool_size"]), int(geo["pool_stride"]), wptr, bptr, self.fc1_size, self.n_prev,
                C.c_void_p(f1.ctypes.data), C.c_void_p(f1b.ctypes.data), C.c_void_p(f2.ctypes.data),
                C.c_void_p(f2b.ctypes.data), C.byref(h)), "rodc_stage_create_general")
        self._h = h
        # the reference default (one 3x3x32 conv, 3x3/1 pool) runs the specialised kernels incl. the tensor fc1
        self.is_default = (n_conv == 1 and int(chans[0]) == 32 and ksize == 3 and int(geo["conv_stride"]) == 1 and
                           int(geo["pool_size"]) == 3 and int(geo["pool_stride"]) == 1 and self.in_size % 4 == 0 and
                           self.fc1_size == 512)

    @property
    def handle(self):
        return self._h

    def workspace_bytes(self, n: int, impl: int = IMPL_TENSOR) -> int:
        return int(_lib.lib().rodc_stage_workspace_bytes(self._h, int(n), impl))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().rodc_stage_destroy(self._h)
                self._h = None
        except Exception:
            pass


def make_source(*, dense: Optional[torch.Tensor] = None, plan: Optional[Plan] = None,
                pyr: Optional[torch.Tensor] = None, table: Optional[FrameTable] = None,
                meta: Optional[torch.Tensor] = None, gather_base: Optional[torch.Tensor] = None,
                win_off: Optional[torch.Tensor] = None, win_pitch: Optional[torch.Tensor] = None) -> WindowSource:
    s = WindowSource()
    if gather_base is not None:
        _require_cuda(gather_base, win_off, win_pitch)
        assert win_off.dtype == torch.int64 and win_pitch.dtype == torch.int32
        s.kind = SRC_GATHER; s.gather_base = gather_base.data_ptr(); s.win_off = win_off.data_ptr()
        s.win_pitch = win_pitch.data_ptr()
    elif dense is not None:
        _require_cuda(dense)
        assert dense.dtype == torch.uint8 and dense.is_contiguous()
        s.kind = SRC_DENSE; s.dense = dense.data_ptr()
    elif plan is not None:
        _require_cuda(pyr)
        s.kind = SRC_PYRAMID; s.plan = plan.handle.value; s.pyramid = pyr.data_ptr()
    else:
        _require_cuda(meta)
        assert meta.dtype == torch.int32 and meta.is_contiguous()
        s.kind = SRC_RECROP; s.frames_base = table.base.data_ptr(); s.frame_off = table.off.data_ptr()
        s.frame_hw = table.hw.data_ptr(); s.meta = meta.data_ptr()
    return s


def stage_forward(stage: Stage, src: WindowSource, first: int, n: int, bottleneck: torch.Tensor,
                  workspace: torch.Tensor, impl: int = IMPL_TENSOR,
                  logits: Optional[torch.Tensor] = None, probs: Optional[torch.Tensor] = None,
                  n_dev: Optional[torch.Tensor] = None):
    """K2.  Writes ``bottleneck[:, :fc1]``; returns ``(logits [n,2], probs [n,2])``.  ``n_dev`` (device i64 [1]):
    ``n`` is an upper bound and the kernels process ``min(n, n_dev)`` windows (``rodc_stage_forward_dyn``)."""
    _require_cuda(bottleneck, workspace)
    dev = bottleneck.device
    assert bottleneck.dtype == torch.float32 and bottleneck.stride(1) == 1 and bottleneck.shape[0] >= n
    if logits is None:
        logits = torch.empty((n, 2), dtype=torch.float32, device=dev)
    if probs is None:
        probs = torch.empty((n, 2), dtype=torch.float32, device=dev)
    check(_lib.lib().rodc_stage_forward_dyn(stage.handle, C.byref(src), int(first), int(n), _ptr(n_dev), _ptr(bottleneck),
                                            int(bottleneck.stride(0)), _ptr(logits), _ptr(probs), _ptr(workspace),
                                            int(workspace.numel() * workspace.element_size()), int(impl), _stream()),
          "rodc_stage_forward")
    return logits, probs


def source_windows(src: WindowSource, first: int, n: int, size: int, device) -> torch.Tensor:
    """The u8 stage input ``[n,size,size,3]`` of any window source (the reference's ``windows_raw``)."""
    out = torch.empty((n, size, size, 3), dtype=torch.uint8, device=device)
    check(_lib.lib().rodc_source_windows(C.byref(src), int(first), int(n), int(size), _ptr(out), _stream()), "rodc_source_windows")
    return out


def stage_features(stage: Stage, src: WindowSource, first: int, n: int, device) -> torch.Tensor:
    out = torch.empty((n, stage.K), dtype=torch.float32, device=device)
    check(_lib.lib().rodc_stage_features_f32(stage.handle, C.byref(src), int(first), int(n), _ptr(out), _stream()),
          "rodc_stage_features_f32")
    return out


# --------------------------------------------------------------------------------- compaction
CONF_LAST, CONF_AVG, CONF_MULT = 0, 1, 2


def compact(probs: torch.Tensor, n: int, threshold: float, *, meta_in=None, meta_out=None,
            bott_in=None, width: int = 0, bott_out=None, col_off: int = 0,
            conf_acc=None, conf_out=None, conf_mode: int = CONF_LAST,
            keep_idx: Optional[torch.Tensor] = None, count: Optional[torch.Tensor] = None,
            n_dev: Optional[torch.Tensor] = None):
    """K3.  Returns ``(keep_idx i32 [n], count i64 [1])`` -- both on the device (no sync).  ``n_dev``: device-side
    count of the input rows (``n`` is then an upper bound)."""
    _require_cuda(probs)
    dev = probs.device
    if keep_idx is None:
        keep_idx = torch.empty((max(n, 1),), dtype=torch.int32, device=dev)
    if count is None:
        count = torch.zeros((1,), dtype=torch.int64, device=dev)
    L = _lib.lib()
    scratch = torch.empty((int(L.rodc_compact_scratch_bytes(int(n))),), dtype=torch.uint8, device=dev)
    check(L.rodc_compact_dyn(_ptr(probs), int(n), _ptr(n_dev), float(threshold), _ptr(meta_in), _ptr(meta_out),
                         _ptr(bott_in), int(bott_in.stride(0)) if bott_in is not None else 0, int(width),
                         _ptr(bott_out), int(bott_out.stride(0)) if bott_out is not None else 0, int(col_off),
                         _ptr(conf_acc), _ptr(conf_out), int(conf_mode), _ptr(keep_idx), _ptr(count), _ptr(scratch),
                         _stream()), "rodc_compact")
    return keep_idx, count


# --------------------------------------------------------------------------------- group rectangles
def group_rectangles(cand: torch.Tensor, k: int, n_frames: int, min_neighbors: int = 1, eps: float = 0.2,
                     k_dev: Optional[torch.Tensor] = None, count: Optional[torch.Tensor] = None):
    """K5.  cand ``[K,5]`` i32 (frame, xmin,ymin,xmax,ymax) grouped by frame ->
    ``(out [K,6] i32 (frame,x,y,w,h,weight), count i64 [1])`` on the device."""
    _require_cuda(cand)
    dev = cand.device
    out = torch.empty((max(k, 1), 6), dtype=torch.int32, device=dev)
    if count is None:
        count = torch.zeros((1,), dtype=torch.int64, device=dev)
    L = _lib.lib()
    scratch = torch.empty((int(L.rodc_group_scratch_bytes(int(k), int(n_frames))),), dtype=torch.uint8, device=dev)
    check(L.rodc_group_rectangles_dyn(_ptr(cand), int(k), _ptr(k_dev), int(n_frames), int(min_neighbors), float(eps), _ptr(out),
                                      _ptr(count), _ptr(scratch), _stream()), "rodc_group_rectangles")
    return out, count
