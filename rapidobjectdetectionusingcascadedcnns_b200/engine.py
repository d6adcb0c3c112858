"""Device-side cascade driver: one CUDA stream, librodc kernels only.

pyramid (K1) -> [stage i: fused window producer + conv/pool (K2a/K4) -> tensor fc1 (K2b) ->
fc2/softmax (K2c) -> threshold + stable compaction (K3)] x n -> groupRectangles (K5).

Restates the control flow of `app/inference_cascade_app.py:119-342` (cascade) and
`app/inference_app.py:117-154` (single net) for a batch of frames; frames are independent
(`app/inference_app.py:187-201` groups NMS per image), so batching is only a throughput device.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import config as cf
from . import native


@dataclass
class GroupResult:
    """Detections of one batch of same-size frames."""
    boxes: np.ndarray                # [K,4] xmin,ymin,xmax,ymax (int32)
    conf: np.ndarray                 # [K]   float (softmax-based before NMS, cluster size after)
    frame: np.ndarray                # [K]   frame index inside the batch, ascending
    n_windows: int = 0
    stage_in: List[int] = field(default_factory=list)     # windows entering each stage
    stage_kept: List[int] = field(default_factory=list)
    trace: Optional[dict] = None     # per-stage device tensors when requested (tests)


@dataclass
class WindowOrigin:
    """Where the stage-0 windows of a run live, for the two next-stage input rules that do not re-crop
    from the original frame (`app/inference_cascade_app.py:234-240,374-394`).

    ``gather``: closure -> ``(base, win_off [n0] i64, win_pitch [n0] i32)`` addressing every stage-0 window
    (cascades whose stages share one input size keep the SAME raw windows).
    ``slot`` / ``groups``: per stage-0 window the index of its (frame group, pyramid level) and, per group,
    ``dict(frames=[n,H,W,3] u8 CUDA, frame_base, scales=float64 [L], width)`` -- the resize-then-crop
    branch (``cascade_scale_patches_individually=False``) resizes whole frames to ``scale * size / width``."""
    gather: Optional[object] = None
    slot: Optional[torch.Tensor] = None
    groups: Optional[List[dict]] = None


@dataclass
class MixedFrames:
    """Device-resident frames of several shapes (``CascadeEngine.upload_mixed``): ``flat`` holds the frames
    grouped by shape, ``offs`` / ``shapes`` the byte offset and (H, W) of every frame in internal order,
    ``groups`` the frame count per shape (insertion order), ``order`` internal id -> index of the caller's list."""
    flat: torch.Tensor
    offs: List[int]
    shapes: List[Tuple[int, int]]
    groups: Dict[Tuple[int, int], int]
    order: List[int]


class CascadeEngine:
    """Runs a list of ``native.Stage`` over frame batches.  ``single=True`` applies the
    single-net decision rule (argmax != background, confidence = p[argmax])."""

    def __init__(self, stage_weights: Sequence[Dict[str, np.ndarray]], device: Optional[int] = None,
                 single: bool = False, geometries: Optional[Sequence[Optional[Dict[str, int]]]] = None):
        if not torch.cuda.is_available():
            raise native._lib.RodcError("CascadeEngine needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        with torch.cuda.device(self.device):
            geos = list(geometries) if geometries is not None else [None] * len(stage_weights)
            self.stages = [native.Stage(w, self.device, g) for w, g in zip(stage_weights, geos)]
        self.single = single
        self.sizes = [s.in_size for s in self.stages]
        self._plans: Dict[Tuple, native.Plan] = {}
        self._scaled_plans: Dict[Tuple, native.Plan] = {}
        self._ws: Optional[torch.Tensor] = None
        import threading
        self._tls = threading.local()
        self._validate_chain()

    # ------------------------------------------------------------------ helpers
    def _validate_chain(self):
        """The bottleneck hand-off trusts nothing: stage i + 1 concatenates ``[own fc1 | previous bottleneck]``
        (`network/net.py:140-145`), so its fc2 must have exactly ``width(stage i)`` extra rows and stage 0 none
        (TensorFlow would refuse to feed a placeholder of another shape, `app/inference_cascade_app.py:203-215`)."""
        st = self.stages
        if st[0].n_prev != 0:
            raise ValueError("stage 0 expects a previous bottleneck of %d values, but a cascade starts without one" % st[0].n_prev)
        for i in range(len(st) - 1):
            if st[i + 1].n_prev not in (0, st[i].width):
                raise ValueError("stage %d expects a previous bottleneck of %d values, stage %d provides %d"
                                 % (i + 1, st[i + 1].n_prev, i, st[i].width))

    @property
    def kernel_launches(self) -> int:
        """Kernel launches issued by librodc in this process so far (measured inside the library)."""
        return int(native._lib.lib().rodc_launch_count())

    def plan_for(self, w: int, h: int) -> native.Plan:
        key = (w, h, self.sizes[0], float(cf.get("window_scale_factor")), float(cf.get("min_window_length")))
        p = self._plans.get(key)
        if p is None:
            p = native.Plan(*key)
            self._plans[key] = p
        return p

    def scaled_plan_for(self, w: int, h: int, scales: Tuple[float, ...]) -> native.Plan:
        key = (w, h, scales)
        p = self._scaled_plans.get(key)
        if p is None:
            if len(self._scaled_plans) > 64:
                self._scaled_plans.clear()
            p = native.Plan.from_scales(w, h, scales)
            self._scaled_plans[key] = p
        return p

    @property
    def increasing_dims(self) -> bool:
        """`app/inference_cascade_app.py:403-420`: first stage input narrower than the last one."""
        return self.sizes[0] < self.sizes[-1]

    @staticmethod
    def _scaled_branch_possible() -> bool:
        """True when some survivor may take the resize-then-crop branch (`:362-365`)."""
        return (not cf.get("cascade_scale_patches_individually")) or \
            bool(cf.get("cascade_scale_patches_individually_iff_not_cached"))

    def _scaled_source(self, origin: WindowOrigin, size: int, meta: torch.Tensor, slot: torch.Tensor, n: int,
                       table: native.FrameTable):
        """Next-stage input by resize-then-crop (`app/inference_cascade_app.py:374-394`): build the frames
        ``raw_scaled(next_scale)`` of every (group, level) that takes the branch (K1 on an explicit scale list)
        and address the survivors' ``size x size`` crops as GATHER rows.  Returns ``None`` when no level takes
        the branch (``..._iff_not_cached`` with nothing cached: every survivor is cropped individually)."""
        individually = bool(cf.get("cascade_scale_patches_individually"))
        per_group, total, any_scaled, any_individual = [], 0, False, False
        for g in origin.groups:
            # next_scale = window.scale * (cf img_width / window.width)                       (`:358-359`)
            ns = np.asarray(g["scales"], np.float64) * (size / np.asarray(g["width"], np.float64))
            # is_raw_scaled_cached(next_scale): the per-image cache holds the levels the stage-0 extraction
            # touched, keyed by the scale's repr; scale 1.0 is never stored under such a key
            # (`data/image_info.py:201-213,252-258`)
            cached = {float(v) for v in np.asarray(g["scales"], np.float64).tolist() if float(v) != 1.0}
            use = [(not individually) or (float(v) in cached) for v in ns.tolist()]
            sel = [k for k, u in enumerate(use) if u]
            any_scaled |= bool(sel); any_individual |= len(sel) < len(use)
            fr = g["frames"]
            plan2 = self.scaled_plan_for(int(fr.shape[2]), int(fr.shape[1]), tuple(float(ns[k]) for k in sel)) if sel else None
            per_group.append((g, ns, sel, plan2, total))
            if plan2 is not None:
                total += int(fr.shape[0]) * plan2.pyramid_bytes
        if not any_scaled:
            return None
        buf = torch.empty((max(total, 16),), dtype=torch.uint8, device=self.device)
        entries = []
        for g, ns, sel, plan2, base in per_group:
            fr = g["frames"]; nf = int(fr.shape[0])
            pos = {k: j for j, k in enumerate(sel)}
            if plan2 is not None:
                native.pyramid(plan2, fr, out=buf[base:base + nf * plan2.pyramid_bytes].view(nf, plan2.pyramid_bytes))
            for k in range(len(ns)):
                if k in pos:
                    j = pos[k]
                    entries.append((float(ns[k]), base + int(plan2.byte_offset[j]), plan2.pyramid_bytes if nf > 1 else 0,
                                    int(plan2.level_w[j]), int(plan2.level_h[j]), int(g["frame_base"]), 0))
                else:
                    entries.append((float(ns[k]), 0, 0, 0, 0, int(g["frame_base"]), 1))
        slots = native.scaled_slots(entries, self.device)
        dense, rel = None, 0
        if any_individual:                    # mixed: those survivors keep the crop + cv2.resize input (K4, materialised)
            dense = native.recrop(table, meta[:n], size)
            rel = dense.data_ptr() - buf.data_ptr()
        win_off, win_pitch, err = native.window_scaled_rows(meta, slot, n, slots, len(entries), size, rel)
        flags = int(err.item())
        if flags & 1:
            raise ValueError("ymax/xmax must not leave the image boundaries")      # `data/rectangles.py:286-289`
        if flags:
            raise native._lib.RodcError("rodc_window_scaled_rows: bad level slot")
        self._keepalive2 = (buf, slots, dense, win_off, win_pitch)
        return native.make_source(gather_base=buf, win_off=win_off, win_pitch=win_pitch)

    @staticmethod
    def _all_pinned(frames: Sequence[np.ndarray]) -> bool:
        """True when every frame is a contiguous array in page-locked memory (registry of ``ArrayImageInfo(pin=True)``
        first, the driver's answer otherwise)."""
        from .data.image_info import PINNED_FRAME_PTRS
        for f in frames:
            if not f.flags["C_CONTIGUOUS"]:
                return False
            if f.__array_interface__["data"][0] in PINNED_FRAME_PTRS:
                continue
            if not f.flags["WRITEABLE"] or not torch.from_numpy(f).is_pinned():
                return False
        return True

    def _pinned_stage(self, nbytes: int) -> torch.Tensor:
        """The pinned staging buffer (grown on demand).  Before it is handed out again the H2D copies issued from
        it by the previous user must have finished: they are asynchronous and the host is about to overwrite it."""
        slots = self.__dict__.setdefault("_stage_slots", {})
        slot = slots.setdefault(self.staging_slot, {"buf": None, "busy": None})
        if slot["busy"] is not None:
            slot["busy"].synchronize()
            slot["busy"] = None
        if slot["buf"] is None or slot["buf"].numel() < nbytes:
            slot["buf"] = torch.empty((nbytes,), dtype=torch.uint8).pin_memory()
        return slot["buf"]

    def _pinned_issued(self):
        """Call after the last asynchronous copy out of the staging buffer has been enqueued."""
        ev = torch.cuda.Event()
        ev.record()
        self._stage_slots[self.staging_slot]["busy"] = ev

    @property
    def staging_slot(self) -> int:
        """Which pinned staging buffer the calling thread uses (thread local; the upload thread of the app's
        decode -> upload pipeline alternates between two of its own so that the upload of chunk k + 1 can be
        staged while chunk k's copy is still in flight)."""
        return getattr(self._tls, "slot", 0)

    @staging_slot.setter
    def staging_slot(self, v: int):
        self._tls.slot = int(v)

    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty((nbytes,), dtype=torch.uint8, device=self.device)
        return self._ws

    def _threshold(self, i: int) -> float:
        t = cf.get("foreground_confidence_threshold")
        try:
            return float(t)
        except TypeError:
            pass
        if len(t) == len(self.stages):
            return float(t[i])
        raise ValueError("Invalid foreground_confidence_threshold.")

    @staticmethod
    def _impl() -> int:
        return native.IMPL_SIMT if cf.get("fc1_impl") == cf.FC1_IMPL_SIMT else native.IMPL_TENSOR

    # ------------------------------------------------------------------ stages on a window set
    def _run_stages(self, src0, n0: int, meta0: torch.Tensor, table: native.FrameTable, n_frames: int,
                    keep_trace: bool, origin: Optional[WindowOrigin] = None, while_busy=None) -> GroupResult:
        """``while_busy``: called ONCE, right after stage 0 has been enqueued and before its survivor count is read --
        host work done there (preparing and uploading the next chunk) is hidden behind the GPU's stage-0 time."""
        dev = self.device
        impl = self._impl()
        n_st = len(self.stages)
        mode_key = cf.get("final_confidence_calculation")
        conf_mode = {cf.FINAL_CONFIDENCE_CALCULATION_LAST_STEP: native.CONF_LAST,
                     cf.FINAL_CONFIDENCE_CALCULATION_AVG: native.CONF_AVG,
                     cf.FINAL_CONFIDENCE_CALCULATION_MULT: native.CONF_MULT}[mode_key]
        res = GroupResult(np.zeros((0, 4), np.int32), np.zeros((0,), np.float32), np.zeros((0,), np.int32), n_windows=n0)
        trace = {"probs": [], "logits": [], "keep": [], "meta": [], "X": []} if keep_trace else None
        n, meta, src = n0, meta0, src0
        bott = torch.empty((n, self.stages[0].width), dtype=torch.float32, device=dev)
        conf_acc = None
        same_windows = n_st > 1 and not self.increasing_dims          # `:234-240`: the raw windows are kept
        scaled = n_st > 1 and self.increasing_dims and self._scaled_branch_possible()
        if (same_windows or scaled) and origin is None:
            raise ValueError("this cascade configuration needs the origin of the stage-0 windows")
        rows = origin.gather() if same_windows else None            # (base, win_off, win_pitch) of the live windows
        slot = origin.slot if scaled else None
        for i, st in enumerate(self.stages):
            if n < 1:
                break                                                       # inference_cascade_app.py:140-142
            last = i == n_st - 1
            ws_bytes = st.workspace_bytes(n, impl)
            ws = self._workspace(ws_bytes)
            if keep_trace:                      # the stage's u8 input, i.e. the reference's windows_raw (tests)
                trace["X"].append(native.source_windows(src, 0, n, st.in_size, dev))
            logits, probs = native.stage_forward(st, src, 0, n, bott, ws[:ws_bytes], impl)
            # single net: keep argmax != background <=> p[1] > p[0] (ties -> background, np.argmax picks index 0),
            # confidence = p[argmax] = p[1] (inference_app.py:140-149); NaN selects that rule in rodc_compact
            thr = float("nan") if self.single else self._threshold(i)
            meta_out = torch.empty((n, 5), dtype=torch.int32, device=dev)
            conf_out = torch.empty((n,), dtype=torch.float32, device=dev)
            use_b = (not last) and bool(cf.get("reuse_bottlenecks"))
            nxt = self.stages[i + 1] if not last else None
            if nxt is not None and nxt.n_prev > 0 and not use_b:
                # the reference would leave the bottleneck_in placeholder unfed and TensorFlow raises
                # (`app/inference_cascade_app.py:203-215`); never run fc2 over uninitialised columns
                raise ValueError("stage %d was built with a %d-wide bottleneck input but reuse_bottlenecks is off"
                                 % (i + 1, nxt.n_prev))
            bott_next = torch.empty((n, nxt.width), dtype=torch.float32, device=dev) if nxt is not None else None
            keep, count = native.compact(
                probs, n, thr, meta_in=meta, meta_out=meta_out,
                bott_in=bott if use_b and nxt.n_prev > 0 else None, width=st.width,
                bott_out=bott_next if use_b and nxt.n_prev > 0 else None, col_off=nxt.fc1_size if nxt is not None else 0,
                conf_acc=conf_acc, conf_out=conf_out, conf_mode=conf_mode if conf_acc is not None else native.CONF_LAST)
            if while_busy is not None:
                hook, while_busy = while_busy, None
                hook()
            kept = int(count.item())                                        # one host sync per stage
            res.stage_in.append(n); res.stage_kept.append(kept)
            if keep_trace:
                trace["probs"].append(probs); trace["logits"].append(logits); trace["keep"].append(keep[:kept].clone())
                trace["meta"].append(meta)
            meta, n = meta_out, kept
            if not last and kept > 0:
                bott = bott_next
                src = None
                if same_windows:
                    # the reference appends windows_raw[window_index_original] only next to the bottleneck copy
                    # (`:234-237`); without reuse_bottlenecks it would feed an empty array -- the windows are kept here too
                    rows = (rows[0], native.select_rows(rows[1], keep, kept), native.select_rows(rows[2], keep, kept))
                    src = native.make_source(gather_base=rows[0], win_off=rows[1], win_pitch=rows[2])
                elif scaled:
                    slot = native.select_rows(slot, keep, kept)
                    src = self._scaled_source(origin, nxt.in_size, meta, slot, kept, table)
                if src is None:
                    src = native.make_source(table=table, meta=meta)
                self._keepalive = (table, meta, rows, slot)
                if mode_key != cf.FINAL_CONFIDENCE_CALCULATION_LAST_STEP:
                    conf_acc = conf_out
            elif last:
                res = self._finish(res, meta, conf_out[:kept], kept, n_frames)
        res.trace = trace
        return res

    def _finish(self, res: GroupResult, meta: torch.Tensor, conf: torch.Tensor, k: int, n_frames: int) -> GroupResult:
        """``_postprocess_bboxes`` up to the vertical enlargement (done on the host by the app)."""
        if k == 0:
            return res
        if cf.get("nms") == cf.NMS_OPENCV:
            out, count = native.group_rectangles(meta, k, n_frames, int(cf.get("nms_opencv_min_neighbors")), 0.2)
            kk = int(count.item())
            rows = out[:kk].cpu().numpy()
            res.frame = rows[:, 0].copy()
            res.boxes = np.column_stack([rows[:, 1], rows[:, 2], rows[:, 1] + rows[:, 3], rows[:, 2] + rows[:, 4]]).astype(np.int32)
            res.conf = rows[:, 5].astype(np.float64)
        else:
            rows = meta[:k].cpu().numpy()
            res.frame = rows[:, 0].copy()
            res.boxes = rows[:, 1:5].astype(np.int32)
            res.conf = self._finalise_confidences(conf.cpu().numpy())
        return res

    def _finalise_confidences(self, c: np.ndarray) -> np.ndarray:
        """`app/inference_cascade_app.py:322-327` in the reference's own arithmetic: a float32 NumPy array divided in
        place by the number of nets (a true IEEE division -- a device-side ``tensor / scalar`` multiplies by the
        reciprocal and can differ in the last bit) or floored at the MULT minimum."""
        mode_key = cf.get("final_confidence_calculation")
        c = np.array(c, np.float32)
        if mode_key == cf.FINAL_CONFIDENCE_CALCULATION_AVG:
            c /= len(self.stages)
        elif mode_key == cf.FINAL_CONFIDENCE_CALCULATION_MULT:
            c[c < cf.MIN_SCORE_FOR_FINAL_CONFIDENCE_CALCULATION_MULT] = cf.MIN_SCORE_FOR_FINAL_CONFIDENCE_CALCULATION_MULT
        return c

    # ------------------------------------------------------------------ public entry points
    def stage_host_frames(self, frames: Sequence[np.ndarray]) -> torch.Tensor:
        """Gather same-size host frames into ONE pinned ``[n,H,W,3]`` u8 buffer (reused across
        calls) so the H2D transfer is a single asynchronous copy."""
        n = len(frames)
        h, w, _ = frames[0].shape
        need = n * h * w * 3
        buf = self._pinned_stage(need)
        view = buf[:need].view(n, h, w, 3)
        dst = view.numpy()
        if n >= 4:                                   # memcpy releases the GIL: gather with a few threads
            pool = getattr(self, "_copy_pool", None)
            if pool is None:
                from concurrent.futures import ThreadPoolExecutor
                pool = self._copy_pool = ThreadPoolExecutor(max_workers=4)

            def _cp(i):
                dst[i] = frames[i]
            list(pool.map(_cp, range(n)))
        else:
            for i, f in enumerate(frames):
                dst[i] = f
        return view

    def upload_host_frames(self, frames: Sequence[np.ndarray]) -> torch.Tensor:
        """Same-size host frames -> ``[n,H,W,3]`` u8 on the device.  The frames are gathered into the pinned
        buffer group by group and every group's H2D copy is issued as soon as it is staged, so the DMA of group g
        runs while the host gathers group g + 1 (the reference re-reads the files per scale,
        `data/image_info.py:228-240`; here a frame crosses PCIe once)."""
        n = len(frames)
        h, w, _ = frames[0].shape
        need = n * h * w * 3
        # frames that already live in page-locked memory (``ArrayImageInfo(..., pin=True)``, capture buffers) are
        # copied straight from where they are: n asynchronous H2D copies, no host-side gather
        tens = frames
        if self._all_pinned(frames):
            with torch.cuda.device(self.device):
                dev = torch.empty((n, h, w, 3), dtype=torch.uint8, device=self.device)
                native.upload_frames(frames, dev, [i * h * w * 3 for i in range(n)])
            self._keepalive_host = tens
            return dev
        buf = self._pinned_stage(need)
        host = buf[:need].view(n, h, w, 3)
        dst = host.numpy()
        with torch.cuda.device(self.device):
            dev = torch.empty((n, h, w, 3), dtype=torch.uint8, device=self.device)
            step = max(1, (n + 3) // 4)
            pool = getattr(self, "_copy_pool", None)
            if pool is None and n >= 4:
                from concurrent.futures import ThreadPoolExecutor
                pool = self._copy_pool = ThreadPoolExecutor(max_workers=4)

            def _cp(i):
                dst[i] = frames[i]
            for g0 in range(0, n, step):
                g1 = min(n, g0 + step)
                if pool is not None and g1 - g0 >= 2:
                    list(pool.map(_cp, range(g0, g1)))
                else:
                    for i in range(g0, g1):
                        dst[i] = frames[i]
                dev[g0:g1].copy_(host[g0:g1], non_blocking=True)
            self._pinned_issued()
        return dev

    # ------------------------------------------------------------------ latency path: no host round trips, CUDA graph
    def _latency_eligible(self, n_windows: int) -> bool:
        """The whole cascade can be enqueued with device-side survivor counts when every stage runs the specialised
        kernels with the tensor fc1 and the next-stage input is the default crop + ``cv2.resize``."""
        mode = cf.get("latency_mode")
        if mode == cf.LATENCY_OFF or self._impl() != native.IMPL_TENSOR:
            return False
        if not all(st.is_default for st in self.stages):
            return False
        n_st = len(self.stages)
        if n_st > 1 and (not self.increasing_dims or self._scaled_branch_possible()):
            return False
        if n_st > 1 and not cf.get("reuse_bottlenecks") and any(st.n_prev > 0 for st in self.stages[1:]):
            return False
        return mode == cf.LATENCY_ALWAYS or n_windows <= int(cf.get("latency_max_windows"))

    def _enqueue_nosync(self, frames: torch.Tensor, plan: native.Plan, ws: torch.Tensor, table: native.FrameTable):
        """Enqueue pyramid -> all stages -> compaction -> NMS for resident same-size frames WITHOUT reading a single
        survivor count on the host: every stage is launched for the upper bound n0 and takes its real window count
        from the previous stage's device-side counter (``rodc_*_dyn``).  Returns the device tensors that hold the
        result: ``counts`` i64 [n_stages + 1] (survivors per stage, then the final row count), ``rows`` i32 [n0, 6]
        (NMS) or ``meta`` i32 [n0, 5] + ``conf`` f32 [n0] (NMS disabled)."""
        dev = self.device
        n_st = len(self.stages)
        n = int(frames.shape[0])
        n0 = n * plan.n_windows
        mode_key = cf.get("final_confidence_calculation")
        conf_mode = {cf.FINAL_CONFIDENCE_CALCULATION_LAST_STEP: native.CONF_LAST,
                     cf.FINAL_CONFIDENCE_CALCULATION_AVG: native.CONF_AVG,
                     cf.FINAL_CONFIDENCE_CALCULATION_MULT: native.CONF_MULT}[mode_key]
        counts = torch.zeros((n_st + 1,), dtype=torch.int64, device=dev)
        pyr = native.pyramid(plan, frames)
        meta = native.window_meta(plan, n, 0, dev)
        src = native.make_source(plan=plan, pyr=pyr)
        bott = torch.empty((n0, self.stages[0].width), dtype=torch.float32, device=dev)
        keep_alive = [pyr, table, meta]
        n_dev, conf_acc, conf_out = None, None, None
        for i, st in enumerate(self.stages):
            last = i == n_st - 1
            _, probs = native.stage_forward(st, src, 0, n0, bott, ws, native.IMPL_TENSOR, n_dev=n_dev)
            thr = float("nan") if self.single else self._threshold(i)
            meta_out = torch.empty((n0, 5), dtype=torch.int32, device=dev)
            conf_out = torch.empty((n0,), dtype=torch.float32, device=dev)
            nxt = self.stages[i + 1] if not last else None
            use_b = nxt is not None and bool(cf.get("reuse_bottlenecks")) and nxt.n_prev > 0
            bott_next = torch.empty((n0, nxt.width), dtype=torch.float32, device=dev) if nxt is not None else None
            native.compact(probs, n0, thr, meta_in=meta, meta_out=meta_out,
                           bott_in=bott if use_b else None, width=st.width, bott_out=bott_next if use_b else None,
                           col_off=nxt.fc1_size if nxt is not None else 0,
                           conf_acc=conf_acc, conf_out=conf_out, conf_mode=conf_mode if conf_acc is not None else native.CONF_LAST,
                           count=counts[i:i + 1], n_dev=n_dev)
            keep_alive += [probs, meta, bott]
            meta, n_dev = meta_out, counts[i:i + 1]
            if not last:
                bott = bott_next
                src = native.make_source(table=table, meta=meta)
                if mode_key != cf.FINAL_CONFIDENCE_CALCULATION_LAST_STEP:
                    conf_acc = conf_out
        out = {"counts": counts, "keep_alive": keep_alive + [meta, conf_out, bott]}
        if cf.get("nms") == cf.NMS_OPENCV:
            rows, _ = native.group_rectangles(meta, n0, n, int(cf.get("nms_opencv_min_neighbors")), 0.2, k_dev=n_dev,
                                              count=counts[n_st:n_st + 1])
            out["rows"] = rows
        else:
            out["meta"], out["conf"] = meta, conf_out
        return out

    def _result_from_nosync(self, counts: np.ndarray, rows: Optional[np.ndarray], meta: Optional[np.ndarray],
                            conf: Optional[np.ndarray], n0: int) -> GroupResult:
        n_st = len(self.stages)
        res = GroupResult(np.zeros((0, 4), np.int32), np.zeros((0,), np.float32), np.zeros((0,), np.int32), n_windows=n0)
        n_in = n0
        for i in range(n_st):
            if n_in < 1:
                break                                                       # the reference stops at the first empty stage
            res.stage_in.append(n_in); res.stage_kept.append(int(counts[i]))
            n_in = int(counts[i])
        k_last = int(counts[n_st - 1])
        if k_last == 0:
            return res
        if rows is not None:
            kk = int(counts[n_st])
            r = rows[:kk]
            res.frame = r[:, 0].copy()
            res.boxes = np.column_stack([r[:, 1], r[:, 2], r[:, 1] + r[:, 3], r[:, 2] + r[:, 4]]).astype(np.int32)
            res.conf = r[:, 5].astype(np.float64)
        else:
            m = meta[:k_last]
            res.frame = m[:, 0].copy()
            res.boxes = m[:, 1:5].astype(np.int32)
            res.conf = self._finalise_confidences(conf[:k_last])
        return res

    def _cf_signature(self):
        t = cf.get("foreground_confidence_threshold")
        return (tuple(t) if isinstance(t, (list, tuple)) else float(t), cf.get("nms"), int(cf.get("nms_opencv_min_neighbors")),
                cf.get("final_confidence_calculation"), bool(cf.get("reuse_bottlenecks")), float(cf.get("window_scale_factor")),
                float(cf.get("min_window_length")))

    def run_frames_latency(self, frames: torch.Tensor) -> GroupResult:
        """``run_frames`` for small batches (one un-batched frame = BASELINE config 2 as written): the cascade is captured
        ONCE per (shape, frame count, cf settings) in a CUDA graph whose stages take their window counts from device
        memory, then every call is one H2D/D2D copy of the frames, one graph launch, one D2H of counts + result rows and
        ONE stream synchronisation -- instead of ~35 launches and 4 blocking count reads."""
        n, h, w, _ = frames.shape
        plan = self.plan_for(w, h)
        key = (n, h, w) + self._cf_signature()
        graphs = getattr(self, "_graphs", None)
        if graphs is None:
            graphs = self._graphs = {}
        g = graphs.get(key)
        n0 = n * plan.n_windows
        if g is None:
            if len(graphs) >= 8:
                graphs.clear()
            n_st = len(self.stages)
            ws_bytes = max(st.workspace_bytes(n0, native.IMPL_TENSOR) for st in self.stages)
            g = {"frames": torch.empty((n, h, w, 3), dtype=torch.uint8, device=self.device),
                 "ws": torch.empty((ws_bytes,), dtype=torch.uint8, device=self.device),
                 "h_counts": torch.empty((n_st + 1,), dtype=torch.int64).pin_memory()}
            # everything that allocates pinned memory or copies from pageable memory happens BEFORE the capture
            g["table"] = native.FrameTable.from_batch(g["frames"])
            if cf.get("nms") == cf.NMS_OPENCV:
                g["h_rows"] = torch.empty((n0, 6), dtype=torch.int32).pin_memory()
            else:
                g["h_meta"] = torch.empty((n0, 5), dtype=torch.int32).pin_memory()
                g["h_conf"] = torch.empty((n0,), dtype=torch.float32).pin_memory()
            g["frames"].copy_(frames)
            self._enqueue_nosync(g["frames"], plan, g["ws"], g["table"])   # eager warm-up: plan upload, kernel attributes
            torch.cuda.current_stream().synchronize()
            graph = torch.cuda.CUDAGraph()
            # thread_local: the app's feeder thread may allocate / copy on its own stream while this thread captures
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                out = self._enqueue_nosync(g["frames"], plan, g["ws"], g["table"])
                g["h_counts"].copy_(out["counts"], non_blocking=True)
                if "rows" in out:
                    g["h_rows"].copy_(out["rows"], non_blocking=True)
                else:
                    g["h_meta"].copy_(out["meta"], non_blocking=True)
                    g["h_conf"].copy_(out["conf"], non_blocking=True)
            g["graph"], g["out"] = graph, out
            graphs[key] = g
        g["frames"].copy_(frames, non_blocking=True)
        g["graph"].replay()
        torch.cuda.current_stream().synchronize()                       # the one host wait of the call
        return self._result_from_nosync(g["h_counts"].numpy(), g["h_rows"].numpy() if "h_rows" in g else None,
                                        g["h_meta"].numpy() if "h_meta" in g else None,
                                        g["h_conf"].numpy() if "h_conf" in g else None, n0)

    def run_frames(self, frames, keep_trace: bool = False, while_busy=None) -> GroupResult:
        """Whole path for ``frames`` ``[n,H,W,3]`` uint8 (numpy: copied H2D; CUDA tensor: used in
        place).  Raises ValueError like the reference when no window can be extracted."""
        with torch.cuda.device(self.device):
            if isinstance(frames, np.ndarray):
                frames = self.stage_host_frames([frames[i] for i in range(frames.shape[0])])
                frames = frames.to(self.device, non_blocking=True)      # pinned host batch -> one async H2D copy
                self._pinned_issued()
            elif not frames.is_cuda:
                frames = frames.to(self.device, non_blocking=True)
            assert frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[3] == 3
            frames = frames.contiguous()
            n, h, w, _ = frames.shape
            plan = self.plan_for(w, h)
            if plan.n_windows < 1:
                raise ValueError("Could not extract any windows from the given image")      # abstract_inference_app.py:168-169
            if not plan.fits:
                raise ValueError("xmax/ymax must not leave the image boundaries")           # rectangles.py:286-289
            if not keep_trace and not getattr(self, "_latency_broken", False) and self._latency_eligible(n * plan.n_windows):
                try:
                    return self.run_frames_latency(frames)
                except (ValueError, native._lib.RodcError):
                    raise
                except RuntimeError as e:
                    # stream capture can be refused by the environment (a tool that forbids it, a driver without graph
                    # support for an op): the stage-by-stage path below runs the same kernels and returns the same bytes
                    self._latency_broken = True
                    self._graphs = {}
                    from .utils import log
                    log.log("CUDA-graph latency path disabled for this engine ({}); using the stage-by-stage path".format(e))
            pyr = native.pyramid(plan, frames)
            meta0 = native.window_meta(plan, n, 0, self.device)
            table = native.FrameTable.from_batch(frames)
            src0 = native.make_source(plan=plan, pyr=pyr)
            origin = None
            if len(self.stages) > 1 and not self.increasing_dims:
                def _rows():
                    off = torch.empty((n * plan.n_windows,), dtype=torch.int64, device=self.device)
                    pitch = torch.empty((n * plan.n_windows,), dtype=torch.int32, device=self.device)
                    native.window_gather_rows(plan, n, 0, off, pitch)
                    return pyr, off, pitch
                origin = WindowOrigin(gather=_rows)
            elif len(self.stages) > 1 and self._scaled_branch_possible():
                origin = WindowOrigin(slot=native.window_levels(plan, n, 0, self.device),
                                      groups=[dict(frames=frames, frame_base=0, scales=plan.scale, width=plan.win)])
            return self._run_stages(src0, n * plan.n_windows, meta0, table, n, keep_trace, origin, while_busy)

    def upload_mixed(self, frames: Sequence[np.ndarray]) -> "MixedFrames":
        """Host frames of DIFFERENT shapes -> one flat device buffer (frames grouped by shape, one asynchronous
        H2D copy from the pinned staging buffer) plus the bookkeeping ``run_mixed_device`` needs."""
        with torch.cuda.device(self.device):
            groups: Dict[Tuple[int, int], List[int]] = {}
            for i, f in enumerate(frames):
                groups.setdefault((f.shape[0], f.shape[1]), []).append(i)
            order = [i for idx in groups.values() for i in idx]
            sizes = [frames[i].size for i in order]
            total = int(sum(sizes))
            shapes = [(frames[i].shape[0], frames[i].shape[1]) for i in order]
            # frames that already live in page-locked memory go straight to their slot of the flat device buffer
            tens = [frames[i] for i in order]
            if self._all_pinned(tens):
                flat = torch.empty((total,), dtype=torch.uint8, device=self.device)
                offs = np.concatenate([[0], np.cumsum(sizes[:-1])]).astype(np.int64).tolist() if sizes else []
                native.upload_frames([frames[i] for i in order], flat, offs)
                self._keepalive_host = tens
                return MixedFrames(flat=flat, offs=offs, shapes=shapes, groups={k: len(v) for k, v in groups.items()}, order=order)
            # otherwise: one pinned staging buffer with all frames (group order) -> one H2D copy
            buf = self._pinned_stage(total)
            host = buf[:total].numpy()
            offs, o = [], 0
            for i, sz in zip(order, sizes):
                host[o:o + sz] = frames[i].reshape(-1)
                offs.append(o); o += sz
            flat = buf[:total].to(self.device, non_blocking=True)
            self._pinned_issued()
            return MixedFrames(flat=flat, offs=offs, shapes=shapes, groups={k: len(v) for k, v in groups.items()}, order=order)

    def run_mixed(self, frames: Sequence[np.ndarray], keep_trace: bool = False) -> Tuple[GroupResult, List[int]]:
        """Whole path for host frames of DIFFERENT shapes in one set of stage launches (the reference's
        merge mode concatenates the windows of all images, `app/abstract_inference_app.py:66-81`).
        Returns the result with INTERNAL frame ids (grouped by shape) and ``order``: internal id -> index
        into ``frames``.  Every frame must yield windows (callers filter with ``plan_for``)."""
        mf = self.upload_mixed(frames)
        return self.run_mixed_device(mf, keep_trace), mf.order

    def run_mixed_device(self, mf: "MixedFrames", keep_trace: bool = False, while_busy=None) -> GroupResult:
        """``run_mixed`` for frames that are already resident (``upload_mixed``).  One K1 launch pair per shape
        writes into a shared pyramid buffer; stage 0 reads its windows through a per-window (offset, pitch)
        table, later stages re-crop from a table of all frames."""
        with torch.cuda.device(self.device):
            flat, offs = mf.flat, mf.offs
            plans = {k: self.plan_for(k[1], k[0]) for k in mf.groups}
            for k, p in plans.items():
                if p.n_windows < 1:
                    raise ValueError("Could not extract any windows from the given image")
                if not p.fits:
                    raise ValueError("xmax/ymax must not leave the image boundaries")
            table = native.FrameTable(flat, offs, mf.shapes)
            n_windows = sum(n * plans[k].n_windows for k, n in mf.groups.items())
            pyr_bytes = sum(n * plans[k].pyramid_bytes for k, n in mf.groups.items())
            pyr = torch.empty((pyr_bytes,), dtype=torch.uint8, device=self.device)
            win_off = torch.empty((n_windows,), dtype=torch.int64, device=self.device)
            win_pitch = torch.empty((n_windows,), dtype=torch.int32, device=self.device)
            meta0 = torch.empty((n_windows, 5), dtype=torch.int32, device=self.device)
            want_slots = len(self.stages) > 1 and self.increasing_dims and self._scaled_branch_possible()
            slot0 = torch.empty((n_windows,), dtype=torch.int32, device=self.device) if want_slots else None
            slot_groups: List[dict] = []
            fbase = wbase = pbase = sbase = 0
            pos = 0
            # ONE horizontal-pass scratch for all shape groups (their K1 launches run one after the other on this stream)
            scratch_all = torch.empty((max(n * plans[k].scratch_bytes for k, n in mf.groups.items()),), dtype=torch.uint8,
                                      device=self.device)
            for k, n in mf.groups.items():
                p = plans[k]
                fsz = k[0] * k[1] * 3
                d = flat[offs[pos]:offs[pos] + n * fsz].view(n, k[0], k[1], 3)
                native.pyramid(p, d, out=pyr[pbase:pbase + n * p.pyramid_bytes].view(n, p.pyramid_bytes),
                               scratch=scratch_all[:n * p.scratch_bytes].view(n, p.scratch_bytes))
                nw = n * p.n_windows
                native.window_gather_rows(p, n, pbase, win_off[wbase:wbase + nw], win_pitch[wbase:wbase + nw])
                native.window_meta(p, n, fbase, self.device, out=meta0[wbase:wbase + nw])
                if want_slots:
                    native.window_levels(p, n, sbase, self.device, out=slot0[wbase:wbase + nw])
                    slot_groups.append(dict(frames=d, frame_base=fbase, scales=p.scale, width=p.win))
                    sbase += p.n_levels
                fbase += n; wbase += nw; pbase += n * p.pyramid_bytes; pos += n
            src0 = native.make_source(gather_base=pyr, win_off=win_off, win_pitch=win_pitch)
            self._keepalive0 = (pyr, win_off, win_pitch, flat, table)
            origin = WindowOrigin(gather=lambda: (pyr, win_off, win_pitch), slot=slot0, groups=slot_groups)
            return self._run_stages(src0, n_windows, meta0, table, len(mf.order), keep_trace, origin, while_busy)

    def run_windows(self, windows_raw, meta: np.ndarray, frames: List[np.ndarray], keep_trace: bool = False,
                    scales: Optional[np.ndarray] = None, widths: Optional[np.ndarray] = None) -> GroupResult:
        """``run_inference_on_windows``: caller-extracted stage-0 windows ``[n,S,S,3]`` u8 plus their
        metadata rows ``(frame_id, xmin_norm, ymin_norm, xmax_norm, ymax_norm)``; ``frames`` are the
        original images the later stages re-crop from.  ``scales`` / ``widths`` (``Window.scale`` and
        ``Window.width`` per window) are needed only by the resize-then-crop branch."""
        with torch.cuda.device(self.device):
            wr = torch.from_numpy(np.ascontiguousarray(windows_raw, np.uint8)).to(self.device)
            m = torch.from_numpy(np.ascontiguousarray(meta, np.int32)).to(self.device)
            dev_frames = [torch.from_numpy(np.ascontiguousarray(f)).to(self.device) for f in frames]
            table = native.FrameTable.from_list(dev_frames)
            src0 = native.make_source(dense=wr)
            self._keepalive0 = (wr, m, table)
            n0, S = int(wr.shape[0]), int(wr.shape[1])
            origin = None
            if len(self.stages) > 1 and not self.increasing_dims:
                origin = WindowOrigin(gather=lambda: (
                    wr, torch.arange(n0, dtype=torch.int64, device=self.device) * (S * S * 3),
                    torch.full((n0,), S * 3, dtype=torch.int32, device=self.device)))
            elif len(self.stages) > 1 and self._scaled_branch_possible():
                if scales is None or widths is None:
                    raise ValueError("the resize-then-crop branch needs Window.scale and Window.width of every window")
                # one slot per distinct (frame, scale, width), one group per frame
                per_frame: Dict[int, List[Tuple[float, int]]] = {}
                slot_of = np.empty((n0,), np.int32)
                fids = np.asarray(meta)[:, 0].tolist()
                order_frames: List[int] = []
                for f in fids:
                    if f not in per_frame:
                        per_frame[f] = []; order_frames.append(f)
                local: Dict[Tuple[int, float, int], int] = {}
                for i, (f, sc, wd) in enumerate(zip(fids, np.asarray(scales, np.float64).tolist(), np.asarray(widths).tolist())):
                    k = (f, sc, int(wd))
                    if k not in local:
                        local[k] = len(per_frame[f]); per_frame[f].append((sc, int(wd)))
                bases, b = {}, 0
                for f in order_frames:
                    bases[f] = b; b += len(per_frame[f])
                for i, (f, sc, wd) in enumerate(zip(fids, np.asarray(scales, np.float64).tolist(), np.asarray(widths).tolist())):
                    slot_of[i] = bases[f] + local[(f, sc, int(wd))]
                groups = [dict(frames=dev_frames[f].view(1, *dev_frames[f].shape), frame_base=f,
                               scales=np.asarray([v[0] for v in per_frame[f]], np.float64),
                               width=np.asarray([v[1] for v in per_frame[f]], np.float64)) for f in order_frames]
                origin = WindowOrigin(slot=torch.from_numpy(slot_of).to(self.device), groups=groups)
            return self._run_stages(src0, n0, m, table, len(frames), keep_trace, origin)

    def softmax_on_raw(self, raw: np.ndarray) -> np.ndarray:
        """``run_inference_on_raw_data`` of the single net: softmax ``[n,2]`` for dense windows."""
        assert len(self.stages) == 1
        st = self.stages[0]
        with torch.cuda.device(self.device):
            wr = torch.from_numpy(np.ascontiguousarray(raw, np.uint8)).to(self.device)
            n = int(wr.shape[0])
            bott = torch.empty((n, st.width), dtype=torch.float32, device=self.device)
            impl = self._impl()
            wsb = st.workspace_bytes(n, impl)
            _, probs = native.stage_forward(st, native.make_source(dense=wr), 0, n, bott, self._workspace(wsb)[:wsb], impl)
            return probs.cpu().numpy()
