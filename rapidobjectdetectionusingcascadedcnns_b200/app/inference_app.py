"""Single-net inference app (mirror of `app/inference_app.py:19-300`).

``InferenceApp(model_session_key)`` looks for ``graph_<key>.pb`` / ``.npz`` in
``cf.get("output_graph_dir")``; ``stage_models=`` injects in-memory models instead (synthetic
weights).  The TensorFlow session is replaced by the device engine; ``_init_tf`` keeps its
name because callers (`app/abstract_inference_app.py:48`) use it.
"""
import os
from typing import List, Optional

import numpy as np

from .. import config as cf
from ..data.db import label
from ..data.box_list import BoxList
from ..data.rectangles import LabeledBoundingBox, Window
from ..network.stage_model import StageModel, model_paths
from ..utils import log
from ..utils.time_watcher import TimeWatcher
from .abstract_inference_app import AbstractInferenceApp


class InferenceApp(AbstractInferenceApp):
    _single = True

    def __init__(self, model_session_key=None, stage_models: Optional[List[StageModel]] = None, device=None):
        self._tf_initialized = False
        self._engine = None
        self._device = device
        self._models = list(stage_models) if stage_models is not None else None
        if model_session_key is None:
            model_session_key = cf.get("default_evaluation_model_single") if self._single \
                else cf.get("default_evaluation_model_cascade")
        self._model_session_key = model_session_key
        self._model_full_path = os.path.join(cf.get("output_graph_dir"), "graph_{}.pb".format(model_session_key))
        log.log("Initializing inference app with the following settings:")
        log.log(" - pre-trained model: {}".format(model_session_key))
        log.log(" - min_window_length: {}".format(cf.get("min_window_length")))
        log.log(" - window_scale_factor: {}".format(cf.get("window_scale_factor")))
        log.log(" - nms: {}".format(cf.get("nms")))
        if cf.get("nms") == cf.NMS_OPENCV:
            log.log("   -> nms_opencv_min_neighbors: {}".format(cf.get("nms_opencv_min_neighbors")))

    # ------------------------------------------------------------------ model files
    @property
    def model_full_path(self):
        return self._model_full_path

    def _stage_files(self) -> List[str]:
        for p in model_paths(cf.get("output_graph_dir"), self._model_session_key, None):
            if os.path.isdir(p):
                raise ValueError("The graph file path does not point to a file, but to a folder: {}".format(p))
            if os.path.exists(p):
                return [p]
        raise FileNotFoundError("Could not find the required graph file: {}".format(self.model_full_path))

    def _load_models(self) -> List[StageModel]:
        if self._models is None:
            self._models = [StageModel.load(p) for p in self._stage_files()]
        return self._models

    @staticmethod
    def _check_supported_config():
        """Options of the reference that this path does not implement fail loudly instead of being ignored."""
        if cf.get("append_inception"):
            raise NotImplementedError("the Inception splice needs a network download and is out of scope")

    def _extract_windows(self, img, convert_raw_to_np=True):
        """`:297-300`: the window size comes from the loaded model, then the common extraction."""
        self._update_input_dims()
        return super()._extract_windows(img, convert_raw_to_np)

    def _init_tf(self):
        """Load the stage model(s) and create the device engine (replaces `:93-115`)."""
        self._check_supported_config()
        if not self._tf_initialized:
            from ..engine import CascadeEngine
            models = self._load_models()
            geos = [dict(in_size=m.in_size, conv_stride=m.conv_stride, pool_size=m.pool_size, pool_stride=m.pool_stride)
                    for m in models]
            self._engine = CascadeEngine([m.weights for m in models], device=self._device, single=self._single,
                                         geometries=geos)
            self._tf_initialized = True

    def clean(self):
        """Release device resources (`:269-274`)."""
        self._tf_initialized = False
        self._engine = None

    @property
    def last_net(self):
        return True

    @property
    def supported_img_width(self):
        return self._load_models()[-1].supported_img_width

    @property
    def supported_img_height(self):
        return self._load_models()[-1].supported_img_height

    def _update_input_dims(self):
        """`:276-285`: the window size comes from the model, not from the current config."""
        cf.set("img_height_max", self.supported_img_height)
        cf.set("img_width_max", self.supported_img_width)
        cf.set("img_height", self.supported_img_height)
        cf.set("img_width", self.supported_img_width)

    def _windows_per_frame(self, w: int, h: int) -> int:
        self._init_tf()
        self._update_input_dims()
        plan = self._engine.plan_for(w, h)
        if plan.n_windows and not plan.fits:
            raise ValueError("xmax/ymax must not leave the image boundaries")          # `data/rectangles.py:286-289`
        return plan.n_windows

    # ------------------------------------------------------------------ results
    @staticmethod
    def _rows_from_result(res) -> np.ndarray:
        """``[K,5]`` float64 rows (xmin, ymin, xmax, ymax, confidence) incl. the optional vertical enlargement
        (`data/rectangles.py:185-220` defaults: ``ymin' = max(ymin - 0.2 height, 0)``)."""
        rows = np.empty((res.boxes.shape[0], 5), np.float64)
        rows[:, :4] = res.boxes
        rows[:, 4] = res.conf
        if cf.get("vertically_enlarge_bboxes") and rows.shape[0]:
            rows[:, 1] = np.maximum(rows[:, 1] - 0.2 * (rows[:, 3] - rows[:, 1]), 0)
        return rows

    @staticmethod
    def _boxes_from_result(res, infos, n_frames: int, flat: bool = False):
        """Per-frame ``BoxList`` views of the result arrays (objects are created on first use).  ``flat=True``
        returns ONE list over all frames in result order (``run_inference_on_windows`` without NMS)."""
        fg = label.get_by_iid(label.IID_FOREGROUND)
        post_nms = cf.get("nms") == cf.NMS_OPENCV
        enlarge = bool(cf.get("vertically_enlarge_bboxes"))
        frames = np.asarray(res.frame)
        if flat:
            bl = BoxList(res.boxes, res.conf, [infos[f] for f in frames.tolist()], fg, post_nms)
            return LabeledBoundingBox.vertically_enlarge_bboxes(bl) if enlarge and len(bl) else bl
        # frame ids ascend (K5 emits frame by frame, compaction is stable): one slice per frame
        cuts = np.searchsorted(frames, np.arange(n_frames + 1), side="left").tolist()
        out = []
        for f in range(n_frames):
            a, b = cuts[f], cuts[f + 1]
            if a == b:
                out.append([])
                continue
            bl = BoxList(res.boxes[a:b], res.conf[a:b], infos[f], fg, post_nms)
            out.append(LabeledBoundingBox.vertically_enlarge_bboxes(bl) if enlarge else bl)
        return out

    def _upload_frame_batch(self, frames):
        """Host frames -> device (on the current stream): one ``[n,H,W,3]`` tensor for one shape, a ``MixedFrames``
        record for several.  Separate from the run so that the app can upload chunk k + 1 while chunk k computes."""
        self._init_tf()
        if len({f.shape for f in frames}) == 1:
            return "same", self._engine.upload_host_frames(frames)
        return "mixed", self._engine.upload_mixed(frames)

    def _run_on_frame_batch(self, infos, frames) -> List[List[LabeledBoundingBox]]:
        return self._run_uploaded(infos, self._upload_frame_batch(frames))

    def _run_uploaded(self, infos, payload, as_rows: bool = False, while_busy=None):
        """Run one uploaded chunk.  Returns one result list per frame, or with ``as_rows`` the arrays
        ``(frame index inside the chunk [K], rows [K,5] = xmin, ymin, xmax, ymax, confidence)`` without any object."""
        self._init_tf()
        self._update_input_dims()
        kind, data = payload
        if kind == "same":
            res = self._engine.run_frames(data, while_busy=while_busy)
            self._log_stage_stats(res)
            if as_rows:
                return np.asarray(res.frame, np.int64), self._rows_from_result(res)
            return self._boxes_from_result(res, infos, len(infos))
        # several shapes: one merged set of stage launches (internal frame ids are grouped by shape)
        res, order = self._engine.run_mixed_device(data, while_busy=while_busy), data.order
        self._log_stage_stats(res)
        if as_rows:
            return np.asarray(order, np.int64)[np.asarray(res.frame, np.int64)], self._rows_from_result(res)
        per_internal = self._boxes_from_result(res, [infos[i] for i in order], len(infos))
        out = [[] for _ in infos]
        for internal, i in enumerate(order):
            out[i] = per_internal[internal]
        return out

    @staticmethod
    def _log_stage_stats(res):
        for i, (n_in, kept) in enumerate(zip(res.stage_in, res.stage_kept)):
            log.log(" -> net {}: used {:.2f}% of the original window set; kept {:.2f}% ({}/{}) as foreground".format(
                i + 1, n_in / max(res.n_windows, 1) * 100, kept / max(n_in, 1) * 100, kept, n_in))

    def run_inference_on_image(self, image) -> List[LabeledBoundingBox]:
        """`:233-267`.  Errors propagate to the caller (`abstract_inference_app.py:110-126`)."""
        self._init_tf()
        tw = TimeWatcher("Run inference on {}".format(getattr(image, "basename", image)))
        arr = np.ascontiguousarray(image.raw_original(True), dtype=np.uint8)
        res = self._run_on_frame_batch([image], [arr])[0]
        tw.stop()
        return res

    def run_inference_on_windows(self, windows_info: List[Window], windows_raw) -> List[LabeledBoundingBox]:
        """`:117-154` (single) / cascade `:119-342`: caller-provided windows of any images."""
        if len(windows_info) < 1:
            raise ValueError("The given window list must not be empty." if self._single
                             else "Could not extract any windows from the given image")
        if len(windows_info) != len(windows_raw):
            raise ValueError("The given window lists (raw vs. info) must have the same length.")
        self._init_tf()
        # frame ids in order of first appearance of the image (the reference groups NMS per
        # image.full_key in insertion order, `:187-193`)
        ids, infos, frames = {}, [], []
        meta = np.empty((len(windows_info), 5), np.int32)
        for k, wi in enumerate(windows_info):
            key = getattr(wi.image, "full_key", None) or id(wi.image)
            if key not in ids:
                ids[key] = len(infos)
                infos.append(wi.image)
                frames.append(np.ascontiguousarray(wi.image.raw_original(True), dtype=np.uint8))
            meta[k] = (ids[key], wi.xmin_norm, wi.ymin_norm, wi.xmax_norm, wi.ymax_norm)
        raw = np.ascontiguousarray(np.asarray(windows_raw, dtype=np.uint8))
        # Window.scale / Window.width feed the resize-then-crop branch (inference_cascade_app.py:358-359)
        scales = np.asarray([wi.scale for wi in windows_info], np.float64)
        widths = np.asarray([wi.width for wi in windows_info], np.int64)
        post_nms = cf.get("nms") == cf.NMS_OPENCV
        if post_nms and (np.diff(meta[:, 0]) < 0).any():    # the device NMS wants candidates grouped by frame
            order = np.argsort(meta[:, 0], kind="stable")
            meta, raw, scales, widths = meta[order], raw[order], scales[order], widths[order]
        res = self._engine.run_windows(raw, meta, frames, scales=scales, widths=widths)
        self._log_stage_stats(res)
        if not post_nms:
            # without NMS the reference returns the survivors in the order of the given windows, also when the
            # windows of several images are interleaved (`:135-154`, cascade `:319-334`)
            return self._boxes_from_result(res, infos, len(infos), flat=True)
        per_frame = self._boxes_from_result(res, infos, len(infos))
        return [b for boxes in per_frame for b in boxes]

    def run_inference_on_raw_data(self, raw_data: np.ndarray) -> np.ndarray:
        """`:156-166`: softmax ``[n,2]`` of dense windows, no batching logic, no metadata."""
        self._init_tf()
        return self._engine.softmax_on_raw(np.asarray(raw_data))
