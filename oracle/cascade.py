"""Oracle (test infrastructure): the cascade-inference path end to end on the CPU.

Restates
* window extraction `data/rectangles.py:309-435` (+ per-scale level cache
  `data/image_info.py:201-250`),
* merge-mode orchestration `app/abstract_inference_app.py:38-156`,
* the cascade loop `app/inference_cascade_app.py:119-342` and the next-stage input
  `:346-401`: crop + ``cv2.resize`` (default, `:363-373`), resize-then-crop
  (``cascade_scale_patches_individually=False`` / ``..._iff_not_cached``, `:374-394`),
  and the unchanged raw windows of cascades with one input size (`:234-240`),
* the single-net path `app/inference_app.py:117-166`,
* ``_run_nms`` / ``_postprocess_bboxes`` `app/inference_app.py:169-231` and
  ``vertically_enlarge_bboxes`` `data/rectangles.py:185-220`.

Third-party arithmetic (Pillow resize, cv2.resize, cv2.groupRectangles) is
taken either from the oracle's integer restatements (``arith="restated"``) or,
for speed in the CPU baseline, from the installed libraries (``arith="lib"``);
``tests/`` pin the two against each other.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import cv_resize, group_rectangles as gr, pillow_resize, plan as planmod, stage_net


@dataclass
class Config:
    """The `cf` keys the path reads, reference defaults."""
    foreground_confidence_threshold: object = 0.5       # `config.py:73` scalar or per-stage list
    min_window_length: float = 0.075                    # :84
    window_scale_factor: float = 1.1                    # :92
    nms: str = "NMS_OPENCV"                             # :95-97
    nms_opencv_min_neighbors: int = 1                   # :101
    vertically_enlarge_bboxes: bool = False             # :109
    final_confidence_calculation: str = "FINAL_CONFIDENCE_CALCULATION_LAST_STEP"   # :118-122
    max_batch_size_original: int = 1200                 # :133,582
    reuse_bottlenecks: bool = True                      # :277
    cascade_scale_patches_individually: bool = True     # :351
    cascade_scale_patches_individually_iff_not_cached: bool = False   # :358


MIN_SCORE_MULT = 0.5001   # `config.py:121`


# ----------------------------------------------------------------------------- extraction
def _level_image(frame: np.ndarray, lvl: planmod.Level, arith: str) -> np.ndarray:
    if lvl.scale == 1.0:
        return frame
    if arith == "lib":
        from PIL import Image
        return np.asarray(Image.fromarray(frame).resize((lvl.width, lvl.height), resample=2))
    return pillow_resize.resize_bilinear(frame, lvl.width, lvl.height)


def pyramid_levels(frame: np.ndarray, p: planmod.Plan, arith: str = "restated") -> List[np.ndarray]:
    return [_level_image(frame, lvl, arith) for lvl in p.levels]


def extract_windows(frame: np.ndarray, win: int, cfg: Config, arith: str = "restated"):
    """``Window.extract_windows``: returns ``(windows_raw [N,win,win,3] u8, plan)``.
    Raises ValueError like ``Window.raw`` when a window leaves its level."""
    h, w = frame.shape[:2]
    p = planmod.plan_pyramid(w, h, win, cfg.window_scale_factor, cfg.min_window_length)
    if not planmod.windows_fit(p):
        raise ValueError("xmax/ymax must not leave the image boundaries")
    levels = pyramid_levels(frame, p, arith)
    raw = np.empty((p.n_windows, win, win, 3), np.uint8)
    for i in range(p.n_windows):                       # per-window slicing like `data/rectangles.py:409`
        lv = levels[p.win_level[i]]
        x, y = p.win_x[i], p.win_y[i]
        raw[i] = lv[y:y + win, x:x + win]
    return raw, p


# ----------------------------------------------------------------------------- recrop
def recrop(frame: np.ndarray, norm_box: Sequence[int], size: int, arith: str = "restated") -> np.ndarray:
    """`app/inference_cascade_app.py:363-373`: ``cv2.resize(window.raw_norm, (size, size))``."""
    x0, y0, x1, y1 = (int(v) for v in norm_box)
    if y1 > frame.shape[0] or x1 > frame.shape[1]:     # `data/rectangles.py:302-305`
        raise ValueError("norm box must not leave the image boundaries")
    crop = frame[y0:y1, x0:x1]
    if arith == "lib":
        import cv2
        return cv2.resize(crop, (size, size))
    return cv_resize.cv_resize_linear(np.ascontiguousarray(crop), size, size)


class ScaledCache:
    """``ImageInfo._raw_img_cache`` of one frame for the keys ``orig_scaled_<ratio>``
    (`data/image_info.py:201-258`): ``raw_scaled(True, ratio)`` resizes the ORIGINAL with
    ``scipy.misc.imresize(arr, ratio)`` -> Pillow BILINEAR to ``(int(W*ratio), int(H*ratio))``."""

    def __init__(self, frame: np.ndarray, arith: str = "restated"):
        self.frame, self.arith, self.store, self.keys = frame, arith, {}, set()

    def is_cached(self, ratio: float) -> bool:              # `:256-258`
        return "orig_scaled_{}".format(ratio) in self.keys

    def raw_scaled(self, ratio: float) -> np.ndarray:       # `:201-213`, always called with cache=True on this path
        if ratio == 1.0:
            return self.frame
        key = "orig_scaled_{}".format(ratio)
        self.keys.add(key)
        if key not in self.store:
            h, w = self.frame.shape[:2]
            lw, lh = int(w * ratio), int(h * ratio)
            if self.arith == "lib":
                from PIL import Image
                self.store[key] = np.asarray(Image.fromarray(self.frame).resize((lw, lh), resample=2))
            else:
                self.store[key] = pillow_resize.resize_bilinear(self.frame, lw, lh)
        return self.store[key]


def recrop_scaled(cache: ScaledCache, norm_box: Sequence[int], scale: float, width: int, size: int) -> np.ndarray:
    """`app/inference_cascade_app.py:358-359,374-394`: crop ``size x size`` at
    ``(int(xmin_norm * next_scale), int(ymin_norm * next_scale))`` of the frame resized to
    ``next_scale = window.scale * (size / window.width)``; ``Window.raw`` bounds check `data/rectangles.py:286-289`."""
    next_scale = scale * (size / width)
    nx = int(int(norm_box[0]) * next_scale)
    ny = int(int(norm_box[1]) * next_scale)
    img = cache.raw_scaled(next_scale)
    if ny + size > img.shape[0]:
        raise ValueError("ymax must not leave the image boundaries")
    if nx + size > img.shape[1]:
        raise ValueError("xmax must not leave the image boundaries")
    return img[ny:ny + size, nx:nx + size]


def next_stage_window(frame: np.ndarray, cache: ScaledCache, norm_box, scale: float, width: int, size: int,
                      cfg: "Config", arith: str) -> np.ndarray:
    """``_get_increased_raw_window_for_next_net`` `:346-401` with its branch condition `:362-365`."""
    next_scale = scale * (size / width)
    if cfg.cascade_scale_patches_individually and \
            (not cfg.cascade_scale_patches_individually_iff_not_cached or not cache.is_cached(next_scale)):
        return recrop(frame, norm_box, size, arith)
    return recrop_scaled(cache, norm_box, scale, width, size)


# ----------------------------------------------------------------------------- nms / post
def run_nms(boxes: np.ndarray, image_ids: np.ndarray, cfg: Config, arith: str = "restated"):
    """``_run_nms``: boxes int [K,4] (xmin,ymin,xmax,ymax).  Returns
    ``(boxes [K',4], conf [K'], image_ids [K'])`` in the reference's order: images by
    first appearance among the candidates, then class order."""
    out_b, out_c, out_i = [], [], []
    if boxes.shape[0] == 0:
        return np.zeros((0, 4), np.int64), np.zeros((0,), np.float64), np.zeros((0,), np.int64)
    seen: Dict[int, List[int]] = {}
    for k, iid in enumerate(image_ids.tolist()):
        seen.setdefault(iid, []).append(k)
    for iid, idx in seen.items():
        b = boxes[idx]
        rects = [(int(r[0]), int(r[1]), int(r[2] - r[0]), int(r[3] - r[1])) for r in b]
        if arith == "lib":
            import cv2
            rr, ww = cv2.groupRectangles(rects, cfg.nms_opencv_min_neighbors)
            rr = np.asarray(rr, np.int32).reshape(-1, 4); ww = np.asarray(ww, np.int32).reshape(-1)
        else:
            rr, ww = gr.group_rectangles(rects, cfg.nms_opencv_min_neighbors)
        for r, wgt in zip(rr, ww):
            out_b.append((r[0], r[1], r[0] + r[2], r[1] + r[3]))
            out_c.append(float(wgt)); out_i.append(iid)
    return (np.asarray(out_b, np.int64).reshape(-1, 4), np.asarray(out_c, np.float64),
            np.asarray(out_i, np.int64))


def vertically_enlarge(boxes: np.ndarray, enlarge_top: float = 0.2) -> np.ndarray:
    """`data/rectangles.py:185-220` defaults: only the top edge moves; float result."""
    b = boxes.astype(np.float64).copy()
    h = b[:, 3] - b[:, 1]
    b[:, 1] = np.maximum(b[:, 1] - enlarge_top * h, 0)
    return b


def postprocess(boxes, conf, image_ids, cfg: Config, arith: str = "restated"):
    """``_postprocess_bboxes``."""
    if cfg.nms != "NMS_DISABLED":
        boxes, conf, image_ids = run_nms(boxes, image_ids, cfg, arith)
    boxes = boxes.astype(np.float64) if cfg.vertically_enlarge_bboxes else boxes
    if cfg.vertically_enlarge_bboxes:
        boxes = vertically_enlarge(boxes)
    return boxes, conf, image_ids


# ----------------------------------------------------------------------------- cascade
@dataclass
class StageTrace:
    n_in: int
    probs: np.ndarray            # [n_in, 2] softmax of this stage (dtype of the run)
    logits: np.ndarray           # [n_in, 2]
    keep_idx: np.ndarray         # indices (into this stage's input order) of the survivors
    windows_raw: Optional[np.ndarray] = None     # [n_in, S, S, 3] u8 input of this stage


@dataclass
class CascadeResult:
    n_windows: int
    stages: List[StageTrace] = field(default_factory=list)
    cand_boxes: np.ndarray = None      # pre-NMS [K,4] norm boxes
    cand_conf: np.ndarray = None
    cand_image: np.ndarray = None
    boxes: np.ndarray = None           # post-processing
    conf: np.ndarray = None
    image: np.ndarray = None
    orig_index: np.ndarray = None      # for each final-stage survivor: index in the stage-0 window list


def _threshold(cfg: Config, stage: int, n_stages: int) -> float:
    t = cfg.foreground_confidence_threshold
    if isinstance(t, (int, float)):
        return float(t)
    if len(t) == n_stages:
        return float(t[stage])
    raise ValueError("Invalid foreground_confidence_threshold.")


def run_cascade(frames: Sequence[np.ndarray], stage_weights: Sequence[Dict[str, np.ndarray]],
                cfg: Optional[Config] = None, *, dtype=torch.float64, arith: str = "restated",
                keep_raw: bool = False) -> CascadeResult:
    """Merge-mode cascade over ``frames`` (`app/abstract_inference_app.py:56-108` +
    `app/inference_cascade_app.py:119-342`).  Frames whose extraction raises are skipped
    like `app/abstract_inference_app.py:165-183`."""
    cfg = cfg or Config()
    n_stages = len(stage_weights)
    sizes = [int(round((w["fc1_W"].shape[0] // w["conv%d_W" % (stage_net.n_conv_layers(w) - 1)].shape[3]) ** 0.5))
             for w in stage_weights]
    raws, norms, img_ids, wscales = [], [], [], []
    caches: Dict[int, ScaledCache] = {}
    for fi, frame in enumerate(frames):
        try:
            raw, p = extract_windows(frame, sizes[0], cfg, arith)
            if raw.shape[0] < 1:
                raise ValueError("Could not extract any windows from the given image")
        except ValueError:
            continue
        raws.append(raw); norms.append(p.norm); img_ids.append(np.full(p.n_windows, fi, np.int64))
        lv_scale = np.asarray([lv.scale for lv in p.levels], np.float64)
        wscales.append(lv_scale[p.win_level])
        caches[fi] = ScaledCache(frame, arith)
        for li, lv in enumerate(p.levels):         # Window.raw cached every level it touched (`data/rectangles.py:281`)
            if lv.scale != 1.0 and (p.win_level == li).any():
                caches[fi].keys.add("orig_scaled_{}".format(lv.scale))         # pixels are rebuilt on first use
    if not raws:
        raise ValueError("Could not extract any windows from the given image")
    windows_raw = np.concatenate(raws); norm = np.concatenate(norms); image = np.concatenate(img_ids)
    wscale = np.concatenate(wscales)
    orig = np.arange(windows_raw.shape[0])
    res = CascadeResult(n_windows=windows_raw.shape[0])
    bott_in = None
    conf_acc = None
    np_dtype = np.float64 if dtype == torch.float64 else np.float32
    final = None
    for i in range(n_stages):
        if windows_raw.shape[0] < 1:
            break                                                       # :140-142
        batch = cfg.max_batch_size_original * (4 ** (n_stages - i - 1)) if sizes[0] < sizes[-1] \
            else cfg.max_batch_size_original                            # :150-163
        t = _threshold(cfg, i, n_stages)
        use_b = i > 0 and cfg.reuse_bottlenecks
        probs, bott, logits = stage_net.stage_forward_batched(
            windows_raw, stage_weights[i], bott_in if use_b else None, dtype=dtype, batch=batch)
        p_fg = probs[:, 1].numpy()
        # strict compare; float32 probability promoted to float64 (NumPy 1.14 scalar rules)
        keep = np.nonzero(p_fg.astype(np.float64) > t)[0]               # :219-222
        if conf_acc is None:
            conf_new = p_fg[keep].astype(np.float32)
        elif cfg.final_confidence_calculation == "FINAL_CONFIDENCE_CALCULATION_AVG":
            conf_new = (conf_acc[keep] + p_fg[keep].astype(np.float32)).astype(np.float32)
        else:
            conf_new = (conf_acc[keep] * p_fg[keep].astype(np.float32)).astype(np.float32)
        res.stages.append(StageTrace(windows_raw.shape[0], probs.numpy().astype(np_dtype),
                                     logits.numpy().astype(np_dtype), keep,
                                     windows_raw if keep_raw else None))
        norm, image, orig, wscale = norm[keep], image[keep], orig[keep], wscale[keep]
        last = i == n_stages - 1
        if not last:
            bott_in = bott[keep] if cfg.reuse_bottlenecks else None
            nxt = sizes[i + 1]
            if nxt != sizes[i]:
                new_raw = np.empty((keep.shape[0], nxt, nxt, 3), np.uint8)
                for k in range(keep.shape[0]):                          # :301-304
                    new_raw[k] = next_stage_window(frames[image[k]], caches[int(image[k])], norm[k], float(wscale[k]),
                                                   sizes[0], nxt, cfg, arith)
                windows_raw = new_raw
            else:
                windows_raw = windows_raw[keep]
            if cfg.final_confidence_calculation != "FINAL_CONFIDENCE_CALCULATION_LAST_STEP":
                conf_acc = conf_new
        else:
            if cfg.final_confidence_calculation == "FINAL_CONFIDENCE_CALCULATION_AVG":
                conf_new = (conf_new / np.float32(n_stages)).astype(np.float32)
            elif cfg.final_confidence_calculation == "FINAL_CONFIDENCE_CALCULATION_MULT":
                conf_new = conf_new.copy(); conf_new[conf_new < MIN_SCORE_MULT] = MIN_SCORE_MULT
            final = conf_new
    if final is None:          # cascade stopped early: no candidates
        norm = np.zeros((0, 4), np.int32); image = np.zeros((0,), np.int64); final = np.zeros((0,), np.float32)
        orig = np.zeros((0,), np.int64)
    res.cand_boxes, res.cand_conf, res.cand_image, res.orig_index = norm, final, image, orig
    res.boxes, res.conf, res.image = postprocess(norm.astype(np.int64), final.astype(np.float64), image, cfg, arith)
    return res


def run_single(frames: Sequence[np.ndarray], weights: Dict[str, np.ndarray], cfg: Optional[Config] = None,
               *, dtype=torch.float64, arith: str = "restated") -> CascadeResult:
    """Single-net path `app/inference_app.py:117-166`: keep ``argmax != background``,
    confidence = p[argmax]."""
    cfg = cfg or Config()
    size = int(round((weights["fc1_W"].shape[0] // weights["conv0_W"].shape[3]) ** 0.5))
    raws, norms, img_ids = [], [], []
    for fi, frame in enumerate(frames):
        try:
            raw, p = extract_windows(frame, size, cfg, arith)
        except ValueError:
            continue
        raws.append(raw); norms.append(p.norm); img_ids.append(np.full(p.n_windows, fi, np.int64))
    windows_raw = np.concatenate(raws); norm = np.concatenate(norms); image = np.concatenate(img_ids)
    probs, _, logits = stage_net.stage_forward_batched(windows_raw, weights, None, dtype=dtype,
                                                       batch=cfg.max_batch_size_original)
    pn = probs.numpy()
    arg = pn.argmax(axis=1)
    keep = np.nonzero(arg != 0)[0]
    res = CascadeResult(n_windows=windows_raw.shape[0])
    res.stages.append(StageTrace(windows_raw.shape[0], pn, logits.numpy(), keep))
    conf = pn[keep, arg[keep]]
    res.cand_boxes, res.cand_conf, res.cand_image, res.orig_index = norm[keep], conf, image[keep], keep
    res.boxes, res.conf, res.image = postprocess(norm[keep].astype(np.int64), conf.astype(np.float64), image[keep], cfg, arith)
    return res


# ----------------------------------------------------------------------------- calibration
def calibrate_offsets(frames, stage_weights, cfg: Optional[Config], pass_rates: Sequence[float],
                      arith: str = "lib") -> List[float]:
    """Documented synthetic-weight knob: random-init cascades pass
    ~99 % of windows, so for configs that need a target pass rate an offset is added to
    ``fc2.b[1]`` per stage.  Offsets are rounded to 1/64 so they are exact in float32 and
    reproducible from the JSON metadata.  Returns the offsets; ``stage_weights`` is not modified."""
    cfg = cfg or Config()
    offs: List[float] = []
    ws = [dict(w) for w in stage_weights]
    for i, rate in enumerate(pass_rates):
        for k, o in enumerate(offs):
            b = stage_weights[k]["fc2_b"].copy(); b[1] += np.float32(o); ws[k]["fc2_b"] = b
        res = _run_prefix(frames, ws, cfg, i, arith)
        margin = res[:, 1] - res[:, 0]
        q = float(np.quantile(margin, 1.0 - rate))
        offs.append(-round(q * 64.0) / 64.0)
    return offs


def _run_prefix(frames, ws, cfg, upto: int, arith: str) -> np.ndarray:
    """Logits (fp64) of stage ``upto`` on the survivors of stages ``< upto``."""
    sub = Config(**{**cfg.__dict__})
    n = len(ws)
    sizes = [int(round((w["fc1_W"].shape[0] // w["conv%d_W" % (stage_net.n_conv_layers(w) - 1)].shape[3]) ** 0.5))
             for w in ws]
    raws, norms, ids = [], [], []
    for fi, fr in enumerate(frames):
        raw, p = extract_windows(fr, sizes[0], sub, arith)
        raws.append(raw); norms.append(p.norm); ids.append(np.full(p.n_windows, fi))
    raw = np.concatenate(raws); norm = np.concatenate(norms); image = np.concatenate(ids)
    bott = None
    for i in range(upto + 1):
        probs, b, logits = stage_net.stage_forward_batched(raw, ws[i], bott if (i > 0 and sub.reuse_bottlenecks) else None,
                                                           dtype=torch.float64)
        if i == upto:
            return logits.numpy()
        keep = np.nonzero(probs[:, 1].numpy() > _threshold(sub, i, n))[0]
        norm, image, bott = norm[keep], image[keep], b[keep]
        raw = np.stack([recrop(frames[image[k]], norm[k], sizes[i + 1], arith) for k in range(len(keep))]) \
            if sizes[i + 1] != sizes[i] else raw[keep]
    raise AssertionError
