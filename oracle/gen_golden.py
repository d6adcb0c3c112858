"""Oracle (test infrastructure): golden-vector generator.

Runs the reference's own inference apps (under ``oracle/ref_shims.py``) on
seeded synthetic frames + seeded Xavier weights and writes small fixtures to
``tests/golden/``.  Must be run in the build container (needs
``/root/reference``):

    python -m oracle.gen_golden

What is reference-derived in each fixture: window table, scales, level sizes,
stage-0 raw windows, re-cropped stage-1/2 inputs (through the reference's
``cv2.resize`` call), survivor counts, candidate boxes, post-NMS boxes and
confidences.  What is NOT: the softmax values themselves (``Session.run`` is
the torch-fp32 restatement, see ``oracle/stage_net.py``).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)

from oracle import ref_shims  # noqa: E402
from oracle import cascade as ocascade  # noqa: E402
from rapidobjectdetectionusingcascadedcnns_b200 import synth  # noqa: E402

GOLDEN = os.path.join(REPO, "tests", "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


CASES = {
    # name: dict(frames=[(H,W,seed)...], cf overrides, pass_rates or None, single flag)
    "cascade_small": dict(frames=[(120, 160, 0), (96, 128, 1)], cf=dict(min_window_length=0.2), rates=None),
    "cascade_small_reject": dict(frames=[(120, 160, 2), (130, 150, 3), (100, 100, 4)],
                                 cf=dict(min_window_length=0.2), rates=[0.4, 0.5, 0.5]),
    "cascade_640x480": dict(frames=[(480, 640, 0)], cf=dict(), rates=[0.3, 0.4, 0.5], store_sub=16),
    "cascade_factor12_enlarge": dict(frames=[(150, 200, 5)],
                                     cf=dict(min_window_length=0.15, window_scale_factor=1.2,
                                             vertically_enlarge_bboxes=True,
                                             foreground_confidence_threshold=[0.5, 0.6, 0.7],
                                             nms_opencv_min_neighbors=2),
                                     rates=[0.5, 0.6, 0.6]),
    "cascade_conf_avg": dict(frames=[(120, 160, 6)],
                             cf=dict(min_window_length=0.2,
                                     final_confidence_calculation="FINAL_CONFIDENCE_CALCULATION_AVG",
                                     nms="NMS_DISABLED"), rates=[0.5, 0.5, 0.5]),
    "cascade_conf_mult": dict(frames=[(120, 160, 6)],
                              cf=dict(min_window_length=0.2,
                                      final_confidence_calculation="FINAL_CONFIDENCE_CALCULATION_MULT",
                                      nms="NMS_DISABLED"), rates=[0.5, 0.5, 0.5]),
    # next-stage input by resize-then-crop (`app/inference_cascade_app.py:374-394`)
    "cascade_scaled_patches": dict(frames=[(120, 160, 8), (96, 128, 9)],
                                   cf=dict(min_window_length=0.2, cascade_scale_patches_individually=False),
                                   rates=[0.4, 0.5, 0.5]),
    # ..._iff_not_cached with window_scale_factor 2: 2 * scale_k IS a cached level, so both branches are taken
    "cascade_iff_not_cached": dict(frames=[(200, 260, 10)],
                                   cf=dict(min_window_length=0.1, window_scale_factor=2.0,
                                           cascade_scale_patches_individually_iff_not_cached=True),
                                   rates=[0.6, 0.6, 0.6]),
    # all stages share one input size: the raw stage-0 windows are carried through (`:234-240`)
    "cascade_equal_dims": dict(frames=[(120, 160, 11)], cf=dict(min_window_length=0.2), rates=[0.4, 0.5, 0.5],
                               sizes=[24, 24, 24]),
    "single_small": dict(frames=[(200, 260, 7)], cf=dict(min_window_length=0.3), rates=[0.3], single=True),
}


def run_case(name: str, spec: dict):
    import config as cf
    from app.inference_cascade_app import InferenceCascadeApp
    from app.inference_app import InferenceApp
    from data.rectangles import Window

    single = spec.get("single", False)
    frames = [synth.synth_frame(h, w, s) for (h, w, s) in spec["frames"]]
    cf.reset()
    for k, v in spec["cf"].items():
        cf.set(k, v)
    ocfg = ocascade.Config(**{k: v for k, v in spec["cf"].items() if k in ocascade.Config().__dict__})
    if single:
        weights = [synth.xavier_stage_weights(2, 48)]
    elif spec.get("sizes"):
        weights = [synth.xavier_stage_weights(i, sz, n_prev_bottleneck=512 * i) for i, sz in enumerate(spec["sizes"])]
    else:
        weights = synth.xavier_cascade()
    offsets = [0.0] * len(weights)
    if spec["rates"] is not None:
        if single:
            # single net: calibrate the one stage on all windows
            offsets = ocascade.calibrate_offsets(frames, weights, ocfg, spec["rates"])
        else:
            offsets = ocascade.calibrate_offsets(frames, weights, ocfg, spec["rates"])
        for w, o in zip(weights, offsets):
            w["fc2_b"][1] += np.float32(o)
    ref_shims.write_model(name, weights if not single else weights[0], cascade=not single)
    infos = ref_shims.make_image_infos(frames, prefix=name)

    out = {}
    # --- window table straight from the reference's Window.extract_windows
    app = (InferenceApp if single else InferenceCascadeApp)(name)
    app._init_tf()
    tables, scales, raws = [], [], []
    for fi, info in enumerate(infos):
        app._update_input_dims()
        raw, winfo = Window.extract_windows(info, True)
        raws.append(raw)
        tables.append(np.asarray([[fi, w.xmin, w.ymin, w.xmin_norm, w.ymin_norm, w.xmax_norm, w.ymax_norm]
                                  for w in winfo], np.int32))
        scales.append(np.asarray([w.scale for w in winfo], np.float64))
        lv_scales = sorted(set(scales[-1].tolist()), reverse=True)
        lv_sizes = [info.raw_scaled(True, s).shape[:2] for s in lv_scales]
        out["level_scale_%d" % fi] = np.asarray(lv_scales, np.float64)
        out["level_hw_%d" % fi] = np.asarray(lv_sizes, np.int32)
        out["level_sha_%d" % fi] = np.asarray([sha(info.raw_scaled(True, s)) for s in lv_scales])
        info.clear_raw_img_cache()
    raw0 = np.concatenate(raws)
    out["win_table"] = np.concatenate(tables)
    out["win_scale"] = np.concatenate(scales)
    sub = spec.get("store_sub", 4)
    out["windows_raw_sha"] = np.asarray(sha(raw0))
    out["windows_raw_sub"] = raw0[::sub]
    out["sub"] = np.asarray(sub)

    # --- full run through the reference's app (merge mode like run_inference_cascade.py)
    ref_shims.TRACE.clear()
    results = app.run_inference_on_images(infos, merge=True)
    trace = list(ref_shims.TRACE)
    # group trace records per graph (stage) in call order
    stage_names = []
    for t in trace:
        if t["graph"] not in stage_names:
            stage_names.append(t["graph"])
    for si, g in enumerate(stage_names):
        X = np.concatenate([t["X"] for t in trace if t["graph"] == g])
        P = np.concatenate([t["probs"] for t in trace if t["graph"] == g])
        L = np.concatenate([t["logits"] for t in trace if t["graph"] == g])
        out["stage%d_n_in" % si] = np.asarray(X.shape[0])
        out["stage%d_X_sha" % si] = np.asarray(sha(X))
        out["stage%d_X_sub" % si] = X[::sub]
        out["stage%d_probs_f32" % si] = P
        out["stage%d_logits_f32" % si] = L
    assert (out["stage0_X_sha"] == out["windows_raw_sha"]), "merge-mode feed must equal extract_windows output"
    boxes = []
    key_to_idx = {info.path_original: i for i, info in enumerate(infos)}
    for per_img in results:
        for b in per_img:
            boxes.append([key_to_idx[b.image.path_original], float(b.xmin), float(b.ymin), float(b.xmax),
                          float(b.ymax), float(b.confidence)])
    out["final_boxes"] = np.asarray(boxes, np.float64).reshape(-1, 6)

    # --- candidates (pre-NMS) through the reference with NMS disabled
    if cf.get("nms") != cf.NMS_DISABLED:
        cf.set("nms", cf.NMS_DISABLED)
        enl = cf.get("vertically_enlarge_bboxes")
        cf.set("vertically_enlarge_bboxes", False)
        for info in infos:
            info.clear_raw_img_cache()
        res2 = app.run_inference_on_images(infos, merge=True)
        cands = [[key_to_idx[b.image.path_original], b.xmin, b.ymin, b.xmax, b.ymax, float(b.confidence)]
                 for per_img in res2 for b in per_img]
        out["cand_boxes"] = np.asarray(cands, np.float64).reshape(-1, 6)
        cf.set("vertically_enlarge_bboxes", enl)
    else:
        out["cand_boxes"] = out["final_boxes"].copy()
    app.clean()

    meta = dict(name=name, frames=spec["frames"], cf=spec["cf"], offsets=offsets, single=single, sizes=spec.get("sizes"),
                n_windows=int(raw0.shape[0]),
                survivors=[int(out["stage%d_n_in" % i]) for i in range(len(stage_names))],
                n_final=int(out["final_boxes"].shape[0]), n_cand=int(out["cand_boxes"].shape[0]))
    out["meta"] = np.asarray(json.dumps(meta))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, meta["n_windows"], meta["survivors"], "cand", meta["n_cand"], "final", meta["n_final"],
          "offsets", offsets, flush=True)


def gen_third_party_vectors():
    """Golden vectors for the third-party functions straight from the installed libraries."""
    import cv2
    from PIL import Image
    rs = np.random.RandomState(1234)
    out = {}
    # cv2.resize
    for k in range(24):
        sh, sw = rs.randint(5, 90, 2)
        d = int(rs.choice([12, 24, 48]))
        if k % 6 == 0:
            sh = sw = 2 * d
        if k % 6 == 1:
            sh, sw = 2 * d, 2 * d + 1
        src = rs.randint(0, 256, (sh, sw, 3)).astype(np.uint8)
        out["cvr_src_%d" % k] = src
        out["cvr_dst_%d" % k] = cv2.resize(src, (d, d))
    # Pillow resize
    img = synth.synth_frame(97, 131, 77)
    out["pil_src"] = img
    for k, (w, h) in enumerate([(31, 23), (131, 50), (60, 97), (12, 9), (130, 96), (200, 150)]):
        out["pil_dst_%d" % k] = np.asarray(Image.fromarray(img).resize((w, h), resample=2))
    # groupRectangles
    for k in range(16):
        n = int(rs.randint(1, 120))
        base = rs.randint(0, 300, (int(rs.randint(1, 8)), 2))
        rects = []
        for _ in range(n):
            b = base[rs.randint(len(base))]
            s = int(rs.randint(12, 90))
            rects.append((int(b[0] + rs.randint(-8, 9)), int(b[1] + rs.randint(-8, 9)),
                          s + int(rs.randint(-4, 5)), s + int(rs.randint(-4, 5))))
        thr = int(k % 4)
        r, w = cv2.groupRectangles(list(rects), thr)
        out["gr_in_%d" % k] = np.asarray(rects, np.int32)
        out["gr_thr_%d" % k] = np.asarray(thr)
        out["gr_out_%d" % k] = np.asarray(r, np.int32).reshape(-1, 4)
        out["gr_w_%d" % k] = np.asarray(w, np.int32).reshape(-1)
    np.savez_compressed(os.path.join(GOLDEN, "third_party.npz"), **out)
    print("third_party vectors written")


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    ref_shims.install()
    only = sys.argv[1:]
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        run_case(name, spec)
    if not only:
        gen_third_party_vectors()


if __name__ == "__main__":
    main()
