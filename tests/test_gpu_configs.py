"""The other BASELINE.json configurations on the GPU.

config 3  FDDB-shaped batch: mixed ~450x400 frame sizes through the app (shape grouping)  -> full oracle compare
config 4  1920x1080, factor 1.2, min window 24 px (min_window_length 0.0125), 93 645 windows -> full oracle compare
config 5  3840x2160, 161 969 windows, > 99.9 % stage-1 rejection -> size-independent properties + sampled oracle
config 1  single 48-px net on 640x480 (3 319 windows)                                       -> oracle compare
Offsets on fc2.b[1] are the documented calibration knob; the same offsets feed both sides.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# north_star: survivor sets may differ from the fp64 oracle only for windows whose logit margin lies within 1e-4 of
# the threshold, at every stage
EPS = [1e-4, 1e-4, 1e-4]


@pytest.fixture()
def pkg():
    import rapidobjectdetectionusingcascadedcnns_b200 as p
    p.cf.reset()
    yield p
    p.cf.reset()


def _weights(offsets):
    from rapidobjectdetectionusingcascadedcnns_b200 import synth
    return synth.xavier_cascade(fg_bias_offsets=offsets)


def _calibrate(rates, cfg, seed=0, shape=(240, 320)):
    """Offsets from a small frame of the same recipe (cheap); exact rates are not needed."""
    from oracle import cascade as oc
    from rapidobjectdetectionusingcascadedcnns_b200 import synth
    small_cfg = oc.Config(**{**cfg.__dict__, "min_window_length": max(cfg.min_window_length, 0.075)})
    return oc.calibrate_offsets([synth.synth_frame(shape[0], shape[1], seed)], synth.xavier_cascade(), small_cfg, rates)


def _rows(results):
    return [np.asarray([[b.xmin, b.ymin, b.xmax, b.ymax, b.confidence] for b in per], np.float64).reshape(-1, 5)
            for per in results]


def _compare_frame(pkg, eng, frame, weights, ocfg, got_rows):
    """Final boxes of one frame against the fp64 oracle; on a mismatch the differing survivors must lie
    inside the epsilon band."""
    from oracle import cascade as oc
    r64 = oc.run_cascade([frame], weights, ocfg, dtype=torch.float64, arith="lib")
    ref = np.column_stack([r64.boxes, r64.conf]).astype(np.float64).reshape(-1, 5)
    if got_rows.shape == ref.shape and np.array_equal(got_rows, ref):
        return True
    res = eng.run_frames(frame[None], keep_trace=True)
    for i, st in enumerate(r64.stages):
        keep_gpu = res.trace["keep"][i].cpu().numpy()
        diff = np.setxor1d(keep_gpu, st.keep_idx)
        if diff.size:
            thr = ocfg.foreground_confidence_threshold
            thr = thr[i] if isinstance(thr, (list, tuple)) else thr
            margin = st.logits[diff, 1] - st.logits[diff, 0] - np.log(thr / (1 - thr))
            assert np.abs(margin).max() < EPS[i], "stage %d flip outside the band: %r" % (i, margin)
            return False
    raise AssertionError("boxes differ although every survivor set matches the oracle")


def test_config3_fddb_shaped_mixed_sizes(pkg):
    from oracle import cascade as oc
    from rapidobjectdetectionusingcascadedcnns_b200 import synth
    ocfg = oc.Config()
    offsets = _calibrate([0.15, 0.3, 0.5], ocfg)
    weights = _weights(offsets)
    idx = [0, 1, 2, 3, 7, 35, 36]                   # 450x400, 440x392, ..., and two frames of the same shape
    frames = [synth.synth_frame(*synth.fddb_shape(j), 100 + j) for j in idx]
    assert frames[0].shape == frames[5].shape == (400, 450, 3) and frames[1].shape == frames[6].shape
    app = pkg.InferenceCascadeApp("fddb", stage_models=[pkg.StageModel(w) for w in weights])
    infos = [pkg.ArrayImageInfo(f, "fddb_%d.png" % j) for j, f in zip(idx, frames)]
    got = _rows(app.run_inference_on_images(infos, merge=True))
    assert app._engine.plan_for(450, 400).n_windows == 6461
    exact = sum(_compare_frame(pkg, app._engine, f, weights, ocfg, g) for f, g in zip(frames, got))
    assert exact >= len(frames) - 1


def test_config4_1080p_dense_pyramid(pkg):
    from oracle import cascade as oc
    from rapidobjectdetectionusingcascadedcnns_b200 import synth
    cfd = dict(window_scale_factor=1.2, min_window_length=0.0125)
    for k, v in cfd.items():
        pkg.cf.set(k, v)
    ocfg = oc.Config(**cfd)
    offsets = _calibrate([0.03, 0.3, 0.5], ocfg, seed=5000)
    weights = _weights(offsets)
    frame = synth.synth_frame(1080, 1920, 5000)
    app = pkg.InferenceCascadeApp("hd", stage_models=[pkg.StageModel(w) for w in weights])
    got = _rows(app.run_inference_on_images([pkg.ArrayImageInfo(frame, "hd.png")], merge=True))[0]
    plan = app._engine.plan_for(1920, 1080)
    assert (plan.n_levels, plan.n_windows) == (21, 93645)
    _compare_frame(pkg, app._engine, frame, weights, ocfg, got)


def test_config5_4k_high_reject_properties(pkg):
    from oracle import cascade as oc, stage_net
    from rapidobjectdetectionusingcascadedcnns_b200 import native, synth
    from rapidobjectdetectionusingcascadedcnns_b200.engine import CascadeEngine
    cfd = dict(min_window_length=0.0125)
    pkg.cf.set("min_window_length", 0.0125)
    ocfg = oc.Config(**cfd)
    offsets = _calibrate([0.01, 0.5, 0.5], ocfg, seed=9000, shape=(480, 640))
    frames = np.stack([synth.synth_frame(2160, 3840, 9000 + j) for j in range(2)])
    # stage-0 offset for > 99.9 % rejection: calibrated on the GPU's own margins of the uncalibrated net
    pkg.cf.set("foreground_confidence_threshold", [2.0, 0.5, 0.5])           # nothing survives: stage 0 only
    lg0 = CascadeEngine(_weights([0.0, 0.0, 0.0])).run_frames(frames, keep_trace=True).trace["logits"][0].cpu().numpy()
    pkg.cf.set("foreground_confidence_threshold", 0.5)
    offsets[0] = -round(float(np.quantile(lg0[:, 1] - lg0[:, 0], 1.0 - 0.0005)) * 64.0) / 64.0
    weights = _weights(offsets)
    eng = CascadeEngine(weights)
    res = eng.run_frames(frames, keep_trace=True)
    n0 = 2 * 161969
    assert res.n_windows == n0 and res.stage_in[0] == n0
    assert res.stage_kept[0] < 0.001 * n0                          # > 99.9 % stage-1 rejection
    assert res.stage_in[1:] == res.stage_kept[:len(res.stage_in) - 1]
    # K3 at scale: survivors are exactly {w : p_fg > t}, in order
    p0 = res.trace["probs"][0].cpu().numpy()
    keep0 = res.trace["keep"][0].cpu().numpy()
    assert np.array_equal(keep0, np.nonzero(p0[:, 1].astype(np.float64) > 0.5)[0])
    # sampled oracle: all survivors + 1500 random rejected windows of frame 1, fp64 stage-0 logits
    plan = eng.plan_for(3840, 2160)
    d = torch.from_numpy(frames).cuda()
    wins = native.extract_windows(plan, native.pyramid(plan, d[1:2])).cpu().numpy()
    rs = np.random.RandomState(0)
    sample = np.unique(np.concatenate([keep0[keep0 >= 161969] - 161969, rs.randint(0, 161969, 1500)]))
    _, _, l64, _ = stage_net.stage_forward(wins[sample], weights[0], None, dtype=torch.float64)
    l64 = l64.numpy()
    lg = res.trace["logits"][0].cpu().numpy()[161969 + sample]
    assert np.abs(lg - l64).max() <= 1e-3 * np.abs(l64).max()
    margin = l64[:, 1] - l64[:, 0]
    flips = (margin > 0) != np.isin(161969 + sample, keep0)
    assert np.abs(margin[flips]).max(initial=0.0) < EPS[0]
    # the sampled raw windows equal the oracle's extraction (pyramid + grid at 4K)
    raw_ref, _ = oc.extract_windows(frames[1], 12, ocfg, "lib")
    assert np.array_equal(wins[sample], raw_ref[sample])
    # stages 1-2 and the final boxes against the fp64 oracle, continued from the stage-0 survivors (a few hundred
    # windows): bottleneck chain, cv2.resize re-crops, thresholds, groupRectangles
    wins0 = native.extract_windows(plan, native.pyramid(plan, d[0:1])).cpu().numpy()
    surv = keep0
    raw = np.stack([(wins0 if w < 161969 else wins)[w % 161969] for w in surv.tolist()])
    meta = res.trace["meta"][0].cpu().numpy()[surv]                    # (frame, *_norm box) of the survivors
    assert np.array_equal(meta[:, 0], surv // 161969)
    _, bott, _, _ = stage_net.stage_forward(raw, weights[0], None, dtype=torch.float64)
    same = True
    for i in (1, 2):
        size = 12 * 2 ** i
        raw = np.stack([oc.recrop(frames[m[0]], m[1:5], size, "lib") for m in meta]) if len(meta) else \
            np.zeros((0, size, size, 3), np.uint8)
        assert i < len(res.trace["X"]) or len(meta) == 0
        if len(meta) == 0:
            break
        assert np.array_equal(res.trace["X"][i].cpu().numpy(), raw), "stage %d input differs" % i
        p64, bott, l64, _ = stage_net.stage_forward(raw, weights[i], bott, dtype=torch.float64)
        l64 = l64.numpy()
        lg = res.trace["logits"][i].cpu().numpy()
        assert np.abs(lg - l64).max() <= 1e-3 * np.abs(l64).max()
        keep = np.nonzero(p64.numpy()[:, 1] > 0.5)[0]
        keep_gpu = res.trace["keep"][i].cpu().numpy()
        diff = np.setxor1d(keep, keep_gpu)
        if diff.size:
            assert np.abs(l64[diff, 1] - l64[diff, 0]).max() < EPS[i], "stage %d flip outside the band" % i
            same = False
            break
        meta, bott = meta[keep], bott[keep]
    if same:
        # final boxes: groupRectangles per frame over the ORACLE's candidates (= the GPU's, survivor sets match)
        boxes, conf, image = oc.postprocess(meta[:, 1:5].astype(np.int64), np.ones(len(meta)), meta[:, 0].astype(np.int64), ocfg, "restated")
        ref = np.column_stack([image, boxes, conf]).astype(np.float64).reshape(-1, 6)
        got = np.column_stack([res.frame, res.boxes, res.conf]).astype(np.float64).reshape(-1, 6)
        assert np.array_equal(got, ref)


def test_config1_single_net_640x480(pkg):
    from oracle import cascade as oc
    from rapidobjectdetectionusingcascadedcnns_b200 import synth
    w = synth.xavier_stage_weights(2, 48, fg_bias_offset=-10.0)
    frame = synth.synth_frame(480, 640, 0)
    app = pkg.InferenceApp("single", stage_models=[pkg.StageModel(w)])
    got = _rows(app.run_inference_on_images([pkg.ArrayImageInfo(frame, "s.png")], merge=True))[0]
    assert app._engine.plan_for(640, 480).n_windows == 3319
    r = oc.run_single([frame], w, oc.Config(), dtype=torch.float64, arith="lib")
    ref = np.column_stack([r.boxes, r.conf]).astype(np.float64).reshape(-1, 5)
    if not (got.shape == ref.shape and np.array_equal(got, ref)):
        res = app._engine.run_frames(frame[None], keep_trace=True)
        diff = np.setxor1d(res.trace["keep"][0].cpu().numpy(), r.stages[0].keep_idx)
        assert diff.size and np.abs(r.stages[0].logits[diff, 1] - r.stages[0].logits[diff, 0]).max() < EPS[2]


def test_fddb_caller_dense_pyramid(pkg, tmp_path):
    """the FDDB caller's regime (factor 1.005, merge=False, vertical enlarge) and writer."""
    from oracle import cascade as oc
    from rapidobjectdetectionusingcascadedcnns_b200 import synth
    from rapidobjectdetectionusingcascadedcnns_b200.app.evaluate_fddb_app import EvaluateFDDBApp
    ocfg = oc.Config(window_scale_factor=1.005, vertically_enlarge_bboxes=True)
    offsets = _calibrate([0.02, 0.3, 0.5], oc.Config())
    weights = _weights(offsets)
    pkg.cf.set("fddb_detection_output_dir", str(tmp_path))
    pkg.cf.set("fddb_img_base_dir", "/fddb/img")
    frames = [synth.synth_frame(*synth.fddb_shape(0), 100), synth.synth_frame(180, 200, 101)]
    infos = [pkg.ArrayImageInfo(frames[0], "/fddb/img/2002/08/11/big/img_591.jpg"),
             pkg.ArrayImageInfo(frames[1], "/fddb/img/2002/08/26/big/img_265.jpg")]
    app = pkg.InferenceCascadeApp("fddb", stage_models=[pkg.StageModel(w) for w in weights])
    ev = EvaluateFDDBApp("t", inference_app=app, folds=[infos[:1], infos[1:]])
    assert pkg.cf.get("window_scale_factor") == 1.005 and pkg.cf.get("inference_merge") is False
    plan = app._engine.plan_for(450, 400)
    assert (plan.n_levels, plan.n_windows) == (496, 116065)
    # the small frame against the fp64 oracle (vertically enlarged float boxes)
    got = _rows(ev.results[2])[0]
    r64 = oc.run_cascade([frames[1]], weights, ocfg, dtype=torch.float64, arith="lib")
    ref = np.column_stack([r64.boxes, r64.conf]).astype(np.float64).reshape(-1, 5)
    if got.shape == ref.shape:
        np.testing.assert_allclose(got, ref)
    else:
        _compare_frame(pkg, app._engine, frames[1], weights, oc.Config(window_scale_factor=1.005), got[:0])
    text = open(os.path.join(ev.export_dir, "fold-01-out.txt")).read().split("\n")
    assert text[0] == "2002/08/11/big/img_591" and int(text[1]) == len(ev.results[1][0])
    if ev.results[1][0]:
        b = ev.results[1][0][0]
        assert text[2] == "{} {} {} {} {}".format(b.xmin, b.ymin, b.width, b.height, b.confidence)


def test_config3_mixed_batch_properties_at_scale(pkg):
    """BASELINE config 3 shapes at a size the oracle cannot follow (140 frames of 35 shapes, ~850 k stage-0 windows,
    several launch chunks): size-independent properties.  Results do not depend on the input order, on the chunking
    or on the frame's neighbours in a batch; a second pass returns the same bytes; row output == box lists."""
    from rapidobjectdetectionusingcascadedcnns_b200 import synth
    weights = _weights([-40.0, -85.0, -24.0])
    app = pkg.InferenceCascadeApp("c3", stage_models=[pkg.StageModel(w) for w in weights])
    n = 140
    frames = [synth.synth_frame(*synth.fddb_shape(j), 100 + j) for j in range(n)]
    infos = [pkg.ArrayImageInfo(f, "c3_%d.png" % j, pin=(j % 3 == 0)) for j, f in enumerate(frames)]
    a = [np.asarray(r.rows() if hasattr(r, "rows") else np.zeros((0, 5))) for r in app.run_inference_on_images(infos, merge=True)]
    assert sum(x.shape[0] for x in a) > 50
    # 1. idempotence
    b = [np.asarray(r.rows() if hasattr(r, "rows") else np.zeros((0, 5))) for r in app.run_inference_on_images(infos, merge=True)]
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    # 2. order independence (a permutation regroups the shape groups and the chunks)
    perm = np.random.RandomState(0).permutation(n)
    c = app.run_inference_on_images([infos[i] for i in perm], merge=True)
    for k, i in enumerate(perm):
        got = np.asarray(c[k].rows() if hasattr(c[k], "rows") else np.zeros((0, 5)))
        assert np.array_equal(got, a[i]), i
    # 3. chunking independence: many small launch sets (the in-memory pipeline with its stage-0 callback) vs one
    pkg.cf.set("max_windows_per_launch", 60000)
    d = [np.asarray(r.rows() if hasattr(r, "rows") else np.zeros((0, 5))) for r in app.run_inference_on_images(infos, merge=True)]
    assert all(np.array_equal(x, y) for x, y in zip(a, d))
    rows = app.run_inference_rows(infos)
    flat = np.concatenate([np.column_stack([np.full((x.shape[0],), float(i)), x]) for i, x in enumerate(a) if x.shape[0]])
    assert np.array_equal(rows, flat)
    # 4. merge=False (one image per call, the FDDB caller's mode) gives the same boxes
    pkg.cf.set("max_windows_per_launch", 1 << 18)
    for i in (0, 17, 139):
        single = app.run_inference_on_images([infos[i]], merge=False)[0]
        got = np.asarray(single.rows() if hasattr(single, "rows") else np.zeros((0, 5)))
        assert np.array_equal(got, a[i])
