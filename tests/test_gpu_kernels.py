"""GPU parity tests, kernel by kernel, through the C ABI (librodc.so) against the CPU oracle.

Integer / byte / index work is bit-exact; floating point (stage net) carries its tolerance in
the test: logits within 1e-3 relative of the fp64 oracle (north_star), asserted much tighter.
"""
import hashlib

import numpy as np
import pytest
import torch

from conftest import case_inputs, load_golden

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def nat():
    from rapidobjectdetectionusingcascadedcnns_b200 import native
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    sm, major, minor = native.device_info()
    assert major >= 10, "built for sm_100a"
    return native


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------ K1 pyramid + windows

def test_scaled_plan_pyramid_bit_exact_incl_upscaling(nat):
    """rodc_plan_create_scales + rodc_pyramid = image.raw_scaled(next_scale) of the resize-then-crop branch
    (app/inference_cascade_app.py:374-394): Pillow BILINEAR to (int(W*s), int(H*s)), down- AND up-scaling,
    scale 1.0 = the frame itself; against the oracle's integer Pillow restatement and the installed Pillow."""
    from PIL import Image
    from oracle import cascade as oc
    from rapidobjectdetectionusingcascadedcnns_b200 import synth
    frames = np.stack([synth.synth_frame(61, 83, 40 + s) for s in range(3)])
    scales = [0.9576, 1.0, 0.5, 0.2393920493, 1.9152, 2.0, 3.3, 0.0625]
    plan = nat.Plan.from_scales(83, 61, scales)
    assert plan.n_levels == len(scales) and plan.n_windows == 0
    pyr = nat.pyramid(plan, dev(frames))
    for f in range(3):
        cache = oc.ScaledCache(frames[f], "restated")
        for l, sc in enumerate(scales):
            got = nat.pyramid_level(plan, pyr, f, l).cpu().numpy()
            ref = cache.raw_scaled(sc)
            assert got.shape == ref.shape == (61, 83, 3) if sc == 1.0 else got.shape == (int(61 * sc), int(83 * sc), 3)
            assert np.array_equal(got, ref), "frame %d scale %g differs from the restatement" % (f, sc)
            if sc != 1.0:
                lib = np.asarray(Image.fromarray(frames[f]).resize((int(83 * sc), int(61 * sc)), resample=2))
                assert np.array_equal(got, lib), "frame %d scale %g differs from Pillow" % (f, sc)


def test_window_scaled_rows_and_select_rows(nat):
    """GATHER rows of the resize-then-crop branch: int(xmin_norm * next_scale) truncation, frame stride, the
    Window.raw bounds flag, the per-window 'individually' escape; rodc_select_rows = fancy indexing."""
    rs = np.random.RandomState(5)
    entries = [(0.47878, 1000, 5000, 76, 57, 0, 0), (0.957584, 40000, 70000, 153, 114, 2, 0), (1.0, 0, 0, 0, 0, 0, 1)]
    slots = nat.scaled_slots(entries, torch.device("cuda"))
    n, size = 500, 24
    slot = rs.randint(0, 3, n).astype(np.int32)
    meta = np.zeros((n, 5), np.int32)
    meta[:, 0] = np.where(slot == 1, rs.randint(2, 5, n), rs.randint(0, 2, n))
    meta[:, 1] = rs.randint(0, 120, n); meta[:, 2] = rs.randint(0, 80, n)
    off, pitch, err = nat.window_scaled_rows(dev(meta), dev(slot), n, slots, 3, size, dense_rel_off=-12345)
    off, pitch = off.cpu().numpy(), pitch.cpu().numpy()
    bad = False
    for i in range(n):
        sc, base, stride, w, h, fb, ind = entries[slot[i]]
        if ind:
            assert off[i] == -12345 + i * size * size * 3 and pitch[i] == size * 3
            continue
        nx, ny = int(int(meta[i, 1]) * sc), int(int(meta[i, 2]) * sc)
        if nx + size > w or ny + size > h:
            bad = True
            continue
        assert off[i] == base + (meta[i, 0] - fb) * stride + (ny * w + nx) * 3, i
        assert pitch[i] == w * 3
    assert bad and int(err.item()) == 1
    ok = (slot == 2)
    _, _, err2 = nat.window_scaled_rows(dev(meta[ok]), dev(slot[ok]), int(ok.sum()), slots, 3, size)
    assert int(err2.item()) == 0
    _, _, err3 = nat.window_scaled_rows(dev(meta[:4]), dev(np.full(4, 7, np.int32)), 4, slots, 3, size)
    assert int(err3.item()) == 2
    src = rs.randint(-1000, 1000, (300, 5)).astype(np.int32)
    idx = rs.randint(0, 300, 1234).astype(np.int32)
    assert np.array_equal(nat.select_rows(dev(src), dev(idx), 1234).cpu().numpy(), src[idx])
    src64 = rs.randint(-2 ** 40, 2 ** 40, 300).astype(np.int64)
    assert np.array_equal(nat.select_rows(dev(src64), dev(idx), 1000).cpu().numpy(), src64[idx[:1000]])
    plan = nat.Plan(160, 120, 12, 1.1, 0.2)
    lv = nat.window_levels(plan, 3, 10, torch.device("cuda")).cpu().numpy()
    assert np.array_equal(lv, np.tile(plan.win_level, 3) + 10)


@pytest.mark.parametrize("name", ["cascade_small", "cascade_640x480", "cascade_factor12_enlarge"])
def test_pyramid_and_windows_bit_exact(nat, name):
    from oracle import cascade as oc, plan as oplan
    g, frames, weights, cfg = case_inputs(name)
    raws = []
    for fi, fr in enumerate(frames):
        plan = nat.Plan(fr.shape[1], fr.shape[0], 12, cfg.window_scale_factor, cfg.min_window_length)
        op = oplan.plan_pyramid(fr.shape[1], fr.shape[0], 12, cfg.window_scale_factor, cfg.min_window_length)
        assert plan.n_windows == op.n_windows and plan.n_levels == len(op.levels)
        d = dev(fr[None])
        pyr = nat.pyramid(plan, d)
        ref_levels = oc.pyramid_levels(fr, op, "restated")
        for l in range(plan.n_levels):
            got = nat.pyramid_level(plan, pyr, 0, l).cpu().numpy()
            assert np.array_equal(got, ref_levels[l]), "level %d of frame %d differs" % (l, fi)
        raws.append(nat.extract_windows(plan, pyr).cpu().numpy())
    raw = np.concatenate(raws)
    assert sha(raw) == str(g["windows_raw_sha"])          # the reference's own extract_windows output


def test_pyramid_batched_frames_and_identity_level(nat):
    from oracle import cascade as oc, plan as oplan
    from rapidobjectdetectionusingcascadedcnns_b200 import synth
    # min_window_length large enough that scale 1.0 itself is a level (identity copy path)
    frames = np.stack([synth.synth_frame(40, 56, s) for s in range(5)])
    plan = nat.Plan(56, 40, 12, 1.1, 0.2)
    assert plan.scale[0] == 1.0
    pyr = nat.pyramid(plan, dev(frames))
    op = oplan.plan_pyramid(56, 40, 12, 1.1, 0.2)
    for f in range(5):
        ref = oc.pyramid_levels(frames[f], op, "restated")
        for l in range(plan.n_levels):
            assert np.array_equal(nat.pyramid_level(plan, pyr, f, l).cpu().numpy(), ref[l])
    wins = nat.extract_windows(plan, pyr).cpu().numpy().reshape(5, plan.n_windows, 12, 12, 3)
    for f in range(5):
        ref_raw, _ = oc.extract_windows(frames[f], 12, oc.Config(min_window_length=0.2))
        assert np.array_equal(wins[f], ref_raw)
    meta = nat.window_meta(plan, 5, 7, "cuda").cpu().numpy().reshape(5, plan.n_windows, 5)
    assert (meta[:, :, 0] == np.arange(7, 12)[:, None]).all()
    assert np.array_equal(meta[2, :, 1:], op.norm)


# ------------------------------------------------------------------ K4 recrop
def test_recrop_matches_cv2_restatement_and_reference_run(nat):
    from oracle import cascade as oc
    g, frames, weights, cfg = case_inputs("cascade_small_reject")
    r = oc.run_cascade(frames, weights, cfg, dtype=torch.float32, arith="restated", keep_raw=True)
    # survivors of stage 0 -> inputs of stage 1 (24 px), survivors of stage 1 -> stage 2 (48 px)
    table = nat.FrameTable.from_list([dev(f) for f in frames])
    sizes = [12, 24, 48]
    # rebuild per-stage metadata from the oracle trace
    from oracle import plan as oplan
    norms, ids = [], []
    for fi, fr in enumerate(frames):
        p = oplan.plan_pyramid(fr.shape[1], fr.shape[0], 12, cfg.window_scale_factor, cfg.min_window_length)
        norms.append(p.norm); ids.append(np.full(p.n_windows, fi, np.int32))
    meta = np.column_stack([np.concatenate(ids), np.concatenate(norms)]).astype(np.int32)
    for i in range(2):
        meta = meta[r.stages[i].keep_idx]
        got = nat.recrop(table, dev(meta), sizes[i + 1]).cpu().numpy()
        assert np.array_equal(got, r.stages[i + 1].windows_raw)
        assert sha(got) == str(g["stage%d_X_sha" % (i + 1)])       # the reference's cv2.resize outputs


def test_recrop_special_paths(nat):
    from oracle import cv_resize
    rs = np.random.RandomState(3)
    frame = rs.randint(0, 256, (200, 220, 3)).astype(np.uint8)
    table = nat.FrameTable.from_list([dev(frame)])
    boxes = [(0, 0, 48, 48), (10, 20, 58, 68),        # 2x area path to 24, identity to 48
             (5, 5, 29, 29), (100, 50, 220, 200), (3, 7, 16, 19), (0, 0, 220, 200), (17, 3, 67, 52)]
    meta = np.asarray([[0, *b] for b in boxes], np.int32)
    for size in (12, 24, 48):
        got = nat.recrop(table, dev(meta), size).cpu().numpy()
        for k, (x0, y0, x1, y1) in enumerate(boxes):
            ref = cv_resize.cv_resize_linear(np.ascontiguousarray(frame[y0:y1, x0:x1]), size, size)
            assert np.array_equal(got[k], ref), (size, boxes[k])


# ------------------------------------------------------------------ K2 stage net
def _stage_case(stage_idx, n=200, seed=0):
    from rapidobjectdetectionusingcascadedcnns_b200 import synth
    sizes = synth.stage_input_sizes(3)
    S = sizes[stage_idx]
    w = synth.xavier_stage_weights(stage_idx, S, n_prev_bottleneck=512 * stage_idx)
    rs = np.random.RandomState(seed)
    # window-like content: smooth frame crops rather than white noise
    frame = synth.synth_frame(S * 6, S * 8, 40 + stage_idx)
    ys = rs.randint(0, frame.shape[0] - S, n); xs = rs.randint(0, frame.shape[1] - S, n)
    x = np.stack([frame[y:y + S, x_:x_ + S] for y, x_ in zip(ys, xs)])
    bott_in = (rs.rand(n, 512 * stage_idx).astype(np.float32) * 50.0) if stage_idx else None
    return S, w, x, bott_in


@pytest.mark.parametrize("n", [1, 37, 1187])       # single window; odd counts (half-filled window pairs at S = 24); > 1 CTA sweep
@pytest.mark.parametrize("stage_idx", [0, 1, 2])
def test_features_match_oracle(nat, stage_idx, n):
    from oracle import stage_net
    if n > 1000 and stage_idx == 2:
        n = 301                                      # keep the fp64 oracle within seconds at S = 48
    S, w, x, _ = _stage_case(stage_idx, n=n)
    st = nat.Stage(w)
    feat = nat.stage_features(st, nat.make_source(dense=dev(x)), 0, x.shape[0], "cuda").cpu().numpy()
    _, _, _, ref = stage_net.stage_forward(x, w, None if stage_idx == 0 else np.zeros((x.shape[0], 512 * stage_idx), np.float32),
                                           dtype=torch.float64)
    np.testing.assert_allclose(feat, ref.numpy(), rtol=2e-6, atol=3e-4)


@pytest.mark.parametrize("impl_name", ["simt", "tensor"])
@pytest.mark.parametrize("stage_idx", [0, 1, 2])
def test_stage_forward_matches_fp64_oracle(nat, stage_idx, impl_name):
    from oracle import stage_net
    impl = nat.IMPL_SIMT if impl_name == "simt" else nat.IMPL_TENSOR
    n = 300 if stage_idx < 2 else 200
    S, w, x, bott_in = _stage_case(stage_idx, n=n)
    st = nat.Stage(w)
    bott = torch.zeros((n, st.width), dtype=torch.float32, device="cuda")
    if bott_in is not None:
        bott[:, 512:] = dev(bott_in)
    ws = torch.empty((st.workspace_bytes(n, impl),), dtype=torch.uint8, device="cuda")
    logits, probs = nat.stage_forward(st, nat.make_source(dense=dev(x)), 0, n, bott, ws, impl)
    torch.cuda.synchronize()
    p64, b64, l64, _ = stage_net.stage_forward(x, w, bott_in, dtype=torch.float64)
    l64 = l64.numpy(); b64 = b64.numpy()
    got_l = logits.cpu().numpy().astype(np.float64)
    got_b = bott.cpu().numpy().astype(np.float64)
    # bottleneck (fc1 + ReLU) and logits: north_star tolerance 1e-3 relative; asserted at 2e-5
    scale_b = np.abs(b64[:, :512]).max()
    assert np.abs(got_b[:, :512] - b64[:, :512]).max() <= 2e-5 * scale_b
    err = np.abs(got_l - l64)
    assert err.max() <= 2e-5 * np.abs(l64).max(), (err.max(), np.abs(l64).max())
    # margin (the quantity the threshold acts on): inside north_star's epsilon band of 1e-4 at EVERY stage.
    # Measured maxima on B200 for this sample (stages 0 / 1 / 2): tensor 6.1e-5 / 9.4e-5 / 6.2e-5, plain fp32 FMA
    # (the arithmetic class of the reference's own Eigen GEMM) 4.1e-5 / 5.6e-5 / 7.4e-5 -- the floor is the fp32
    # rounding of the 512..1536 bottleneck values that fc2 sums, not the tensor path.
    m_err = np.abs((got_l[:, 1] - got_l[:, 0]) - (l64[:, 1] - l64[:, 0]))
    print("stage %d %s: max |logit err| %.3g, max |margin err| %.3g, max |fc1 err| %.3g" % (
        stage_idx, impl_name, err.max(), m_err.max(), np.abs(got_b[:, :512] - b64[:, :512]).max()))
    assert m_err.max() < 1e-4, m_err.max()
    np.testing.assert_allclose(probs.cpu().numpy(), p64.numpy(), rtol=1e-3, atol=1e-6)


def test_stage_forward_chunked_and_ragged(nat):
    """A workspace smaller than the problem forces several chunks; n not a multiple of 128."""
    from oracle import stage_net
    S, w, x, _ = _stage_case(0, n=333)
    st = nat.Stage(w)
    outs = []
    for impl in (nat.IMPL_SIMT, nat.IMPL_TENSOR):
        bott = torch.zeros((333, 512), dtype=torch.float32, device="cuda")
        ws = torch.empty((128 * st.K * 4,), dtype=torch.uint8, device="cuda")      # exactly one 128-window chunk
        logits, _ = nat.stage_forward(st, nat.make_source(dense=dev(x)), 0, 333, bott, ws, impl)
        outs.append(logits.cpu().numpy())
    _, _, l64, _ = stage_net.stage_forward(x, w, None, dtype=torch.float64)
    for o in outs:
        assert np.abs(o - l64.numpy()).max() <= 2e-5 * np.abs(l64.numpy()).max()


def test_stage_sources_agree(nat):
    """PYRAMID and RECROP sources feed the same bytes as the materialised DENSE windows."""
    from rapidobjectdetectionusingcascadedcnns_b200 import synth
    frames = np.stack([synth.synth_frame(96, 128, s) for s in (11, 12)])
    plan = nat.Plan(128, 96, 12, 1.1, 0.2)
    d = dev(frames)
    pyr = nat.pyramid(plan, d)
    wins = nat.extract_windows(plan, pyr)
    w0 = synth.xavier_stage_weights(0, 12)
    st0 = nat.Stage(w0)
    n = wins.shape[0]
    f_dense = nat.stage_features(st0, nat.make_source(dense=wins), 0, n, "cuda")
    f_pyr = nat.stage_features(st0, nat.make_source(plan=plan, pyr=pyr), 0, n, "cuda")
    assert torch.equal(f_dense, f_pyr)
    # first/offset handling
    f_tail = nat.stage_features(st0, nat.make_source(plan=plan, pyr=pyr), 100, n - 100, "cuda")
    assert torch.equal(f_dense[100:], f_tail)
    meta = nat.window_meta(plan, 2, 0, "cuda")
    table = nat.FrameTable.from_batch(d)
    w1 = synth.xavier_stage_weights(1, 24, n_prev_bottleneck=512)
    st1 = nat.Stage(w1)
    crops = nat.recrop(table, meta, 24)
    f_d1 = nat.stage_features(st1, nat.make_source(dense=crops), 0, n, "cuda")
    f_r1 = nat.stage_features(st1, nat.make_source(table=table, meta=meta), 0, n, "cuda")
    assert torch.equal(f_d1, f_r1)


# ------------------------------------------------------------------ K3 compaction
@pytest.mark.parametrize("n,rate", [(1, 1.0), (1, 0.0), (1023, 0.5), (1024, 0.01), (5061, 0.3), (100000, 0.999), (70001, 0.0)])
def test_compact_stable_and_exact(nat, n, rate):
    rs = np.random.RandomState(n)
    p1 = rs.rand(n).astype(np.float32)
    thr = float(np.float32(1.0 - rate)) if 0.0 < rate < 1.0 else (2.0 if rate == 0.0 else -1.0)
    p1[::7] = np.float32(thr) if 0.0 < rate < 1.0 else p1[::7]       # exact ties must be dropped (strict >)
    probs = np.stack([1 - p1, p1], 1).astype(np.float32)
    meta = rs.randint(0, 1000, (n, 5)).astype(np.int32)
    bott = rs.rand(n, 512).astype(np.float32)
    acc = rs.rand(n).astype(np.float32)
    keep_ref = np.nonzero(p1.astype(np.float64) > thr)[0]
    for mode in (0, 1, 2):
        meta_out = torch.zeros((n, 5), dtype=torch.int32, device="cuda")
        bott_out = torch.zeros((n, 1024), dtype=torch.float32, device="cuda")
        conf_out = torch.zeros((n,), dtype=torch.float32, device="cuda")
        keep, count = nat.compact(dev(probs), n, thr, meta_in=dev(meta), meta_out=meta_out, bott_in=dev(bott), width=512,
                                  bott_out=bott_out, col_off=512, conf_acc=dev(acc), conf_out=conf_out, conf_mode=mode)
        c = int(count.item())
        assert c == len(keep_ref)
        assert np.array_equal(keep.cpu().numpy()[:c], keep_ref)
        assert np.array_equal(meta_out.cpu().numpy()[:c], meta[keep_ref])
        assert np.array_equal(bott_out.cpu().numpy()[:c, 512:], bott[keep_ref])
        ref_conf = [p1[keep_ref], acc[keep_ref] + p1[keep_ref], acc[keep_ref] * p1[keep_ref]][mode]
        assert np.array_equal(conf_out.cpu().numpy()[:c], ref_conf.astype(np.float32))


# ------------------------------------------------------------------ K5 group rectangles
def _group_ref(cand, n_frames, thr):
    from oracle import group_rectangles as gr
    out = []
    for f in range(n_frames):
        c = cand[cand[:, 0] == f]
        rects = [(int(r[1]), int(r[2]), int(r[3] - r[1]), int(r[4] - r[2])) for r in c]
        rr, ww = gr.group_rectangles(rects, thr)
        for r, w in zip(rr, ww):
            out.append([f, r[0], r[1], r[2], r[3], w])
    return np.asarray(out, np.int32).reshape(-1, 6)


def test_group_rectangles_golden_vectors(nat):
    g = load_golden("third_party")
    for k in range(16):
        rects = g["gr_in_%d" % k]; thr = int(g["gr_thr_%d" % k])
        cand = np.column_stack([np.zeros(len(rects), np.int32), rects[:, 0], rects[:, 1],
                                rects[:, 0] + rects[:, 2], rects[:, 1] + rects[:, 3]]).astype(np.int32)
        out, count = nat.group_rectangles(dev(cand), len(cand), 1, thr)
        c = int(count.item())
        got = out.cpu().numpy()[:c]
        assert np.array_equal(got[:, 1:5], g["gr_out_%d" % k].reshape(-1, 4)), k
        assert np.array_equal(got[:, 5], g["gr_w_%d" % k]), k


@pytest.mark.parametrize("thr", [0, 1, 2, 3])
def test_group_rectangles_multi_frame_random(nat, thr):
    rs = np.random.RandomState(100 + thr)
    rows = []
    n_frames = 9
    for f in range(n_frames):
        if f in (2, 6):
            continue                                     # frames without candidates
        k = int(rs.randint(1, 400))
        base = rs.randint(0, 300, (int(rs.randint(1, 9)), 2))
        for _ in range(k):
            b = base[rs.randint(len(base))]
            s = int(rs.randint(12, 120))
            x, y = int(b[0] + rs.randint(-9, 10)), int(b[1] + rs.randint(-9, 10))
            rows.append([f, x, y, x + s + int(rs.randint(-4, 5)), y + s + int(rs.randint(-4, 5))])
    cand = np.asarray(rows, np.int32)
    out, count = nat.group_rectangles(dev(cand), len(cand), n_frames, thr)
    c = int(count.item())
    assert np.array_equal(out.cpu().numpy()[:c], _group_ref(cand, n_frames, thr))


def test_group_rectangles_empty(nat):
    out, count = nat.group_rectangles(torch.zeros((0, 5), dtype=torch.int32, device="cuda"), 0, 3, 1)
    assert int(count.item()) == 0


def test_group_rectangles_many_frames_many_tiles(nat):
    """12k candidates over 40 frames: the 2-D tile triangle spans several tiles per frame and frame borders
    fall inside tiles."""
    rs = np.random.RandomState(5)
    rows = []
    n_frames = 40
    for f in range(n_frames):
        k = int(rs.randint(200, 420))
        base = rs.randint(0, 400, (int(rs.randint(2, 12)), 2))
        for _ in range(k):
            b = base[rs.randint(len(base))]
            s = int(rs.randint(12, 100))
            x, y = int(b[0] + rs.randint(-9, 10)), int(b[1] + rs.randint(-9, 10))
            rows.append([f, x, y, x + s + int(rs.randint(-3, 4)), y + s + int(rs.randint(-3, 4))])
    cand = np.asarray(rows, np.int32)
    out, count = nat.group_rectangles(dev(cand), len(cand), n_frames, 1)
    c = int(count.item())
    assert np.array_equal(out.cpu().numpy()[:c], _group_ref(cand, n_frames, 1))


def test_fc1_kernel_variants_are_bit_identical(nat):
    """The three tensor fc1 kernels (cta_group::2 pair kernel = default for large m, single-CTA 128x256 tiles,
    single-CTA 128x128 tiles = default for small m; chosen by environment variables read once per process)
    see the same k order per output element: bit-identical logits, also at a ragged m spanning several tiles."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from rapidobjectdetectionusingcascadedcnns_b200 import native, synth
w = synth.xavier_stage_weights(1, 24, n_prev_bottleneck=512)
rs = np.random.RandomState(0)
x = torch.from_numpy(rs.randint(0, 256, (1100, 24, 24, 3)).astype(np.uint8)).cuda()
st = native.Stage(w)
bott = torch.zeros((1100, st.width), dtype=torch.float32, device="cuda")
ws = torch.empty((st.workspace_bytes(1100),), dtype=torch.uint8, device="cuda")
logits, _ = native.stage_forward(st, native.make_source(dense=x), 0, 1100, bott, ws)
np.save(sys.argv[1], logits.cpu().numpy())
'''
    from conftest import REPO
    outs = []
    variants = [dict(RODC_FC1_BN="256", RODC_FC1_2CTA="1"), dict(RODC_FC1_BN="256", RODC_FC1_2CTA="0"),
                dict(RODC_FC1_BN="128"), dict()]
    for k, var in enumerate(variants):
        path = os.path.join("/tmp", "rodc_fc1_%d.npy" % k)
        env = {kk: vv for kk, vv in os.environ.items() if not kk.startswith("RODC_FC1")}
        env.update(var)
        subprocess.run([sys.executable, "-c", code % REPO, path], check=True, env=env, timeout=300)
        outs.append(np.load(path))
    for o in outs[1:]:
        assert np.array_equal(outs[0], o)            # same MMA order per element -> bit-identical


@pytest.mark.parametrize("stage_idx", [0, 1, 2])
def test_generic_feature_kernel_writes_the_same_k_order(nat, stage_idx):
    """The split fc1 operand is stored in a permuted K order (stage.cuh kperm) that the prepared weights follow.  Both
    producers must agree on it: the generic feature kernel (RODC_FEATURES_GENERIC=1, shared-memory pooling, bias added
    after the MMA) gives the same logits as the streaming kernel up to the rounding of their different bias handling,
    and both agree with the fp64 oracle."""
    import os
    import subprocess
    import sys
    from oracle import stage_net
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
sys.path.insert(0, %r)
from rapidobjectdetectionusingcascadedcnns_b200 import native
from test_gpu_kernels import _stage_case
S, w, x, bott_in = _stage_case(%d, n=97)
st = native.Stage(w)
bott = torch.zeros((97, st.width), dtype=torch.float32, device="cuda")
if bott_in is not None:
    bott[:, 512:] = torch.from_numpy(bott_in).cuda()
ws = torch.empty((st.workspace_bytes(97),), dtype=torch.uint8, device="cuda")
logits, _ = native.stage_forward(st, native.make_source(dense=torch.from_numpy(x).cuda()), 0, 97, bott, ws)
np.save(sys.argv[1], logits.cpu().numpy())
'''
    from conftest import REPO
    outs = []
    for k, var in enumerate([dict(), dict(RODC_FEATURES_GENERIC="1")]):
        path = os.path.join("/tmp", "rodc_feat_%d_%d.npy" % (stage_idx, k))
        env = {kk: vv for kk, vv in os.environ.items() if not kk.startswith("RODC_FEATURES")}
        env.update(var)
        subprocess.run([sys.executable, "-c", code % (REPO, os.path.join(REPO, "tests"), stage_idx), path], check=True,
                       env=env, timeout=300)
        outs.append(np.load(path).astype(np.float64))
    S, w, x, bott_in = _stage_case(stage_idx, n=97)
    _, _, l64, _ = stage_net.stage_forward(x, w, bott_in, dtype=torch.float64)
    l64 = l64.numpy()
    scale = np.abs(l64).max()
    for o in outs:
        assert np.abs(o - l64).max() <= 2e-5 * scale
    assert np.abs(outs[0] - outs[1]).max() <= 4e-6 * scale


@pytest.mark.parametrize("spec", [
    dict(S=24, filters=(16, 32), ksize=3, conv_stride=1, pool_size=2, pool_stride=2, n_prev=512),
    dict(S=48, filters=(24,), ksize=5, conv_stride=2, pool_size=3, pool_stride=2, n_prev=0),
    dict(S=12, filters=(8, 16, 32), ksize=3, conv_stride=1, pool_size=3, pool_stride=1, n_prev=0),
])
def test_general_stage_architectures_match_fp64_oracle(nat, spec):
    """Several conv layers / strides (`network/net.py:189-221`, TF SAME padding) through the generic kernel."""
    from oracle import stage_net
    from rapidobjectdetectionusingcascadedcnns_b200 import synth
    S = spec["S"]
    w = synth.xavier_stage_weights(1, S, n_prev_bottleneck=spec["n_prev"], conv_filters=spec["filters"], ksize=spec["ksize"],
                                   conv_stride=spec["conv_stride"], pool_stride=spec["pool_stride"])
    rs = np.random.RandomState(1)
    for j in range(len(spec["filters"])):
        w["conv%d_b" % j] = (rs.randn(spec["filters"][j]) * 0.5).astype(np.float32)
    n = 150
    frame = synth.synth_frame(S * 5, S * 6, 77)
    ys = rs.randint(0, frame.shape[0] - S, n); xs = rs.randint(0, frame.shape[1] - S, n)
    x = np.stack([frame[y:y + S, x_:x_ + S] for y, x_ in zip(ys, xs)])
    prev = (rs.rand(n, spec["n_prev"]).astype(np.float32) * 20.0) if spec["n_prev"] else None
    geo = dict(in_size=S, conv_stride=spec["conv_stride"], pool_size=spec["pool_size"], pool_stride=spec["pool_stride"])
    st = nat.Stage(w, geometry=geo)
    assert st.K == w["fc1_W"].shape[0]
    for impl in (nat.IMPL_TENSOR, nat.IMPL_SIMT):
        bott = torch.zeros((n, st.width), dtype=torch.float32, device="cuda")
        if prev is not None:
            bott[:, 512:] = dev(prev)
        ws = torch.empty((st.workspace_bytes(n, impl),), dtype=torch.uint8, device="cuda")
        logits, probs = nat.stage_forward(st, nat.make_source(dense=dev(x)), 0, n, bott, ws, impl)
        p64, b64, l64, _ = stage_net.stage_forward(x, w, prev, dtype=torch.float64, conv_stride=spec["conv_stride"],
                                                   pool_size=spec["pool_size"], pool_stride=spec["pool_stride"])
        l64 = l64.numpy()
        assert np.abs(logits.cpu().numpy() - l64).max() <= 2e-5 * max(np.abs(l64).max(), 1.0)
        np.testing.assert_allclose(bott.cpu().numpy()[:, :512], b64.numpy()[:, :512], rtol=1e-4, atol=2e-5 * np.abs(b64.numpy()).max())
