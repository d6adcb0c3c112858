"""Per-kernel rooflines of the secondary kernels (K1 pyramid, K2a features, K3 compaction, K5 group
rectangles), each timed in isolation with CUDA events (median of 20 after 3 warm-ups, L2 flushed between
runs) on the shapes of the bench workload (16 x 640x480 frames).  Algorithmic bytes as in SURVEY.md §8d /
DESIGN.md §4; peak = MEASURED_PEAKS.json hbm_gbs.

    python profiles/kernel_rooflines.py > profiles/<round>_kernel_rooflines.json
"""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from rapidobjectdetectionusingcascadedcnns_b200 import native, synth  # noqa: E402


def timeit(fn, flush, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    peak = 6545.9
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    dev = torch.device("cuda", 0)
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)
    B = 16
    frames = torch.from_numpy(np.stack([synth.synth_frame(480, 640, j) for j in range(B)])).to(dev)
    plan = native.Plan(640, 480, 12)
    out = []

    def row(name, ms, nbytes, note):
        gbs = nbytes / (ms * 1e-3) / 1e9
        out.append({"kernel": name, "ms": ms, "algorithmic_bytes": nbytes, "achieved_gbs": gbs, "peak_gbs": peak,
                    "frac": gbs / peak, "note": note})

    # K1
    pyr = torch.empty((B, plan.pyramid_bytes), dtype=torch.uint8, device=dev)
    scr = torch.empty((B, plan.scratch_bytes), dtype=torch.uint8, device=dev)
    ms = timeit(lambda: native.pyramid(plan, frames, pyr, scr), flush)
    alg = B * sum(3 * 640 * 480 + 3 * int(w) * int(h) for w, h in zip(plan.level_w, plan.level_h))
    row("K1 pyramid (h-pass + v-pass)", ms, alg, "SURVEY formula: source counted once per level; real DRAM traffic is ~4x lower")

    # K4 + K2a per stage on the product path: rodc_source_windows materialises the windows (pyramid crops at stage 0,
    # cv2.resize re-crops after), the streaming feature kernel reads them densely (stage 2 shape is the one that matters)
    ws = synth.xavier_cascade()
    from rapidobjectdetectionusingcascadedcnns_b200 import _lib
    import ctypes as C
    L = _lib.lib()
    table = native.FrameTable.from_batch(frames)
    for i, (S, n) in enumerate([(12, B * plan.n_windows), (24, 37888), (48, 18944)]):
        st = native.Stage(ws[i])
        meta = native.window_meta(plan, B, 0, dev)[:n].contiguous()
        src = native.make_source(plan=plan, pyr=pyr) if i == 0 else native.make_source(table=table, meta=meta)
        res = {}

        def copy_windows():
            res["w"] = native.source_windows(src, 0, n, S, dev)
        ms = timeit(copy_windows, flush, reps=10)
        row("K4 window copy / re-crop S=%d (%d windows)" % (S, n), ms, n * S * S * 3,
            "output bytes only; the re-crop is bound by the L1 sector rate of its scattered tap loads")
        dsrc = native.make_source(dense=res["w"])
        feat = torch.empty((n, st.K), dtype=torch.float32, device=dev)
        ms = timeit(lambda: native._lib.check(L.rodc_stage_features_f32(st.handle, C.byref(dsrc), 0, n, C.c_void_p(feat.data_ptr()),
                                                                         C.c_void_p(torch.cuda.current_stream().cuda_stream))), flush, reps=10)
        row("K2a features S=%d (%d windows)" % (S, n), ms, n * st.K * 4, "4 B per feature written (pure write stream: compare with the 6.3-7.0 TB/s write-only rate)")
        del st, feat, res, dsrc

    # K3 on stage-0 sized input with bottleneck gather
    n = B * plan.n_windows
    rs = np.random.RandomState(0)
    probs = torch.from_numpy(np.stack([1 - (p1 := rs.rand(n).astype(np.float32)), p1], 1)).to(dev)
    meta = native.window_meta(plan, B, 0, dev)
    bott = torch.rand((n, 512), device=dev)
    meta_o = torch.empty_like(meta); bott_o = torch.empty((n, 1024), device=dev); conf_o = torch.empty((n,), device=dev)
    keep = torch.empty((n,), dtype=torch.int32, device=dev); cnt = torch.zeros((1,), dtype=torch.int64, device=dev)
    ms = timeit(lambda: native.compact(probs, n, 0.13, meta_in=meta, meta_out=meta_o, bott_in=bott, width=512, bott_out=bott_o,
                                       col_off=512, conf_out=conf_o, keep_idx=keep, count=cnt), flush)
    s = int(cnt.item())
    row("K3 compaction (n=%d, kept=%d, 512-wide rows)" % (n, s), ms, 8 * n + s * (20 + 20 + 4 + 4) + s * 512 * 4 * 2,
        "scores read + ids/meta/conf written + bottleneck rows moved")

    # K5 on the bench's final candidates (~3450 per frame)
    rows = []
    for f in range(B):
        idx = np.sort(rs.choice(plan.n_windows, 3450, replace=False))
        rows.append(np.column_stack([np.full(3450, f), plan.norm[idx]]))
    cand = torch.from_numpy(np.concatenate(rows).astype(np.int32)).to(dev)
    k = cand.shape[0]
    res = {}

    def run_k5():
        res["o"] = native.group_rectangles(cand, k, B, 1)
    ms = timeit(run_k5, flush)
    kk = int(res["o"][1].item())
    row("K5 group rectangles (K=%d, K'=%d)" % (k, kk), ms, 20 * k + 24 * kk,
        "compute bound by the K^2/2 = %.0fM predicate evaluations, not by bytes" % (B * 3450 * 3450 / 2e6))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
