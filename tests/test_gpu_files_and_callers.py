"""GPU parity of the DEPLOYMENT interface and the callers around the path.

* cascades / single nets discovered on disk as ``graph_<key>[_<i>].pb|.npz`` (the reference's only way to get a
  model: `app/inference_cascade_app.py:36-65,89-117`, `app/inference_app.py:93-115`; export
  `app/train_app.py:177-227`) must give the golden boxes of the reference's own run,
* ``EvaluateRuntimeApp`` (`app/evaluate_runtime_app.py:50-97`),
* N-rank ``run_sharded`` over NCCL == the 1-rank result (skipped below 2 GPUs).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch

from conftest import REPO, case_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture()
def pkg():
    import rapidobjectdetectionusingcascadedcnns_b200 as p
    p.cf.reset()
    yield p
    p.cf.reset()


def _set_cf(pkg, meta_cf):
    pkg.cf.reset()
    for k, v in meta_cf.items():
        pkg.cf.set(k, v)


def _rows(results):
    return np.asarray([[i, b.xmin, b.ymin, b.xmax, b.ymax, b.confidence] for i, per in enumerate(results) for b in per],
                      np.float64).reshape(-1, 6)


def _write_cascade(pkg, weights, graph_dir, key, fmt):
    from rapidobjectdetectionusingcascadedcnns_b200.network.graph_pb import write_stage_pb
    for i, w in enumerate(weights):
        m = pkg.StageModel(w)
        if fmt == "npz":
            m.save(os.path.join(graph_dir, "graph_%s_%d.npz" % (key, i)))
        elif fmt == "pb":
            m.save(os.path.join(graph_dir, "graph_%s_%d.pb" % (key, i)))
        else:       # a TF-shaped export: scope prefix + <var>/read identities + op nodes
            write_stage_pb(os.path.join(graph_dir, "graph_%s_%d.pb" % (key, i)), w, i > 0, m.in_size,
                           prefix="import/", identity_reads=True)


@pytest.mark.parametrize("fmt", ["npz", "pb", "pb_tf_shaped"])
@pytest.mark.parametrize("name", ["cascade_small", "cascade_640x480"])
def test_cascade_loaded_from_graph_files_matches_reference_run(pkg, tmp_path, name, fmt):
    g, frames, weights, _ = case_inputs(name)
    _set_cf(pkg, g["meta"]["cf"])
    pkg.cf.set("output_graph_dir", str(tmp_path))
    _write_cascade(pkg, weights, str(tmp_path), "disk", fmt)
    app = pkg.InferenceCascadeApp("disk")                       # no stage_models=: discovery on disk
    assert app._models is None and app._n_nets == len(weights)
    assert app.model_full_path == os.path.join(str(tmp_path), "graph_disk_0.pb")
    infos = [pkg.ArrayImageInfo(f, "%s_%d.png" % (name, i)) for i, f in enumerate(frames)]
    got = _rows(app.run_inference_on_images(infos, merge=True))
    assert (app.supported_img_width, app.supported_img_height) == (48, 48)
    assert pkg.cf.get("img_width") == 12 and pkg.cf.get("cascade_increasing_input_dimensions") is True
    np.testing.assert_array_equal(got, g["final_boxes"])       # boxes AND cluster sizes of the reference's run
    loaded = app._load_models()
    for m, w in zip(loaded, weights):
        assert all(np.array_equal(m.weights[k], np.asarray(w[k], np.float32)) for k in m.weights)
    app.clean()
    assert app._engine is None


def test_single_net_loaded_from_graph_file(pkg, tmp_path):
    g, frames, weights, _ = case_inputs("single_small")
    _set_cf(pkg, g["meta"]["cf"])
    pkg.cf.set("output_graph_dir", str(tmp_path))
    pkg.StageModel(weights[0]).save(str(tmp_path / "graph_solo.pb"))
    app = pkg.InferenceApp("solo")
    assert app.model_full_path == str(tmp_path / "graph_solo.pb") and app.last_net
    infos = [pkg.ArrayImageInfo(f, "single_%d.png" % i) for i, f in enumerate(frames)]
    np.testing.assert_array_equal(_rows(app.run_inference_on_images(infos, merge=True)), g["final_boxes"])
    assert pkg.cf.get("img_width") == 48


class _CountingInfo:
    """ArrayImageInfo that counts ``clear_raw_img_cache`` calls."""

    def __init__(self, pkg, frame, key):
        self._i = pkg.ArrayImageInfo(frame, key)
        self.cleared = 0

    def __getattr__(self, name):
        return getattr(self._i, name)

    def clear_raw_img_cache(self):
        self.cleared += 1


def test_evaluate_runtime_app(pkg, tmp_path):
    """`app/evaluate_runtime_app.py:50-97`: cascade then single net over the same images, both built from their
    session keys, ``clean()`` + cache clearing after each run, runtime lines in the log."""
    from rapidobjectdetectionusingcascadedcnns_b200.app.evaluate_runtime_app import EvaluateRuntimeApp
    from rapidobjectdetectionusingcascadedcnns_b200.utils import log
    gc, frames, wc, _ = case_inputs("cascade_small")
    gs, frames_s, wsingle, _ = case_inputs("single_small")
    _set_cf(pkg, gc["meta"]["cf"])
    pkg.cf.set("output_graph_dir", str(tmp_path))
    _write_cascade(pkg, wc, str(tmp_path), "casc", "pb")
    pkg.StageModel(wsingle[0]).save(str(tmp_path / "graph_sing.npz"))
    infos = [_CountingInfo(pkg, f, "rt_%d.png" % i) for i, f in enumerate(frames)]
    log.log_clear()
    ev = EvaluateRuntimeApp("casc", "sing", 2, 2, img_infos=infos)
    assert set(ev.results) == {"cascade", "single"} and all(i.cleared == 2 for i in infos)
    assert pkg.cf.get("cache_dataset") is False
    np.testing.assert_array_equal(_rows(ev.results["cascade"]), gc["final_boxes"])
    # the single net over the same frames == a directly built single app
    direct = pkg.InferenceApp("sing").run_inference_on_images(
        [pkg.ArrayImageInfo(f, "rt_%d.png" % i) for i, f in enumerate(frames)], merge=True)
    np.testing.assert_array_equal(_rows(ev.results["single"]), _rows(direct))
    text = "\n".join(log.log_cache)
    assert "Cascade / GPU" in text and "Single Net / GPU" in text
    assert text.count("Runtime inference including window extraction: {} images".format(len(infos))) >= 2
    assert text.count("Runtime window extraction:") >= 2 and text.count("Runtime inference only:") >= 2
    assert ev.runtimes["cascade"] > 0 and ev.runtimes["single"] > 0
    with pytest.raises(FileNotFoundError):
        EvaluateRuntimeApp("missing_cascade", "sing", img_infos=infos)
    with pytest.raises(NotImplementedError):
        ev._run_inference(False, True)


def test_run_inference_on_windows_keeps_input_order_without_nms(pkg):
    """Interleaved windows of two images, NMS disabled: the reference returns the survivors in the order of the
    given windows (`app/inference_cascade_app.py:319-334`)."""
    from oracle import cascade as oc
    g, frames, weights, ocfg = case_inputs("cascade_small_reject")
    _set_cf(pkg, g["meta"]["cf"])
    pkg.cf.set("nms", pkg.cf.NMS_DISABLED)
    models = [pkg.StageModel(w) for w in weights]
    app = pkg.InferenceCascadeApp("w", stage_models=models)
    infos = [pkg.ArrayImageInfo(f, "ord_%d.png" % i) for i, f in enumerate(frames[:2])]
    per = []
    for info, fr in zip(infos, frames[:2]):
        raw, p = oc.extract_windows(fr, 12, ocfg)
        per.append([(pkg.Window(int(p.win_x[k]), int(p.win_y[k]), int(p.win_x[k]) + 12, int(p.win_y[k]) + 12, info,
                                p.levels[p.win_level[k]].scale), raw[k]) for k in range(p.n_windows)])
    n = min(len(per[0]), len(per[1]))
    inter = [x for pair in zip(per[0][:n], per[1][:n]) for x in pair]          # a0 b0 a1 b1 ...
    boxes = app.run_inference_on_windows([w for w, _ in inter], np.stack([r for _, r in inter]))
    key = lambda b: (b.image.full_key, b.xmin, b.ymin, b.xmax, b.ymax)
    # expected: the per-image survivors (image path, no NMS) merged back in interleaved window order
    pos = {}
    for j, (w, _) in enumerate(inter):
        pos.setdefault((w.image.full_key, w.xmin_norm, w.ymin_norm, w.xmax_norm, w.ymax_norm), []).append(j)
    order = [pos[key(b)][0] for b in boxes]
    assert len(boxes) > 4 and order == sorted(order)
    assert len({b.image.full_key for b in boxes}) == 2


# ------------------------------------------------------------------ N ranks == 1 rank (NCCL)
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import torch.distributed as dist
    import rapidobjectdetectionusingcascadedcnns_b200 as pkg
    from rapidobjectdetectionusingcascadedcnns_b200 import dist as rdist
    from conftest import case_inputs as ci
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        g, frames, weights, _ = ci("cascade_640x480")
        g2, frames2, _, _ = ci("cascade_small")
        pkg.cf.reset()
        for k, v in g["meta"]["cf"].items():
            pkg.cf.set(k, v)
        frames = list(frames) + list(frames2) + [frames[0]]
        app = pkg.InferenceCascadeApp("n", stage_models=[pkg.StageModel(w) for w in weights], device=rank)
        infos = [pkg.ArrayImageInfo(f, "n_%d.png" % i) for i, f in enumerate(frames)]
        sharded = rdist.run_sharded(app, infos, as_rows=True)
        local = app.run_inference_on_images(infos, merge=True)                  # the whole list on ONE rank
        ref = np.asarray([[i, b.xmin, b.ymin, b.xmax, b.ymax, b.confidence] for i, per in enumerate(local) for b in per],
                         np.float64).reshape(-1, 6)
        q.put((rank, bool(sharded.shape == ref.shape and np.array_equal(sharded, ref)), int(ref.shape[0])))
    except Exception as e:              # noqa: BLE001
        q.put((rank, False, repr(e)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_run_sharded_over_nccl_equals_single_rank():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, same, n in got:
        assert same is True, (rank, n)
        assert n > 0


def test_decode_upload_pipeline_from_files(pkg, tmp_path):
    """files are decoded ONCE by a thread pool and uploaded on a side stream while the previous
    chunk computes.  Results from files (PNG, lossless) == results from the same frames in memory, for one chunk
    and for many small chunks; an unreadable file yields [] without disturbing its neighbours
    (`app/abstract_inference_app.py:165-183`)."""
    from PIL import Image
    g, frames, weights, _ = case_inputs("cascade_small_reject")
    _set_cf(pkg, g["meta"]["cf"])
    frames = list(frames) + [frames[0][:, ::-1].copy(), frames[1][::-1].copy()]
    paths = []
    for i, f in enumerate(frames):
        p = str(tmp_path / ("img_%d.png" % i))
        Image.fromarray(f).save(p)
        paths.append(p)
    bad = str(tmp_path / "broken.png")
    with open(bad, "wb") as fh:
        fh.write(b"this is not a png")
    missing = str(tmp_path / "missing.png")
    app = pkg.InferenceCascadeApp("files", stage_models=[pkg.StageModel(w) for w in weights])
    ref = _rows(app.run_inference_on_images([pkg.ArrayImageInfo(f, "a_%d.png" % i) for i, f in enumerate(frames)], merge=True))
    assert ref.shape[0] > 0

    def file_infos():
        return [pkg.FileImageInfo(p) for p in paths]
    got = _rows(app.run_inference_on_images(file_infos(), merge=True))
    np.testing.assert_array_equal(got, ref)
    # header-only planning: sizes are known before any pixel is decoded
    fi = pkg.FileImageInfo(paths[0])
    assert (fi.img_height_original, fi.img_width_original) == frames[0].shape[:2] and fi._array is None
    # many small chunks -> the feeder thread runs ahead of the GPU; same boxes
    pkg.cf.set("max_windows_per_launch", 600)
    got2 = _rows(app.run_inference_on_images(file_infos(), merge=True))
    np.testing.assert_array_equal(got2, ref)
    # unreadable / missing files in the middle: [] for them, the rest unchanged
    infos = file_infos()
    infos.insert(2, pkg.FileImageInfo(bad))
    infos.insert(4, pkg.FileImageInfo(missing))
    res = app.run_inference_on_images(infos, merge=True)
    assert res[2] == [] and res[4] == []
    rest = [r for k, r in enumerate(res) if k not in (2, 4)]
    np.testing.assert_array_equal(_rows(rest), ref)
    pkg.cf.set("max_windows_per_launch", 1 << 18)
    with pytest.raises(ValueError):
        app.run_inference_on_images([pkg.FileImageInfo(bad), pkg.FileImageInfo(missing)], merge=True)


@pytest.mark.parametrize("enlarge", [False, True])
def test_row_output_equals_box_lists(pkg, enlarge):
    """``run_inference_rows`` (what ``dist.run_sharded`` exchanges) == the flattened per-image box lists, for frames of
    several shapes, with and without the vertical enlargement (`data/rectangles.py:185-220`)."""
    g, frames, weights, _ = case_inputs("cascade_small_reject")
    _set_cf(pkg, g["meta"]["cf"])
    pkg.cf.set("vertically_enlarge_bboxes", enlarge)
    app = pkg.InferenceCascadeApp("rows", stage_models=[pkg.StageModel(w) for w in weights])
    infos = [pkg.ArrayImageInfo(f, "r_%d.png" % i, pin=(i % 2 == 0)) for i, f in enumerate(frames)]
    infos.insert(1, pkg.ArrayImageInfo(np.zeros((8, 8, 3), np.uint8), "tiny.png"))      # yields no window: skipped
    lists = app.run_inference_on_images(infos, merge=True)
    rows = app.run_inference_rows(infos)
    np.testing.assert_array_equal(rows, _rows(lists))
    assert lists[1] == [] and not (rows[:, 0] == 1).any()
    if enlarge:
        assert rows.shape[0] and (rows[:, 2] != np.floor(rows[:, 2])).any()              # float ymin after the enlargement


def test_pinned_frames_are_uploaded_in_place(pkg):
    """Frames in page-locked memory (``ArrayImageInfo(pin=True)``) take the one-call ``rodc_upload_frames`` path: the
    device copy equals the frames (same shape and mixed shapes) and the results equal those of pageable frames."""
    from rapidobjectdetectionusingcascadedcnns_b200.engine import CascadeEngine
    g, frames, weights, _ = case_inputs("cascade_small_reject")
    _set_cf(pkg, g["meta"]["cf"])
    eng = CascadeEngine(weights)
    pinned = [pkg.ArrayImageInfo(f, "p_%d" % i, pin=True).raw_original(True) for i, f in enumerate(frames)]
    assert eng._all_pinned(pinned) and not eng._all_pinned(list(frames))
    same = [pinned[0], pkg.ArrayImageInfo(frames[0][::-1].copy(), "p_r", pin=True).raw_original(True)]
    dev = eng.upload_host_frames(same)
    assert np.array_equal(dev.cpu().numpy(), np.stack(same))
    mf = eng.upload_mixed(pinned)
    for k, i in enumerate(mf.order):
        h, w = mf.shapes[k]
        assert np.array_equal(mf.flat[mf.offs[k]:mf.offs[k] + h * w * 3].cpu().numpy().reshape(h, w, 3), frames[i])
    app = pkg.InferenceCascadeApp("pin", stage_models=[pkg.StageModel(w) for w in weights])
    a = _rows(app.run_inference_on_images([pkg.ArrayImageInfo(f, "x_%d" % i, pin=True) for i, f in enumerate(frames)], merge=True))
    b = _rows(app.run_inference_on_images([pkg.ArrayImageInfo(f, "x_%d" % i) for i, f in enumerate(frames)], merge=True))
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(a, g["final_boxes"])
