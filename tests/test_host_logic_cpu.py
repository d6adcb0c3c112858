"""Host-side logic that needs no GPU: cf mirror, box semantics, stage-model files, app errors."""
import os

import numpy as np
import pytest


def test_cf_defaults_and_reset():
    from rapidobjectdetectionusingcascadedcnns_b200 import config as cf
    cf.reset()
    assert cf.get("foreground_confidence_threshold") == 0.5 and cf.get("window_scale_factor") == 1.1
    assert cf.get("min_window_length") == 0.075 and cf.get("nms") == cf.NMS_OPENCV
    assert cf.get("nms_opencv_min_neighbors") == 1 and cf.get("max_batch_size_original") == 1200
    assert cf.get("final_confidence_calculation") == cf.FINAL_CONFIDENCE_CALCULATION_LAST_STEP
    cf.set("min_window_length", 0.2)
    assert cf.get("min_window_length") == 0.2
    cf.reset()
    assert cf.get("min_window_length") == 0.075
    with pytest.raises(KeyError):
        cf.get("no_such_key")


def test_rectangle_intersects_like_reference_test():
    """The five pairs of the reference's `test/test_rectangles.py:9-39`."""
    from rapidobjectdetectionusingcascadedcnns_b200.data.rectangles import Rectangle
    pairs = [((100, 100, 200, 200), (150, 150, 175, 175)), ((150, 150, 175, 175), (100, 100, 200, 200)),
             ((100, 100, 200, 200), (150, 150, 250, 250)), ((100, 100, 200, 200), (150, 0, 175, 250)),
             ((100, 100, 200, 200), (50, 150, 250, 175))]
    for a, b in pairs:
        assert Rectangle(*a).intersects(Rectangle(*b)) and Rectangle(*b).intersects(Rectangle(*a))
    assert not Rectangle(0, 0, 10, 10).intersects(Rectangle(11, 0, 20, 10))
    assert Rectangle(0, 0, 9, 9).intersection_over_union(Rectangle(0, 0, 9, 9)) == 1.0


def test_window_norm_and_vertical_enlarge():
    from rapidobjectdetectionusingcascadedcnns_b200 import ArrayImageInfo, LabeledBoundingBox, Window
    from rapidobjectdetectionusingcascadedcnns_b200.data.db import label
    img = ArrayImageInfo(np.zeros((100, 120, 3), np.uint8), "a.png")
    w = Window(4, 8, 16, 20, img, 0.23939204936916333)
    assert (w.xmin_norm, w.ymin_norm, w.xmax_norm, w.ymax_norm) == (16, 33, 66, 83)
    assert w.raw_norm.shape == (50, 50, 3)
    with pytest.raises(ValueError):
        Window(0, 0, 12, 30, img, 0.25).raw_norm
    fg = label.get_by_iid(label.IID_FOREGROUND)
    assert fg.is_foreground and label.get_by_iid(1) is fg
    with pytest.raises(ValueError):
        label.get_by_iid(7)
    b = LabeledBoundingBox(10, 20, 40, 70, fg, 3.0, img)
    e = LabeledBoundingBox.vertically_enlarge_bboxes([b])[0]
    assert (e.xmin, e.ymin, e.xmax, e.ymax) == (10, 10.0, 40, 70) and e.confidence == 3.0
    assert LabeledBoundingBox.vertically_enlarge_bboxes([LabeledBoundingBox(0, 2, 5, 30, fg)])[0].ymin == 0


def test_stage_model_files_roundtrip(tmp_path):
    from rapidobjectdetectionusingcascadedcnns_b200 import StageModel, synth
    from rapidobjectdetectionusingcascadedcnns_b200.network.graph_pb import write_stage_pb
    from rapidobjectdetectionusingcascadedcnns_b200.network.stage_model import model_paths
    ws = synth.xavier_cascade()
    for i, w in enumerate(ws[:2]):
        m = StageModel(w)
        assert (m.in_size, m.n_prev, m.bottleneck_out_width) == ([12, 24][i], 512 * i, 512 * (i + 1))
        npz, pb = str(tmp_path / ("graph_k_%d.npz" % i)), str(tmp_path / ("graph_k_%d.pb" % i))
        m.save(npz)
        write_stage_pb(pb, w, stage_has_bottleneck_in=i > 0, in_size=m.in_size)
        for path in (npz, pb):
            r = StageModel.load(path)
            assert all(np.array_equal(r.weights[k], m.weights[k]) for k in m.weights)
    assert model_paths(str(tmp_path), "k", 1) == [str(tmp_path / "graph_k_1.pb"), str(tmp_path / "graph_k_1.npz")]
    with pytest.raises(ValueError):
        StageModel({k: v for k, v in ws[0].items() if k != "fc2_b"})


def test_pb_reader_on_tf_shaped_exports(tmp_path):
    """A GraphDef shaped like the reference's export (`app/train_app.py:177-227`): op nodes present, an
    ``import/`` scope, the ``<var>/read`` Identity nodes of an un-optimised freeze.  Weights come from the Const
    nodes, geometry from X / Conv2D / MaxPool -- including a non-default pooling that leaves the flattened size
    unchanged (5x5 / stride 1) and a strided net."""
    from rapidobjectdetectionusingcascadedcnns_b200 import StageModel, synth
    from rapidobjectdetectionusingcascadedcnns_b200.network.graph_pb import read_stage_graph, write_stage_pb
    from tensorboard.compat.proto import graph_pb2
    w1 = synth.xavier_stage_weights(1, 24, n_prev_bottleneck=512)
    pb = str(tmp_path / "graph_tf_1.pb")
    write_stage_pb(pb, w1, True, 24, prefix="import/", identity_reads=True)
    gd = graph_pb2.GraphDef(); gd.ParseFromString(open(pb, "rb").read())
    ops = {n.op for n in gd.node}
    assert {"Placeholder", "Const", "Identity", "Conv2D", "BiasAdd", "Relu", "MaxPool", "Reshape", "MatMul", "ConcatV2",
            "Softmax"} <= ops
    names = {n.name for n in gd.node}
    assert {"import/X", "import/bottleneck_in", "import/bottleneck_out_part1/W", "import/bottleneck_out_part1/W/read",
            "import/bottleneck_out/Relu", "import/final_softmax", "import/conv_layer_0/conv",
            "import/pool_layer_0/MaxPool"} <= names
    m = StageModel.load(pb)
    assert all(np.array_equal(m.weights[k], w1[k]) for k in m.weights)
    assert (m.in_size, m.conv_stride, m.pool_size, m.pool_stride, m.n_prev) == (24, 1, 3, 1, 512)
    # another pooling size with the same flattened size must NOT be read as the default 3x3
    pb5 = str(tmp_path / "graph_pool5_0.pb")
    w0 = synth.xavier_stage_weights(0, 12)
    write_stage_pb(pb5, w0, False, 12, pool_size=5)
    assert StageModel.load(pb5).pool_size == 5
    # strided net: 24 -> conv stride 2 -> 12 -> pool stride 2 -> 6; K = 6*6*32
    ws = dict(w0); rs = np.random.RandomState(3)
    ws["fc1_W"] = rs.uniform(-0.01, 0.01, (6 * 6 * 32, 512)).astype(np.float32)
    pbs = str(tmp_path / "graph_strided_0.pb")
    write_stage_pb(pbs, ws, False, 24, conv_stride=2, pool_stride=2)
    ms = StageModel.load(pbs)
    assert (ms.in_size, ms.conv_stride, ms.pool_stride, ms.weights["fc1_W"].shape[0]) == (24, 2, 2, 1152)
    # a Const-only file carries no geometry: refuse to guess
    bare = str(tmp_path / "graph_bare_0.pb")
    write_stage_pb(bare, w0, False, 12, ops=False)
    with pytest.raises(ValueError, match="refusing to guess"):
        read_stage_graph(bare)
    # geometry that contradicts the weights is an error, not a silent default
    bad = str(tmp_path / "graph_bad_0.pb")
    write_stage_pb(bad, w0, False, 16)
    with pytest.raises(ValueError):
        StageModel.load(bad)
    # StageModel.save(.pb) round trip incl. geometry
    out = str(tmp_path / "graph_rt_0.pb")
    ms.save(out)
    r = StageModel.load(out)
    assert (r.in_size, r.conv_stride, r.pool_stride) == (24, 2, 2) and np.array_equal(r.weights["fc1_W"], ws["fc1_W"])


def test_cascade_app_model_discovery_errors(tmp_path):
    import rapidobjectdetectionusingcascadedcnns_b200 as pkg
    pkg.cf.reset()
    pkg.cf.set("output_graph_dir", str(tmp_path))
    with pytest.raises(FileNotFoundError):
        pkg.InferenceCascadeApp("missing")
    ws = pkg.synth.xavier_cascade() if hasattr(pkg, "synth") else None
    from rapidobjectdetectionusingcascadedcnns_b200 import synth
    pkg.StageModel(synth.xavier_stage_weights(0, 12)).save(str(tmp_path / "graph_solo_0.npz"))
    with pytest.raises(FileNotFoundError):                       # a cascade needs at least two stage files
        pkg.InferenceCascadeApp("solo")
    pkg.StageModel(synth.xavier_stage_weights(1, 24, n_prev_bottleneck=512)).save(str(tmp_path / "graph_solo_1.npz"))
    app = pkg.InferenceCascadeApp("solo")
    assert app._n_nets == 2 and app.supported_img_width == 24 and not app.last_net
    assert app.model_full_path.endswith("graph_solo_0.pb")
    with pytest.raises(NotImplementedError):
        app.run_inference_on_raw_data(np.zeros((1, 12, 12, 3), np.uint8))
    single = pkg.InferenceApp("nope")
    with pytest.raises(FileNotFoundError):
        single._stage_files()
    pkg.cf.reset()


def test_box_list_behaves_like_the_reference_list():
    """`data/box_list.py`: the lazily materialised per-image result behaves like the reference's
    ``List[LabeledBoundingBox]`` (len, iteration, indexing, ==, +, list methods) and exposes the arrays."""
    from collections.abc import Sequence
    from rapidobjectdetectionusingcascadedcnns_b200.data.box_list import BoxList
    from rapidobjectdetectionusingcascadedcnns_b200.data.db import label
    from rapidobjectdetectionusingcascadedcnns_b200.data.rectangles import LabeledBoundingBox
    fg = label.get_by_iid(label.IID_FOREGROUND)
    boxes = np.asarray([[1, 2, 13, 14], [5, 6, 29, 30], [7, 8, 55, 56]], np.int32)
    img = object()
    post = BoxList(boxes, np.asarray([3.0, 2.0, 5.0]), img, fg, post_nms=True)
    assert isinstance(post, Sequence) and len(post) == 3 and bool(post) and post._list is None      # nothing built yet
    np.testing.assert_array_equal(post.rows(), [[1, 2, 13, 14, 3.0], [5, 6, 29, 30, 2.0], [7, 8, 55, 56, 5.0]])
    assert post._list is None                                       # rows() needs no object either
    b = post[1]
    assert isinstance(b, LabeledBoundingBox) and (b.xmin, b.ymin, b.xmax, b.ymax, b.width, b.height) == (5, 6, 29, 30, 24, 24)
    assert isinstance(b.xmin, np.integer) and isinstance(b.confidence, float) and b.confidence == 2.0   # `app/inference_app.py:203-212`
    assert b.image is img and b.label.is_foreground
    assert [x.xmax for x in post] == [13, 29, 55] and post[-1].confidence == 5.0 and len(post[0:2]) == 2
    assert list(post) == post._list and post == list(post) and not (post != list(post))
    assert (post + [7])[-1] == 7 and ([7] + post)[0] == 7
    post.sort(key=lambda x: x.confidence)                           # any other list method acts on the materialised list
    assert [x.confidence for x in post] == [2.0, 3.0, 5.0]
    post.append("x")
    assert len(post) == 4 and post.rows.__self__ is post
    pre = BoxList(boxes, np.asarray([0.75, 0.5, 1.0], np.float32), [img, None, img], fg, post_nms=False)
    first = pre[0]
    assert type(first.xmin) is int and isinstance(first.confidence, np.float32) and pre[1].image is None   # cascade `:319-334`
    assert BoxList(boxes[:0], np.zeros((0,)), img, fg, True) == [] and not BoxList(boxes[:0], np.zeros((0,)), img, fg, True)
