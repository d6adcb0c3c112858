"""Oracle (test infrastructure): run the reference's OWN Python under shims.

Only usable where ``/root/reference`` exists (the build container); never
imported by the GPU-side tests, ``smoke()`` or ``bench.py``.  It makes

    config.py, data/rectangles.py, data/image_info.py, data/db/label.py,
    app/abstract_inference_app.py, app/inference_app.py,
    app/inference_cascade_app.py

importable and executable UNMODIFIED by providing:

1. a ``config_local.py`` in a temp dir (``project_extension_root`` with the
   directories `config.py:30-49,592-597` insists on);
2. ``scipy.misc.imread`` / ``imresize`` restored on top of Pillow with SciPy
   1.0.1's logic (``imresize(arr, float)`` -> ``im.resize((int(W*r), int(H*r)),
   resample=BILINEAR)``), because SciPy >= 1.3 removed them;
3. a fake ``tensorflow`` package.  Everything the inference apps call on it is
   implemented (``Session``, ``graph.get_tensor_by_name``, ``gfile.FastGFile``,
   ``GraphDef``, ``import_graph_def``, ``reset_default_graph``); the rest of the
   import chain (training, augmentation, ...) gets inert mocks.
   ``Session.run`` evaluates the stage graph with ``oracle/stage_net.py`` in
   float32 -- the ONE piece of the reference path that is restated rather
   than executed (TensorFlow 1.4 cannot be installed here).

With this, the reference's own ``Window.extract_windows``, cascade loop,
``cv2.resize`` re-crop, ``_run_nms`` (``cv2.groupRectangles``) and box building
run as written and generate the golden vectors in ``tests/golden/``.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import io
import os
import sys
import tempfile
from typing import Dict, List
from unittest import mock

import numpy as np
import torch

REFERENCE_ROOT = "/root/reference"
_installed = {}


# ----------------------------------------------------------------------------- fake tensorflow
class _Dim:
    def __init__(self, v):
        self.value = v

    def __index__(self):
        return int(self.value)

    __int__ = __index__

    def __repr__(self):
        return "Dim(%r)" % (self.value,)


class _Tensor:
    def __init__(self, graph_name: str, kind: str, shape):
        self.graph_name, self.kind = graph_name, kind
        self.shape = tuple(_Dim(s) for s in shape)


class _GraphDef:
    def __init__(self):
        self.weights = None

    def ParseFromString(self, data: bytes):
        with np.load(io.BytesIO(data)) as z:
            self.weights = {k: z[k] for k in z.files}


class _Graph:
    def __init__(self):
        self.models: Dict[str, Dict[str, np.ndarray]] = {}

    def get_tensor_by_name(self, name: str) -> _Tensor:
        gname, rest = name.split("/", 1)
        w = self.models[gname]
        cout = w["conv0_W"].shape[3]
        j = 0
        while "conv%d_W" % (j + 1) in w:
            j += 1
            cout = w["conv%d_W" % j].shape[3]
        size = int(round((w["fc1_W"].shape[0] // cout) ** 0.5))
        fc1 = w["fc1_W"].shape[1]
        n_prev = w["fc2_W"].shape[0] - fc1
        if rest == "X:0":
            return _Tensor(gname, "X", (None, size, size, 3))
        if rest == "bottleneck_in:0":
            return _Tensor(gname, "bottleneck_in", (None, n_prev))
        if rest == "bottleneck_out/Relu:0":
            return _Tensor(gname, "bottleneck_out", (None, fc1 + n_prev))
        if rest == "final_softmax:0":
            return _Tensor(gname, "softmax", (None, 2))
        raise KeyError(name)


_GRAPH = _Graph()
TRACE: List[dict] = []          # one record per Session.run call


class _Session:
    def __init__(self, *a, **k):
        self.graph = _GRAPH

    def run(self, fetches, feed_dict):
        from . import stage_net
        single = not isinstance(fetches, (list, tuple))
        flist = [fetches] if single else list(fetches)
        gname = flist[0].graph_name
        x = bott = None
        for t, v in feed_dict.items():
            if t.kind == "X":
                x = np.asarray(v)
            elif t.kind == "bottleneck_in":
                bott = np.asarray(v, np.float32)
        assert x is not None and x.dtype == np.uint8, "the reference feeds raw uint8 windows"
        probs, bo, logits, _ = stage_net.stage_forward(x, _GRAPH.models[gname], bott, dtype=torch.float32)
        out = {"softmax": probs.numpy().astype(np.float32), "bottleneck_out": bo.numpy().astype(np.float32)}
        TRACE.append({"graph": gname, "X": x.copy(), "probs": out["softmax"].copy(),
                      "logits": logits.numpy().astype(np.float32)})
        res = [out[t.kind] for t in flist]
        return res[0] if single else res

    def close(self):
        pass


class _FastGFile:
    def __init__(self, path, mode="rb"):
        self._f = open(path, mode)

    def __enter__(self):
        return self._f

    def __exit__(self, *a):
        self._f.close()


def _import_graph_def(graph_def, name="", **kw):
    _GRAPH.models[name] = graph_def.weights


class _FakeTFFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    PREFIXES = ("tensorflow", "matplotlib")

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.PREFIXES:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__path__, m.__name__, m.__spec__, m.__loader__ = [], spec.name, spec, self
        if spec.name == "tensorflow":
            m.Session = _Session
            m.GraphDef = _GraphDef
            m.import_graph_def = _import_graph_def
            m.gfile.FastGFile = _FastGFile
            m.reset_default_graph = lambda: None
        return m

    def exec_module(self, module):
        pass


# ----------------------------------------------------------------------------- scipy.misc shim
def _imread(path, mode=None, flatten=False):
    from PIL import Image
    im = Image.open(path)
    if mode is not None and im.mode != mode:
        im = im.convert(mode)
    return np.array(im)


def _imresize(arr, size, interp="bilinear", mode=None):
    """SciPy 1.0.1 ``misc.imresize`` for uint8 RGB input (``toimage`` is the identity
    for uint8; a float ``size`` is a ratio; 'bilinear' = PIL resample 2)."""
    from PIL import Image
    im = Image.fromarray(np.asarray(arr, np.uint8))
    if isinstance(size, (float, np.floating)):
        size = tuple((np.array(im.size) * size).astype(int))
    elif isinstance(size, (int, np.integer)):
        size = tuple((np.array(im.size) * (size / 100.0)).astype(int))
    else:
        size = (size[1], size[0])
    func = {"nearest": 0, "lanczos": 1, "bilinear": 2, "bicubic": 3, "cubic": 3}
    return np.array(im.resize(tuple(int(s) for s in size), resample=func[interp]))


# ----------------------------------------------------------------------------- install
def install():
    """Idempotently install all shims; returns the temp ``project_extension_root``."""
    if _installed:
        return _installed["root"]
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError("the reference checkout is not present; golden vectors are generated in the "
                           "build container only")
    root = tempfile.mkdtemp(prefix="rodc_ref_")
    for d in ("output/input_augmented/ds0", "input/ds0/images/original/foreground", "output/graph"):
        os.makedirs(os.path.join(root, d), exist_ok=True)
    with open(os.path.join(root, "config_local.py"), "w") as f:
        f.write("_cf = dict(project_extension_root=%r, project_extension_root_fast=%r)\n" % (root, root))
    sys.path.insert(0, root)
    sys.meta_path.insert(0, _FakeTFFinder())
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import scipy.misc as misc
    misc.imread, misc.imresize = _imread, _imresize
    sys.path.insert(0, REFERENCE_ROOT)
    prev = os.environ.get("CUDA_VISIBLE_DEVICES")
    import config as cf                          # noqa: F401  (reference module; has import side effects)
    from utils import log
    log.disable_console_output()
    if prev is None:
        os.environ.pop("CUDA_VISIBLE_DEVICES", None)
    else:
        os.environ["CUDA_VISIBLE_DEVICES"] = prev
    _installed["root"] = root
    return root


def write_model(key: str, stage_weights, cascade: bool = True):
    """Serialise stage weights where the reference looks for ``graph_<key>[_<i>].pb``
    (`app/inference_app.py:38-40`, `app/inference_cascade_app.py:89-117`)."""
    import config as cf
    gdir = cf.get("output_graph_dir")
    ws = stage_weights if cascade else [stage_weights]
    for i, w in enumerate(ws):
        buf = io.BytesIO()
        np.savez(buf, **w)
        name = "graph_%s_%d.pb" % (key, i) if cascade else "graph_%s.pb" % key
        with open(os.path.join(gdir, name), "wb") as f:
            f.write(buf.getvalue())


def make_image_infos(frames, prefix="frame"):
    """Write frames as PNG and wrap them in the reference's ``ImageInfo``."""
    from PIL import Image
    from data.image_info import ImageInfo
    from data.db import label
    root = _installed["root"]
    d = os.path.join(root, "input/ds0/images/original/foreground")
    infos = []
    for j, fr in enumerate(frames):
        p = os.path.join(d, "%s_%04d.png" % (prefix, j))
        Image.fromarray(fr).save(p)
        infos.append(ImageInfo(p, label.get_by_key("foreground"), "ds0"))
    return infos
