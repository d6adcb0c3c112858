"""FDDB evaluation caller (mirror of `app/evaluate_fddb_app.py:13-179`).

The accuracy-facing consumer of the cascade: it switches the path to its dense-pyramid regime
(`window_scale_factor = 1.005` -> 496 levels / 116 065 windows on a 450x400 frame), per-image mode
(`inference_merge = False`) and vertically enlarged boxes (`:34-41`), runs fold by fold and writes
``fold-XX-out.txt`` in FDDB's format ``<x y w h score>`` (`:95-137`).  The external Perl / gnuplot
scoring (`:50-60,173-179`) is out of scope; ``evaluation_hook`` lets a caller plug it in.
"""
import os
from typing import Callable, List, Optional

from .. import config as cf
from ..data.image_info import FileImageInfo
from ..utils import log
from .inference_cascade_app import InferenceCascadeApp


class EvaluateFDDBApp:
    def __init__(self, session_suffix="", inference_app: Optional[InferenceCascadeApp] = None,
                 folds: Optional[List[List]] = None, evaluation_hook: Optional[Callable[[str], None]] = None,
                 run_now=True):
        if not cf.get("foreground_equals_face"):
            raise AttributeError("This app does not make sense, if you're not looking for a face detector")
        cf.set("window_scale_factor", 1.005)            # `:34`
        cf.set("inference_merge", False)                # `:37`
        cf.set("vertically_enlarge_bboxes", True)       # `:41`
        self._session_suffix = session_suffix
        if self._session_suffix != "" and not self._session_suffix.startswith("_"):
            self._session_suffix = "_" + self._session_suffix
        self._export_dir = os.path.join(cf.get("fddb_detection_output_dir"), cf.get("session_key") + self._session_suffix)
        os.makedirs(self._export_dir, exist_ok=True)
        self._app = inference_app
        self._folds = folds                              # optional: explicit image-info lists (synthetic runs)
        self._hook = evaluation_hook
        self.results = {}
        if run_now:
            self._main()

    @property
    def export_dir(self) -> str:
        return self._export_dir

    def _get_img_infos_for_fddb_images(self, fold_nr: int) -> List:
        """`:66-93`: ``FDDB-fold-XX.txt`` lists image paths relative to the FDDB image dir, without '.jpg'."""
        if self._folds is not None:
            return self._folds[fold_nr - 1]
        path = os.path.join(cf.get("fddb_folds_dir"), "FDDB-fold-{:02d}.txt".format(fold_nr))
        with open(path) as f:
            rel = [x.strip() for x in f.readlines()]
        return [FileImageInfo(os.path.join(cf.get("fddb_img_base_dir"), r) + ".jpg", "fddb") for r in rel if r]

    def _persist_fold_results(self, fold_nr: int, img_infos: List, results_per_img) -> str:
        """`:95-137`: one block per image -- name, number of faces, ``left_x top_y width height score`` lines."""
        export_file_path = os.path.join(self._export_dir, "fold-{:02d}-out.txt".format(fold_nr))
        content = ""
        for img, bboxes in zip(img_infos, results_per_img):
            img_key = img.path_original.replace(cf.get("fddb_img_base_dir"), "").replace(".jpg", "")
            if img_key.startswith("/"):
                img_key = img_key[1:]
            content += "{}\n{}\n".format(img_key, len(bboxes))
            for b in bboxes:
                content += "{} {} {} {} {}\n".format(b.xmin, b.ymin, b.width, b.height, b.confidence)
        with open(export_file_path, "w") as f:
            f.write(content)
        return export_file_path

    def _main(self):
        log.log("Running inference on the FDDB dataset")
        app = self._app if self._app is not None else InferenceCascadeApp()
        n_folds = len(self._folds) if self._folds is not None else 10
        for fold_nr in range(1, n_folds + 1):
            log.log("*******************  Fold {}/{}  *************************".format(fold_nr, n_folds))
            infos = self._get_img_infos_for_fddb_images(fold_nr)
            res = app.run_inference_on_images(infos, merge=cf.get("inference_merge"))
            self.results[fold_nr] = res
            self._persist_fold_results(fold_nr, infos, res)
        if self._hook is not None:
            self._hook(self._export_dir)
