"""Cascade inference app (mirror of `app/inference_cascade_app.py:16-439`)."""
import os
from typing import List, Optional

from .. import config as cf
from ..network.stage_model import StageModel, model_paths
from ..utils import log
from .inference_app import InferenceApp


class InferenceCascadeApp(InferenceApp):
    _single = False

    def __init__(self, model_session_key=None, stage_models: Optional[List[StageModel]] = None, device=None):
        self._curr_net_index = None
        InferenceApp.__init__(self, model_session_key, stage_models, device)
        # count the stage files graph_<key>_0, _1, ... (`:36-50`); a cascade needs at least two
        if self._models is not None:
            self._n_nets = len(self._models)
        else:
            n = 0
            while any(os.path.exists(p) for p in model_paths(cf.get("output_graph_dir"), self._model_session_key, n)):
                n += 1
            self._n_nets = n
        if self._n_nets < 1:
            raise FileNotFoundError("Could not find any graph files belonging to the specified cascade: {}".format(
                self.model_full_path))
        elif self._n_nets == 1:
            raise FileNotFoundError("Found a single graph file, but a cascade must consist of at least two: {}".format(
                self.model_full_path))
        log.log("the loaded cascade contains {} nets".format(self._n_nets))

    @property
    def model_full_path(self):
        """`:89-117`: ``graph_<key>_<current net index>.pb``."""
        base = self._model_full_path
        if base.endswith(".pb"):
            base = base[:-3]
        if base.endswith("_0"):
            base = base[:-2]
        return "{}_{}{}".format(base, self._curr_net_index if self._curr_net_index is not None else 0, ".pb")

    def _stage_files(self) -> List[str]:
        files = []
        for i in range(self._n_nets):
            hit = [p for p in model_paths(cf.get("output_graph_dir"), self._model_session_key, i) if os.path.exists(p)]
            if not hit:
                raise FileNotFoundError("Could not find the required graph file: graph_{}_{}.pb".format(
                    self._model_session_key, i))
            files.append(hit[0])
        return files

    @property
    def last_net(self):
        return self._curr_net_index == (self._n_nets - 1)

    @property
    def supported_img_width(self):
        return self._load_models()[-1].supported_img_width

    @property
    def supported_img_height(self):
        return self._load_models()[-1].supported_img_height

    def _update_input_dims(self):
        """`:403-428`: max dims from the last net, window size from the first; flags increasing dims."""
        models = self._load_models()
        first, lastm = models[0], models[-1]
        cf.set("img_height_max", lastm.supported_img_height)
        cf.set("img_width_max", lastm.supported_img_width)
        cf.set("cascade_increasing_input_dimensions", first.supported_img_width < lastm.supported_img_width)
        cf.set("img_height", first.supported_img_height)
        cf.set("img_width", first.supported_img_width)

    def run_inference_on_raw_data(self, raw_data):
        raise NotImplementedError("The cascade does not support running raw data directly.")     # `:438-439`
