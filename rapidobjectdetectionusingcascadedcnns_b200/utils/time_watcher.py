"""Wall-clock timers with the reference's interface (`utils/time_watcher.py:10-47`)."""
from datetime import timedelta
from time import time

from . import log


class TimeWatcher:
    def __init__(self, name: str):
        self._name = name
        self.start()

    def start(self):
        self._start_time = time()
        self._end_time = None
        self._elapsed_seconds = None
        log.log("TimeWatcher Start: {}".format(self._name))

    def stop(self):
        self._end_time = time()
        self._elapsed_seconds = self._end_time - self._start_time
        log.log("TimeWatcher Stop {}: {}".format(self._name, self.seconds_to_str(self._elapsed_seconds)))

    @property
    def elapsed_seconds(self):
        return self._elapsed_seconds

    @staticmethod
    def seconds_to_str(time_in_seconds) -> str:
        return str(timedelta(seconds=time_in_seconds))
