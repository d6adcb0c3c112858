"""Wall-clock stopwatch used for the three runtime figures the inference apps log
("window extraction", "inference only", "including window extraction").

Interface-compatible with the reference's `utils/time_watcher.py` (a named watcher that starts on
construction, ``stop()``, ``elapsed_seconds``, ``seconds_to_str``); built on ``perf_counter`` so that
short GPU-side phases are not quantised by the system clock.
"""
import datetime
import time

from . import log as _log


class TimeWatcher:
    __slots__ = ("label", "_t0", "_elapsed")

    def __init__(self, name: str):
        self.label = str(name)
        self._elapsed = None
        self._t0 = 0.0
        self.start()

    def start(self) -> None:
        _log.log("TimeWatcher Start: " + self.label)
        self._elapsed = None
        self._t0 = time.perf_counter()

    def stop(self) -> float:
        self._elapsed = time.perf_counter() - self._t0
        _log.log("TimeWatcher Stop %s: %s" % (self.label, self.seconds_to_str(self._elapsed)))
        return self._elapsed

    @property
    def elapsed_seconds(self):
        """Seconds between ``start`` and the last ``stop`` (``None`` while running)."""
        return self._elapsed

    @staticmethod
    def seconds_to_str(time_in_seconds) -> str:
        return str(datetime.timedelta(seconds=float(time_in_seconds)))
