"""tic/toc wall-clock timer with the attribute surface of the reference's
nms_net/tools.py:11-33 (`total_time`, `calls`, `diff`, `average_time`), which
test.py uses around each forward call."""
import time


class Timer(object):

    def __init__(self):
        self.total_time, self.calls, self.diff = 0.0, 0, 0.0
        self._t0 = None

    @property
    def average_time(self):
        return self.total_time / self.calls if self.calls else 0.0

    def tic(self):
        self._t0 = time.perf_counter()

    def toc(self, average=True):
        if self._t0 is None:
            raise RuntimeError('toc() without tic()')
        self.diff = time.perf_counter() - self._t0
        self._t0 = None
        self.total_time += self.diff
        self.calls += 1
        return self.average_time if average else self.diff
