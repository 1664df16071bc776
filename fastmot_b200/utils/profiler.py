"""Stage timers for `MOT.step`.

Keeps the observable contract of the reference's stage profiler (fastmot/utils/profiler.py:5-33, used by
fastmot/mot.py:138-178): `with Profiler('track'):` times a stage on the host clock, `aggregate=True` charges the time
to an already-counted call of that stage, and `Profiler.get_avg_millis(name)` is total / calls.  Storage is one
record per stage instead of two name-mangled counters, and a stage can additionally be bracketed by CUDA events
(`device=True`) so the same names report the time the GPU spent, resolved lazily in `get_avg_device_millis`.
"""
import time


class _Stage:
    __slots__ = ("calls", "seconds", "events", "device_ms")

    def __init__(self):
        self.calls = 0
        self.seconds = 0.0
        self.events = []
        self.device_ms = 0.0


class Profiler:
    _stages = {}

    def __init__(self, name, aggregate=False, device=False):
        self.name = name
        self._rec = Profiler._stages.setdefault(name, _Stage())
        self._device = device
        self._ev = None
        self.duration = 0.0
        if not aggregate:
            self._rec.calls += 1

    def __enter__(self):
        if self._device:
            import torch
            self._ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self._ev[0].record()
        self._t0 = time.perf_counter()
        return self

    def __exit__(self, *exc_info):
        self.duration = time.perf_counter() - self._t0
        self._rec.seconds += self.duration
        if self._ev is not None:
            self._ev[1].record()
            self._rec.events.append(self._ev)
        return False

    @classmethod
    def reset(cls):
        cls._stages = {}

    @classmethod
    def get_avg_millis(cls, name):
        rec = cls._stages.get(name)
        return 1e3 * rec.seconds / rec.calls if rec and rec.calls else 0.

    @classmethod
    def get_avg_device_millis(cls, name):
        """Mean CUDA-event time of the stage (only for stages entered with device=True); synchronises the events."""
        rec = cls._stages.get(name)
        if not rec or not rec.calls:
            return 0.
        for e0, e1 in rec.events:
            e1.synchronize()
            rec.device_ms += e0.elapsed_time(e1)
        rec.events = []
        return rec.device_ms / rec.calls
