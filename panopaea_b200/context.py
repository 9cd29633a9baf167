"""Device context: one per process (one process per GPU)."""
from __future__ import annotations

import ctypes as C
import os

from . import _lib
from ._lib import check


class Context:
    def __init__(self, device: int | None = None, stream: int | None = None):
        L = _lib.load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = C.c_void_p()
        check(L.pano_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h)))
        self._h = h
        self.device = device
        self._L = L

    @property
    def handle(self):
        return self._h

    def sync(self):
        check(self._L.pano_ctx_sync(self._h))

    def num_sms(self) -> int:
        n = C.c_int()
        check(self._L.pano_ctx_num_sms(self._h, C.byref(n)))
        return n.value

    def launch_count(self) -> int:
        n = C.c_uint64()
        check(self._L.pano_ctx_launch_count(self._h, C.byref(n)))
        return n.value

    def timer_start(self):
        check(self._L.pano_timer_start(self._h))

    def timer_stop_ms(self) -> float:
        ms = C.c_double()
        check(self._L.pano_timer_stop_ms(self._h, C.byref(ms)))
        return ms.value

    def timer_mark(self):
        check(self._L.pano_timer_mark(self._h))

    def timer_marks_ms(self):
        """Elapsed ms between consecutive timer_mark() calls since the last read (synchronises)."""
        buf = (C.c_double * 4096)()
        n = C.c_int()
        check(self._L.pano_timer_marks_ms(self._h, buf, 4096, C.byref(n)))
        return [buf[i] for i in range(max(0, n.value - 1))]

    def set_option(self, key: str, value: int):
        check(self._L.pano_ctx_set_option(self._h, key.encode(), int(value)))

    def step_times(self):
        """(ms per phase accumulated [inflow, advect, neg_div, cg, project], steps) since the last call."""
        ms = (C.c_double * 5)()
        n = C.c_int64()
        check(self._L.pano_ctx_step_times(self._h, ms, C.byref(n)))
        return list(ms), n.value

    def stream(self) -> int:
        s = C.c_void_p()
        check(self._L.pano_ctx_stream(self._h, C.byref(s)))
        return s.value or 0

    def close(self):
        if self._h:
            self._L.pano_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default = None


def default_context() -> Context:
    global _default
    if _default is None:
        _default = Context()
    return _default


def device_count() -> int:
    n = C.c_int()
    check(_lib.load().pano_device_count(C.byref(n)))
    return n.value
