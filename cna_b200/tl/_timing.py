"""Optional wall-clock stage timer (``CNA_B200_TIMING=1``): prints where one association() call
spends its host time.  Off by default; costs one dict lookup per mark when off."""
import os
import threading
import time

ENABLED = bool(os.environ.get("CNA_B200_TIMING")) and os.environ.get("RANK", "0") == "0"
_marks = []
_lock = threading.Lock()


def mark(label):
    if ENABLED:
        with _lock:
            _marks.append((time.perf_counter(), threading.current_thread().name, label))


def report(reset=True):
    if not ENABLED or not _marks:
        return
    t0 = _marks[0][0]
    for t, th, label in _marks:
        print(f"[cna timing] {1e3 * (t - t0):8.2f} ms  {th:18s} {label}")
    if reset:
        _marks.clear()
