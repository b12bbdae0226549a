"""Per-step timeline for tuning the sharded loop: KB200_TIMELINE=1 makes `mark(name)` record a
CUDA event on the current (or given) stream plus the host clock; `report()` prints, for every
mark, the device time and the host time since the first mark of the step.  Off: no-ops."""
import os
import time

import torch

ON = os.environ.get("KB200_TIMELINE", "0") == "1"
_marks = []


def mark(name, stream=None):
    if not ON:
        return
    ev = torch.cuda.Event(enable_timing=True)
    ev.record(stream if stream is not None else torch.cuda.current_stream())
    _marks.append((name, time.perf_counter(), ev))


def report(prefix=""):
    """-> list of (name, device ms since first mark, host ms since first mark); prints on rank 0."""
    if not ON or not _marks:
        return []
    torch.cuda.synchronize()
    t0, e0 = _marks[0][1], _marks[0][2]
    rows = [(n, e0.elapsed_time(e), (t - t0)*1e3) for n, t, e in _marks]
    del _marks[:]
    if int(os.environ.get("RANK", "0")) == 0:
        print(prefix + "  ".join("%s d%.2f h%.2f" % r for r in rows), flush=True)
    return rows
