"""Drop-in for the computational half of the reference's correlation plot (``main/plotting.py:486-572, 730-764``):
the two co-occurrence histograms (column of device x vs column of device y, row vs row) per event.

``correlation_histograms`` does the numpy pre-processing exactly as ``plot_correlation`` does it (event window, event
shift, keep only events present in both tables, histogram shapes) and runs the event loop (``_eventloop_fast``) on the
GPU (``ptb_correlate``).  Plotting stays with the reference.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import h5min
from .. import native as N


def prepare(device_x, device_y, max_cols=(None, None), max_rows=(None, None), offset=0, event_size=None,
            event_numb_shift=0):
    """Pre-processing of ``main/plotting.py:501-556``; returns (device_x, device_y, max_cols, max_rows)."""
    device_x, device_y = np.copy(device_x), np.copy(device_y)
    if event_size is not None:
        device_x = device_x[(device_x["event_number"] <= offset + event_size) & (offset <= device_x["event_number"])]
        device_y = device_y[(device_y["event_number"] <= offset + event_size) & (offset <= device_y["event_number"])]
    device_y["event_number"] = device_y["event_number"] + event_numb_shift
    device_x = device_x[np.isin(device_x["event_number"], device_y["event_number"])]
    device_y = device_y[np.isin(device_y["event_number"], device_x["event_number"])]
    max_cols, max_rows = list(max_cols), list(max_rows)
    if max_cols[0] is None:
        if len(device_x) == 0 or len(device_y) == 0:
            raise ValueError("the two hit tables share no event")
        max_cols = [int(np.max(device_x["column"])) + 1, int(np.max(device_y["column"])) + 1]
        max_rows = [int(np.max(device_x["row"])) + 1, int(np.max(device_y["row"])) + 1]
    return device_x, device_y, max_cols, max_rows


def correlation_histograms(device_x, device_y, max_cols=(None, None), max_rows=(None, None), offset=0,
                           event_size=None, event_numb_shift=0, device=None):
    """(x_corr_hist, y_corr_hist) as ``plot_correlation`` has them before plotting (int32 ndarrays)."""
    if not torch.cuda.is_available():
        raise RuntimeError("pytestbeam_b200 needs a CUDA device; there is no CPU path")
    lib = N.load()
    dx, dy, max_cols, max_rows = prepare(device_x, device_y, max_cols, max_rows, offset, event_size, event_numb_shift)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    xh = torch.zeros(max_cols, dtype=torch.int32, device=dev)
    yh = torch.zeros(max_rows, dtype=torch.int32, device=dev)
    if len(dx) and len(dy):
        up = lambda a, t: torch.from_numpy(np.ascontiguousarray(a).astype(t)).to(dev)  # noqa: E731
        ev1, c1, r1 = up(dx["event_number"], np.int64), up(dx["column"], np.int16), up(dx["row"], np.int16)
        ev2, c2, r2 = up(dy["event_number"], np.int64), up(dy["column"], np.int16), up(dy["row"], np.int16)
        ev_min, ev_max = int(dx["event_number"].min()), int(dx["event_number"].max())
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            rc = lib.ptb_correlate(ev1.data_ptr(), c1.data_ptr(), r1.data_ptr(), len(dx), ev2.data_ptr(), c2.data_ptr(),
                                   r2.data_ptr(), len(dy), ev_min, ev_max, xh.data_ptr(), max_cols[0], max_cols[1],
                                   yh.data_ptr(), max_rows[0], max_rows[1], C.c_void_p(stream))
        if rc != 0:
            raise RuntimeError(f"ptb_correlate failed ({rc})")
    return xh.cpu().numpy(), yh.cpu().numpy()


def correlation_from_files(path_in_device_x: str, path_in_device_y: str, **kw):
    """Same, reading the two ``*_dut.h5`` files like ``plot_correlation`` does (``main/plotting.py:502-508``)."""
    return correlation_histograms(h5min.read_table(path_in_device_x, "Hits"), h5min.read_table(path_in_device_y, "Hits"),
                                  **kw)
