"""ctypes binding of ``libptb_b200.so`` (C ABI in ``include/ptb.h``).

There is no CPU fallback: if the shared library has not been built the import of this module fails
loudly (``python -c "import __graft_entry__ as g; g.build()"`` or ``python -m pytestbeam_b200.build``).
"""
from __future__ import annotations

import ctypes as C
import os

PTB_MAX_PLANES = 32

PTB_DIGITISE = 1
PTB_TRACKS_FROM_TABLE = 2
PTB_WRITE_TRACKS = 4
PTB_KEEP_CHARGE64 = 8
PTB_RECYCLE_SLABS = 16
PTB_ZERO_RADIUS = 32
PTB_COMPACT = 64
PTB_COMPACT_NARROW = 128
PTB_COMPACT_PACKED = 256

PTB_DEV_SCRATCH_OVERFLOW = 1
PTB_DEV_SLAB_OVERFLOW = 2
PTB_DEV_BOX_TOO_LARGE = 4
PTB_DEV_COUNT_OVERFLOW = 8
PTB_DEV_PACK_OVERFLOW = 16

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PTB_LIB_PATH") or os.path.join(_HERE, "_lib", "libptb_b200.so")   # env: developer builds

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)


class PtbBeam(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("energy", "sigma_x", "sigma_y", "loc_x", "loc_y", "disp_x", "disp_y",
                                          "angle_x", "angle_y", "particle_rate")]


class PtbDevice(C.Structure):
    _fields_ = [("ncol", C.c_int32), ("nrow", C.c_int32), ("pitch_c", C.c_double), ("pitch_r", C.c_double),
                ("thickness", C.c_double), ("z", C.c_double), ("delta_x", C.c_double), ("delta_y", C.c_double),
                ("threshold_e", C.c_double), ("triggered", C.c_int32), ("_pad", C.c_int32),
                ("rad_length", C.c_double), ("Z", C.c_double), ("A", C.c_double), ("rho", C.c_double)]


class PtbOptions(C.Structure):
    _fields_ = [("pool_entries", C.c_int32), ("_pad", C.c_int32), ("trigger_accept", C.c_double)]


class PtbInjected(C.Structure):
    _fields_ = [("zstart", C.c_void_p), ("gap", C.c_void_p), ("zkick", C.c_void_p), ("eloss", C.c_void_p),
                ("accepted", C.c_void_p), ("dig_z", C.c_void_p), ("dig_off", C.c_void_p),
                ("eloss0_prev", C.c_double)]


class PtbTrackTable(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("energy", "angle_x", "angle_y", "x", "y", "z", "time_stamp", "eloss",
                                          "accepted")]


class PtbHitSlab(C.Structure):
    _fields_ = [("event", C.c_void_p), ("col", C.c_void_p), ("row", C.c_void_p), ("charge", C.c_void_p),
                ("charge64", C.c_void_p), ("capacity", C.c_uint64)]


class PtbCompactOut(C.Structure):
    _fields_ = [("hits", C.c_void_p), ("counts", C.c_void_p), ("accepted", C.c_void_p),
                ("plane_first", C.c_uint64 * (PTB_MAX_PLANES + 1)), ("charge", C.c_void_p), ("pos_lo", C.c_void_p),
                ("pos_hi", C.c_void_p), ("escapes", C.c_void_p), ("escape_capacity", C.c_uint64),
                ("word0", C.c_void_p), ("word1", C.c_void_p), ("word2", C.c_void_p)]


class PtbPackedLayout(C.Structure):
    _fields_ = [("exp_min", C.c_uint8 * PTB_MAX_PLANES), ("col_bits", C.c_uint8 * PTB_MAX_PLANES)]


class PtbBases(C.Structure):
    _fields_ = [("event", C.c_int64), ("time", C.c_int64), ("hits", C.c_uint64 * PTB_MAX_PLANES)]


class PtbTotals(C.Structure):
    _fields_ = [("hits", C.c_uint64 * PTB_MAX_PLANES), ("accepted", C.c_uint64), ("time_sum", C.c_int64),
                ("pairs", C.c_uint64), ("inside", C.c_uint64), ("pixels", C.c_uint64), ("error", C.c_uint32),
                ("escapes", C.c_uint32), ("reserved", C.c_uint64 * PTB_MAX_PLANES)]


class PtbRun(C.Structure):
    _fields_ = [("first", C.c_uint64), ("n", C.c_uint64), ("seed", C.c_uint64), ("flags", C.c_uint32),
                ("_pad", C.c_uint32), ("injected", C.POINTER(PtbInjected)), ("tracks", PtbTrackTable),
                ("slabs", PtbHitSlab * PTB_MAX_PLANES), ("bases_in", C.c_void_p), ("bases_out", C.c_void_p),
                ("totals", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
                ("scratch_rows", C.c_uint64 * PTB_MAX_PLANES), ("compact", PtbCompactOut)]


EXPORTS = ("ptb_create", "ptb_destroy", "ptb_last_error", "ptb_n_planes", "ptb_workspace_bytes",
           "ptb_prefix", "ptb_simulate", "ptb_pack_hits", "ptb_launch_count", "ptb_measure_fp64_peak",
           "ptb_timing_enable", "ptb_timing_read", "ptb_timing_read_stages", "ptb_correlate", "ptb_digest", "ptb_expand_hits",
           "ptb_expand_hits_narrow", "ptb_expand_hits_fd", "ptb_expand_hits_narrow_fd", "ptb_packed_layout",
           "ptb_expand_hits_packed", "ptb_expand_hits_packed_fd")

_lib = None


def default_scratch_rows(capacity: int) -> int:
    """PTB_DEFAULT_SCRATCH_ROWS of include/ptb.h."""
    return 4 * int(capacity) + 4096


def load() -> C.CDLL:
    """Load the CUDA library; raise if it is missing (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m pytestbeam_b200.build` "
                          "(nvcc, sm_100a); pytestbeam_b200 has no CPU path")
    L = C.CDLL(LIB_PATH)
    L.ptb_create.argtypes = [C.POINTER(PtbBeam), C.POINTER(PtbDevice), C.c_int32, C.POINTER(PtbOptions),
                             C.POINTER(C.c_void_p)]
    L.ptb_create.restype = C.c_int
    L.ptb_destroy.argtypes = [C.c_void_p]
    L.ptb_destroy.restype = None
    L.ptb_last_error.argtypes = [C.c_void_p]
    L.ptb_last_error.restype = C.c_char_p
    L.ptb_n_planes.argtypes = [C.c_void_p]
    L.ptb_n_planes.restype = C.c_int32
    L.ptb_workspace_bytes.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                      C.c_uint32]
    L.ptb_workspace_bytes.restype = C.c_size_t
    L.ptb_prefix.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(PtbInjected), C.c_void_p,
                             C.c_void_p]
    L.ptb_prefix.restype = C.c_int
    L.ptb_simulate.argtypes = [C.c_void_p, C.POINTER(PtbRun), C.c_void_p]
    L.ptb_simulate.restype = C.c_int
    L.ptb_pack_hits.argtypes = [C.POINTER(PtbHitSlab), C.c_uint64, C.c_void_p, C.c_void_p]
    L.ptb_pack_hits.restype = C.c_int
    L.ptb_timing_enable.argtypes = [C.c_void_p, C.c_int32]
    L.ptb_timing_enable.restype = C.c_int
    L.ptb_timing_read.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    L.ptb_timing_read.restype = C.c_int
    L.ptb_timing_read_stages.argtypes = [C.c_void_p, C.POINTER(C.c_double * 4), C.POINTER(C.c_uint64)]
    L.ptb_timing_read_stages.restype = C.c_int
    L.ptb_correlate.argtypes = [C.c_void_p] * 3 + [C.c_uint64] + [C.c_void_p] * 3 + [C.c_uint64, C.c_int64, C.c_int64,
                                C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    L.ptb_correlate.restype = C.c_int
    L.ptb_digest.argtypes = [C.POINTER(PtbHitSlab), C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
    L.ptb_digest.restype = C.c_int
    L.ptb_expand_hits.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64, C.c_uint64, C.c_void_p,
                                  C.c_int32]
    L.ptb_expand_hits.restype = C.c_int64
    L.ptb_expand_hits_narrow.argtypes = [C.c_void_p] * 5 + [C.c_uint64, C.c_int32, C.c_void_p, C.c_uint64, C.c_int64,
                                         C.c_uint64, C.c_void_p, C.c_int32]
    L.ptb_expand_hits_narrow.restype = C.c_int64
    L.ptb_expand_hits_fd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64, C.c_uint64, C.c_int32,
                                     C.c_uint64, C.c_int32]
    L.ptb_expand_hits_fd.restype = C.c_int64
    L.ptb_expand_hits_narrow_fd.argtypes = [C.c_void_p] * 5 + [C.c_uint64, C.c_int32, C.c_void_p, C.c_uint64, C.c_int64,
                                            C.c_uint64, C.c_int32, C.c_uint64, C.c_int32]
    L.ptb_expand_hits_narrow_fd.restype = C.c_int64
    L.ptb_packed_layout.argtypes = [C.c_void_p, C.POINTER(PtbPackedLayout)]
    L.ptb_packed_layout.restype = C.c_int
    L.ptb_expand_hits_packed.argtypes = [C.c_void_p] * 5 + [C.c_uint64, C.c_int32, C.POINTER(PtbPackedLayout), C.c_uint64,
                                         C.c_void_p, C.c_uint64, C.c_int64, C.c_uint64, C.c_void_p, C.c_int32]
    L.ptb_expand_hits_packed.restype = C.c_int64
    L.ptb_expand_hits_packed_fd.argtypes = [C.c_void_p] * 5 + [C.c_uint64, C.c_int32, C.POINTER(PtbPackedLayout),
                                            C.c_uint64, C.c_void_p, C.c_uint64, C.c_int64, C.c_uint64, C.c_int32,
                                            C.c_uint64, C.c_int32]
    L.ptb_expand_hits_packed_fd.restype = C.c_int64
    L.ptb_launch_count.argtypes = []
    L.ptb_launch_count.restype = C.c_uint64
    L.ptb_measure_fp64_peak.argtypes = [C.c_void_p]
    L.ptb_measure_fp64_peak.restype = C.c_double
    _lib = L
    return L


def pack_beam(beam: dict) -> PtbBeam:
    return PtbBeam(beam["energy"], beam["sigma_x"], beam["sigma_y"], beam["loc_x"], beam["loc_y"], beam["x_disp"],
                   beam["y_disp"], beam["x_angle"], beam["y_angle"], beam["particle_rate"])


def pack_devices(devices: list, materials: list):
    arr = (PtbDevice * len(devices))()
    for i, (d, m) in enumerate(zip(devices, materials)):
        arr[i] = PtbDevice(int(d["column"]), int(d["row"]), d["column_pitch"], d["row_pitch"], d["thickness"],
                           d["z_position"], d["delta_x"], d["delta_y"], d["threshold"],
                           int(d["trigger"] == "triggered"), 0, m["rad_length"], m["Z"], m["A"], m["rho"])
    return arr
