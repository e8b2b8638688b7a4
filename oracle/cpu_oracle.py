"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle.

Combines
  * ``oracle/_build/libptb_oracle.so`` (plain C, ``oracle/ptb_oracle.c``): the two sequential loops of
    the reference (``main/hit.py:209-241``, ``main/device.py:132-168``) and the MT19937 stream they
    consume, and
  * a numpy restatement of the interpreter-side set-up: the event-0 draws (``main/hit.py:70-73``), the
    accept/reject + bootstrap Landau table (``main/hit.py:84-102, 345-359, 385, 399-400, 428-431``) and
    the trigger mask (``main/device.py:56``), all drawn from ``numpy.random.RandomState`` in the
    reference's order.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module.
The recorded variate tables it returns are what the CUDA library's deterministic mode is fed with.

Parity pin: see the header of ``oracle/ptb_oracle.c``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libptb_oracle.so")
_lib = None


class OrcBeam(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "energy", "sigma_x", "sigma_y", "loc_x", "loc_y", "disp_x", "disp_y", "angle_x", "angle_y",
        "particle_distance")]


class OrcDevice(C.Structure):
    _fields_ = [("ncol", C.c_int32), ("nrow", C.c_int32),
                ("pitch_c", C.c_double), ("pitch_r", C.c_double), ("thickness", C.c_double), ("z", C.c_double),
                ("delta_x", C.c_double), ("delta_y", C.c_double), ("threshold", C.c_double),
                ("triggered", C.c_int32), ("_pad", C.c_int32), ("rad_length", C.c_double)]


class OrcDigi(C.Structure):
    _fields_ = [("n_hits", C.c_int64), ("cap_hits", C.c_int64),
                ("event", C.POINTER(C.c_int64)), ("col", C.POINTER(C.c_uint16)), ("row", C.POINTER(C.c_uint16)),
                ("charge", C.POINTER(C.c_float)), ("charge64", C.POINTER(C.c_double)),
                ("n_z", C.c_int64), ("cap_z", C.c_int64), ("z", C.POINTER(C.c_double)),
                ("z_off", C.POINTER(C.c_int64)),
                ("n_pairs", C.c_int64), ("n_inside", C.c_int64), ("n_pixels", C.c_int64)]


def build(force: bool = False) -> str:
    """Compile the C oracle with gcc (idempotent)."""
    src = os.path.join(_HERE, "ptb_oracle.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int64)
        L.orc_sizeof_mt.restype = C.c_size_t
        L.orc_mt_seed.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_mt_double.argtypes = [C.c_void_p]
        L.orc_mt_double.restype = C.c_double
        L.orc_mt_gauss.argtypes = [C.c_void_p]
        L.orc_mt_gauss.restype = C.c_double
        L.orc_mt_poisson.argtypes = [C.c_void_p, C.c_double]
        L.orc_mt_poisson.restype = C.c_int64
        L.orc_record_tracks.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_double, dp, ip, dp]
        L.orc_record_tracks.restype = None
        L.orc_highland.argtypes = [C.c_double] * 3
        L.orc_highland.restype = C.c_double
        L.orc_cloud_sigma.argtypes = [C.c_double]
        L.orc_cloud_sigma.restype = C.c_double
        L.orc_noise_sigma.argtypes = [C.c_double] * 3
        L.orc_noise_sigma.restype = C.c_double
        L.orc_pixel_index.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int32]
        L.orc_pixel_index.restype = C.c_long
        L.orc_pixel_centre.argtypes = [C.c_long, C.c_double, C.c_double, C.c_int32]
        L.orc_pixel_centre.restype = C.c_double
        L.orc_tracks.argtypes = [C.POINTER(OrcBeam), C.POINTER(OrcDevice), C.c_int32, C.c_int64, dp, ip, dp, dp,
                                 dp, dp, dp, dp, dp, dp, ip]
        L.orc_tracks.restype = C.c_int
        L.orc_digitise_plane.argtypes = [C.POINTER(OrcDevice), C.c_int32, C.c_int32, C.c_int64, dp, dp, dp,
                                         C.POINTER(C.c_uint8), C.c_void_p, dp]
        L.orc_digitise_plane.restype = C.POINTER(OrcDigi)
        L.orc_digitise_plane_ex.argtypes = L.orc_digitise_plane.argtypes + [C.c_int32]
        L.orc_digitise_plane_ex.restype = C.POINTER(OrcDigi)
        u16p = C.POINTER(C.c_uint16)
        L.orc_eventloop.argtypes = [ip, u16p, u16p, C.c_int64, ip, u16p, u16p, C.c_int64, C.POINTER(C.c_int32),
                                    C.c_int64, C.POINTER(C.c_int32), C.c_int64]
        L.orc_eventloop.restype = None
        L.orc_digi_free.argtypes = [C.POINTER(OrcDigi)]
        L.orc_digi_free.restype = None
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


class MT:
    """Numba's / RandomState's MT19937 stream (seeded like ``np.random.seed(int)``)."""

    def __init__(self, seed: int):
        self._buf = C.create_string_buffer(lib().orc_sizeof_mt())
        lib().orc_mt_seed(self._buf, int(seed) & 0xFFFFFFFF)

    @property
    def ptr(self):
        return C.cast(self._buf, C.c_void_p)

    def double(self):
        return lib().orc_mt_double(self.ptr)

    def gauss(self):
        return lib().orc_mt_gauss(self.ptr)

    def poisson(self, lam):
        return lib().orc_mt_poisson(self.ptr, float(lam))


# ------------------------------------------------------------------------------------------------
# flattening, as main/hit.py:24-61 and main/device.py:30-53 do it
# ------------------------------------------------------------------------------------------------

def particle_distance(beam: dict) -> float:
    return 1 / beam["particle_rate"] * 10**9


def pack_beam(beam: dict) -> OrcBeam:
    return OrcBeam(beam["energy"], beam["sigma_x"], beam["sigma_y"], beam["loc_x"], beam["loc_y"], beam["x_disp"],
                   beam["y_disp"], beam["x_angle"], beam["y_angle"], particle_distance(beam))


def pack_devices(devices: list, materials: list):
    arr = (OrcDevice * len(devices))()
    for i, (d, m) in enumerate(zip(devices, materials)):
        arr[i] = OrcDevice(d["column"], d["row"], d["column_pitch"], d["row_pitch"], d["thickness"], d["z_position"],
                           d["delta_x"], d["delta_y"], d["threshold"], int(d["trigger"] == "triggered"), 0,
                           m["rad_length"])
    return arr


# ------------------------------------------------------------------------------------------------
# Landau table (interpreter side of the reference)
# ------------------------------------------------------------------------------------------------

def moyal_pdf(delta, x, z, Z, A, energy, rho, a):
    """Moyal-form Landau density used by the reference (``main/hit.py:385, 399-400, 428-431``)."""
    beta = np.sqrt(1 - 1 / (1 + (energy / 0.511) ** 2))
    xi = 0.1535 * z**2 * Z * rho * x / (A * beta**2)
    lam = 1 / xi * (delta - 0.025) - beta**2 - np.log(xi / 0.000001) - 1 + np.e
    return a * (1 / np.sqrt(2 * np.pi) * np.exp(-(lam + np.exp(-lam)) / 2))


def landau_args(dev_thickness, Z, A, rho, energy):
    return (dev_thickness * 10 ** (-4), -1, Z, A, energy, rho, 1)


def accept_pool(rs: np.random.RandomState, args, n: int, xmin=0, xmax=0.5):
    """One accept/reject pass over n candidates (``main/hit.py:347-356``): the accepted x values."""
    grid = np.linspace(xmin, xmax, n)
    top = np.max(moyal_pdf(grid, *args))
    bottom = np.min(moyal_pdf(grid, *args))
    cand = rs.uniform(xmin, xmax, n)
    level = rs.uniform(bottom, top, n)
    return cand[level < moyal_pdf(cand, *args)]


def landau_table(rs: np.random.RandomState, beam: dict, devices: list, materials: list, n: int) -> np.ndarray:
    """energy_lost[D, N] before the track loop zeroes the misses (``main/hit.py:84-102``)."""
    D = len(devices)
    table = np.zeros((D, n))
    for d in range(D):
        args = landau_args(devices[d]["thickness"], materials[d]["Z"], materials[d]["A"], materials[d]["rho"],
                           beam["energy"] - np.mean(table[d - 1]))
        pool = accept_pool(rs, args, n)
        table[d] = rs.choice(pool, size=n)
    return table


# ------------------------------------------------------------------------------------------------
# full oracle run
# ------------------------------------------------------------------------------------------------

@dataclass
class Variates:
    """Everything random about one run, in the layout the deterministic CUDA mode consumes."""
    zstart: np.ndarray        # f8 [N,4]   (z_ax, z_ay, z_x, z_y)
    gap: np.ndarray           # i8 [N]     Poisson gaps in ns
    zkick: np.ndarray         # f8 [N,D-1,2]
    eloss: np.ndarray         # f8 [D,N]   Landau table BEFORE zeroing
    accepted: np.ndarray      # u1 [N]
    dig_z: list = field(default_factory=list)     # per plane: f8 [n_z]
    dig_off: list = field(default_factory=list)   # per plane: i8 [N+1]


@dataclass
class OracleRun:
    tracks: tuple             # (energy, ax, ay, x, y, z, t) flat [N*D], particle-major
    eloss: np.ndarray         # f8 [D,N] AFTER zeroing
    hits: list                # per plane: structured array (event_number,frame,column,row,charge)
    charge64: list            # per plane: f8 charges before the f4 store
    variates: Variates
    census: dict


HITS_DTYPE = np.dtype([("event_number", "<i8"), ("frame", "<u2"), ("column", "<u2"), ("row", "<u2"),
                       ("charge", "<f4")])


def draw_interpreter_side(beam, devices, materials, n, seed_interp, table_override=None):
    """Consume RandomState(seed_interp) in the reference's order; return (zx0, zy0, t0, table, accepted).

    With ``table_override`` the Landau table is not drawn (tiny N, where the reference itself raises)."""
    rs = np.random.RandomState(seed_interp)
    zx0 = rs.standard_normal()      # np.random.normal(loc, s) == loc + s * this   (main/hit.py:70)
    zy0 = rs.standard_normal()      # main/hit.py:71
    t0 = rs.poisson(particle_distance(beam))                     # main/hit.py:73
    if table_override is not None:
        table = np.ascontiguousarray(table_override, dtype=np.float64)
    else:
        table = landau_table(rs, beam, devices, materials, n)    # main/hit.py:84-102
    accepted = rs.uniform(0, 1, n) < 0.7                         # main/device.py:56
    return zx0, zy0, int(t0), table, accepted.astype(np.uint8)


def run_tracks(beam, devices, materials, n, zstart, gap, zkick, eloss_pre):
    L = lib()
    D = len(devices)
    ob, od = pack_beam(beam), pack_devices(devices, materials)
    eloss = np.ascontiguousarray(eloss_pre, dtype=np.float64).copy()
    out = [np.empty(n * D, dtype=np.float64) for _ in range(6)]
    t = np.empty(n * D, dtype=np.int64)
    rc = L.orc_tracks(C.byref(ob), od, D, n, _dp(zstart), _ip(gap), _dp(zkick), _dp(eloss), *[_dp(a) for a in out],
                      _ip(t))
    if rc != 0:
        raise ValueError("oracle: plane 0 must sit at z=0 and consecutive planes must differ in z (SURVEY Q2)")
    return tuple(out) + (t,), eloss


def run_digitise_plane(devices, materials, dut, n, x, y, eloss, accepted, mt: MT | None = None, inj=None,
                       zero_radius=False):
    L = lib()
    od = pack_devices(devices, materials)
    acc = np.ascontiguousarray(accepted, dtype=np.uint8)
    injp = _dp(np.ascontiguousarray(inj, dtype=np.float64)) if inj is not None else None
    r = L.orc_digitise_plane_ex(od, len(devices), dut, n, _dp(x), _dp(y), _dp(eloss),
                                acc.ctypes.data_as(C.POINTER(C.c_uint8)), mt.ptr if mt is not None else None, injp,
                                int(bool(zero_radius)))
    rr = r.contents
    nh, nz = rr.n_hits, rr.n_z
    hits = np.zeros(nh, dtype=HITS_DTYPE)
    if nh:
        hits["event_number"] = np.ctypeslib.as_array(rr.event, (nh,))
        hits["column"] = np.ctypeslib.as_array(rr.col, (nh,))
        hits["row"] = np.ctypeslib.as_array(rr.row, (nh,))
        hits["charge"] = np.ctypeslib.as_array(rr.charge, (nh,))
        q64 = np.ctypeslib.as_array(rr.charge64, (nh,)).copy()
    else:
        q64 = np.zeros(0)
    z = np.ctypeslib.as_array(rr.z, (nz,)).copy() if nz else np.zeros(0)
    off = np.ctypeslib.as_array(rr.z_off, (n + 1,)).copy()
    census = {"pairs": rr.n_pairs, "inside": rr.n_inside, "pixels": rr.n_pixels}
    L.orc_digi_free(r)
    return hits, q64, z, off, census


def run(beam: dict, devices: list, materials: list, n: int | None = None, seed_interp: int = 1,
        seed_numba: int = 2, table_override=None) -> OracleRun:
    """Replay one full reference run (tracks + every plane's digitisation) on the CPU."""
    n = int(beam["nmb_particles"] if n is None else n)
    D = len(devices)
    zx0, zy0, t0, table, accepted = draw_interpreter_side(beam, devices, materials, n, seed_interp, table_override)
    mt = MT(seed_numba)
    zstart = np.zeros((n, 4))
    gap = np.zeros(n, dtype=np.int64)
    zkick = np.zeros((n, max(D - 1, 1), 2))
    lib().orc_record_tracks(mt.ptr, n, D, particle_distance(beam), _dp(zstart), _ip(gap), _dp(zkick))
    zstart[0, 2], zstart[0, 3] = zx0, zy0
    gap[0] = t0
    tracks, eloss = run_tracks(beam, devices, materials, n, zstart, gap, zkick, table)
    var = Variates(zstart, gap, zkick, table, accepted)
    hits, q64s = [], []
    census = {"pairs": 0, "inside": 0, "pixels": 0}
    for d in range(D):
        h, q64, z, off, cs = run_digitise_plane(devices, materials, d, n, tracks[3], tracks[4], eloss, accepted, mt=mt)
        hits.append(h)
        q64s.append(q64)
        var.dig_z.append(z)
        var.dig_off.append(off)
        for k in census:
            census[k] += cs[k]
    return OracleRun(tracks, eloss, hits, q64s, var, census)


def threshold_scan(inj_array, numb_inj, column_pitch, row_pitch, threshold, thickness, column_numb, row_numb, seed_numba,
                   column=0, row=0):
    """S-curve of main/threshold_scan.py:42-77 replayed on the CPU: (hit fractions, the normals drawn [bins,inj,2])."""
    inj_array = np.asarray(inj_array, dtype=np.float64)
    bins, numb_inj = len(inj_array), int(numb_inj)
    n = bins * numb_inj
    dev = [{"column": int(column_numb), "row": int(row_numb), "column_pitch": column_pitch, "row_pitch": row_pitch,
            "thickness": thickness, "z_position": 0, "delta_x": column * column_pitch, "delta_y": row * row_pitch,
            "threshold": threshold, "trigger": "untriggered", "material": "silicon"}]
    mat = [{"rad_length": 93650, "Z": 14, "A": 24, "rho": 2.33}]
    energy = np.repeat(inj_array * 1e-6 * 3.6, numb_inj).reshape(1, n)
    hits, q64, z, off, census = run_digitise_plane(dev, mat, 0, n, np.zeros(n), np.zeros(n), energy,
                                                   np.ones(n, np.uint8), mt=MT(seed_numba), zero_radius=True)
    frac = np.bincount((hits["event_number"] - 1) // numb_inj, minlength=bins)[:bins] / numb_inj
    return frac, z.reshape(bins, numb_inj, 2)


def correlation_prepare(device_x, device_y, max_cols=(None, None), max_rows=(None, None), offset=0, event_size=None,
                        event_numb_shift=0):
    """The numpy pre-processing of plot_correlation (main/plotting.py:501-556): event window, event shift, keep only
    events present in both tables, histogram shapes."""
    device_x, device_y = np.copy(device_x), np.copy(device_y)
    if event_size is not None:
        device_x = device_x[(device_x["event_number"] <= offset + event_size) & (offset <= device_x["event_number"])]
        device_y = device_y[(device_y["event_number"] <= offset + event_size) & (offset <= device_y["event_number"])]
    device_y["event_number"] = device_y["event_number"] + event_numb_shift
    device_x = device_x[np.isin(device_x["event_number"], device_y["event_number"])]
    device_y = device_y[np.isin(device_y["event_number"], device_x["event_number"])]
    max_cols, max_rows = list(max_cols), list(max_rows)
    if max_cols[0] is None:
        max_cols = [int(np.max(device_x["column"])) + 1, int(np.max(device_y["column"])) + 1]
        max_rows = [int(np.max(device_x["row"])) + 1, int(np.max(device_y["row"])) + 1]
    return device_x, device_y, max_cols, max_rows


def correlation(device_x, device_y, **kw):
    """(x_corr_hist, y_corr_hist) exactly as plot_correlation builds them before plotting."""
    dx, dy, max_cols, max_rows = correlation_prepare(device_x, device_y, **kw)
    xh = np.zeros(max_cols, dtype=np.int32)
    yh = np.zeros(max_rows, dtype=np.int32)
    c = lambda a, t: np.ascontiguousarray(a, dtype=t)   # noqa: E731
    ev1, r1, c1 = c(dx["event_number"], np.int64), c(dx["row"], np.uint16), c(dx["column"], np.uint16)
    ev2, r2, c2 = c(dy["event_number"], np.int64), c(dy["row"], np.uint16), c(dy["column"], np.uint16)
    u16 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint16))   # noqa: E731
    lib().orc_eventloop(_ip(ev1), u16(r1), u16(c1), len(ev1), _ip(ev2), u16(r2), u16(c2), len(ev2),
                        xh.ctypes.data_as(C.POINTER(C.c_int32)), xh.shape[1],
                        yh.ctypes.data_as(C.POINTER(C.c_int32)), yh.shape[1])
    return xh, yh
