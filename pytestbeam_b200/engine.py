"""Host side of the hot path: owns the library handle, allocates the device buffers with PyTorch and
drives ``ptb_simulate`` / ``ptb_prefix`` / ``ptb_pack_hits`` on the current CUDA stream.

PyTorch is plumbing here (device memory, streams, ``torch.distributed``); all arithmetic happens in the
hand-written kernels of ``csrc/ptb_kernels.cu``.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from . import native as N

HITS_DTYPE = np.dtype([("event_number", "<i8"), ("frame", "<u2"), ("column", "<u2"), ("row", "<u2"),
                       ("charge", "<f4")])

_TOTALS_WORDS = C.sizeof(N.PtbTotals) // 8
_BASES_WORDS = C.sizeof(N.PtbBases) // 8


class PtbError(RuntimeError):
    pass


@dataclass
class Injected:
    """Pre-drawn variates of particles [first, first+n) on the device (deterministic mode)."""
    first: int
    n: int
    tensors: dict
    eloss0_prev: float = 0.0

    def struct(self) -> N.PtbInjected:
        t = self.tensors
        p = lambda k: t[k].data_ptr() if k in t and t[k] is not None else None  # noqa: E731
        return N.PtbInjected(p("zstart"), p("gap"), p("zkick"), p("eloss"), p("accepted"), p("dig_z"), p("dig_off"),
                             float(self.eloss0_prev))

    @staticmethod
    def from_numpy(device, first: int, n: int, *, zstart=None, gap=None, zkick=None, eloss=None, accepted=None,
                   dig_z=None, dig_off=None):
        """Slice whole-run variate tables (numpy; layout documented at PtbInjected in include/ptb.h) to a particle range.

        ``dig_z`` / ``dig_off`` are per-plane lists: normals in consumption order and their CSR offsets [N+1].
        """
        t = {}
        lo, hi = first, first + n
        up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(device)  # noqa: E731
        prev = 0.0
        if zstart is not None:
            t["zstart"] = up(zstart[lo:hi], np.float64)
        if gap is not None:
            t["gap"] = up(gap[lo:hi], np.int64)
        if zkick is not None:
            t["zkick"] = up(zkick[lo:hi], np.float64)
        if eloss is not None:
            t["eloss"] = up(eloss[:, lo:hi], np.float64)
            prev = float(eloss[0, lo - 1]) if lo > 0 else 0.0
        if accepted is not None:
            t["accepted"] = up(accepted[lo:hi], np.uint8)
        if dig_z is not None:
            parts, offs, base = [], [], 0
            for z, off in zip(dig_z, dig_off):
                a, b = int(off[lo]), int(off[hi])
                parts.append(z[a:b])
                offs.append(off[lo:hi + 1] - a + base)
                base += b - a
            t["dig_z"] = up(np.concatenate(parts) if base else np.zeros(1), np.float64)
            t["dig_off"] = up(np.stack(offs), np.int64)
        return Injected(first, n, t, prev)


@dataclass
class RunResult:
    first: int
    n: int
    totals: dict
    slabs: list = field(default_factory=list)      # per plane: dict of device tensors trimmed to the rows produced
    tracks: dict | None = None                     # device tensors, layout of the reference's hit_tables
    bases_out: torch.Tensor | None = None
    compact: dict | None = None                    # PTB_COMPACT: device tensors of the PtbCompactOut
    event_base: int = 0

    def compact_host(self) -> dict:
        """The PtbCompactOut on the host, plane by plane: uint64 rows + uint8 counts [D, n] (wide form); the (charge,
        pos_lo, pos_hi) slices (narrow) or the (word0, word1, word2) slices + row layout (packed) with nibble pairs
        [D, (n+1)//2] and escape entries; uint32 trigger words."""
        D, hits, c = len(self.totals["hits"]), self.totals["hits"], self.compact
        acc = c["accepted"].cpu().numpy().view(np.uint32)
        form = c.get("form", "wide")
        if form in ("narrow", "packed"):
            first = np.concatenate([[0], np.cumsum(hits)]).astype(np.int64)      # the planes lie back to back
            cut = lambda k, dt: [c[k][first[d]:first[d + 1]].cpu().numpy().view(dt) for d in range(D)]  # noqa: E731
            out = {"form": form, "narrow": True, "counts": c["counts"].cpu().numpy().reshape(D, (self.n + 1) // 2),
                   "escapes": c["escapes"][:self.totals["escapes"]].cpu().numpy().view(np.uint64), "accepted": acc}
            if form == "narrow":
                out.update(charge=cut("charge", np.float32), pos_lo=cut("pos_lo", np.uint16), pos_hi=cut("pos_hi", np.uint8))
            else:
                out.update(word0=cut("word0", np.uint16), word1=cut("word1", np.uint16), word2=cut("word2", np.uint16),
                           layout=c["layout"], row_base=[int(v) for v in first[:D]])
            return out
        pf = c["plane_first"]
        return {"form": "wide", "hits": [c["hits"][pf[d]:pf[d] + hits[d]].cpu().numpy().view(np.uint64) for d in range(D)],
                "counts": c["counts"].cpu().numpy().reshape(D, self.n), "accepted": acc}

    def hits_numpy(self, plane: int) -> np.ndarray:
        """Rows of one plane as the reference's structured array (main/device.py:20-28)."""
        if self.compact is not None:
            return expand_compact(N.load(), plane_of(self.compact_host(), plane), self.n, self.event_base)
        s = self.slabs[plane]
        out = np.zeros(int(s["event"].numel()), dtype=HITS_DTYPE)
        out["event_number"] = s["event"].cpu().numpy()
        out["column"] = s["col"].cpu().numpy().view(np.uint16)
        out["row"] = s["row"].cpu().numpy().view(np.uint16)
        out["charge"] = s["charge"].cpu().numpy()
        return out


class Simulator:
    """One telescope set-up on one GPU."""

    def __init__(self, beam: dict, devices: list, materials: list, device=None, pool_entries: int = 0):
        if not torch.cuda.is_available():
            raise PtbError("pytestbeam_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.lib = N.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.beam, self.devices, self.materials = dict(beam), list(devices), list(materials)
        self.D = len(devices)
        self._opts = dict(pool_entries=int(pool_entries))
        self._h = None
        self._create()
        self._hits_per_particle = self._rows_per_particle = None

    # -- handle ------------------------------------------------------------------------------------
    def _create(self):
        self.close()
        h = C.c_void_p()
        opt = N.PtbOptions(self._opts["pool_entries"], 0, 0.0)
        with torch.cuda.device(self.device):
            rc = self.lib.ptb_create(C.byref(N.pack_beam(self.beam)), N.pack_devices(self.devices, self.materials),
                                     self.D, C.byref(opt), C.byref(h))
        if rc != 0:
            msg = self.lib.ptb_last_error(h).decode() if h else "allocation failed"
            if h:
                self.lib.ptb_destroy(h)
            raise ValueError(f"ptb_create failed ({rc}): {msg}")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.lib.ptb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise PtbError(f"{what} failed ({rc}): {self.lib.ptb_last_error(self._h).decode()}")

    # -- buffers -----------------------------------------------------------------------------------
    def alloc_slabs(self, capacities, keep64=False):
        slabs = []
        for c in capacities:
            c = int(c)
            s = {"event": torch.empty(c, dtype=torch.int64, device=self.device),
                 "col": torch.empty(c, dtype=torch.int16, device=self.device),
                 "row": torch.empty(c, dtype=torch.int16, device=self.device),
                 "charge": torch.empty(c, dtype=torch.float32, device=self.device)}
            if keep64:
                s["charge64"] = torch.empty(c, dtype=torch.float64, device=self.device)
            slabs.append(s)
        return slabs

    def alloc_compact(self, n, capacities, narrow=False, form: str | None = None):
        """Device buffers of a PtbCompactOut for n particles and the given rows per plane; form = "wide", "narrow" or
        "packed" (include/ptb.h)."""
        form = form or ("narrow" if narrow else "wide")
        first = np.concatenate([[0], np.cumsum([int(c) for c in capacities])]).astype(np.int64)
        rows, dev = int(first[-1]), self.device
        out = {"plane_first": [int(v) for v in first], "form": form, "narrow": form != "wide",
               "accepted": torch.empty((int(n) + 31) // 32, dtype=torch.int32, device=dev)}
        if form == "wide":
            out.update(hits=torch.empty(rows, dtype=torch.int64, device=dev),
                       counts=torch.empty(self.D * int(n), dtype=torch.uint8, device=dev))
            return out
        # counts of 15 and more: one entry per 256 (particle, plane) pairs is far above any telescope's need
        out.update(counts=torch.empty(self.D * ((int(n) + 1) // 2), dtype=torch.uint8, device=dev),
                   escapes=torch.empty(max(1024, self.D * int(n) // 256), dtype=torch.int64, device=dev))
        if form == "narrow":
            out.update(charge=torch.empty(rows, dtype=torch.float32, device=dev),
                       pos_lo=torch.empty(rows, dtype=torch.int16, device=dev),
                       pos_hi=torch.empty(rows, dtype=torch.uint8, device=dev))
        elif form == "packed":
            layout = self.packed_layout()
            if layout is None:
                raise PtbError("the packed compressed rows do not fit this set-up (ptb_packed_layout)")
            out.update(layout=layout, **{k: torch.empty(rows, dtype=torch.int16, device=dev) for k in ("word0", "word1", "word2")})
        else:
            raise ValueError(f"unknown compressed-row form {form!r}")
        return out

    def packed_layout(self):
        """The PtbPackedLayout of this set-up, or None when PTB_COMPACT_PACKED does not fit it."""
        if getattr(self, "_packed_layout", 0) == 0:
            L = N.PtbPackedLayout()
            self._packed_layout = L if self.lib.ptb_packed_layout(self._h, C.byref(L)) == 0 else None
        return self._packed_layout

    def compact_form(self, seed=12345, pilot=32768) -> str:
        """The most compressed row form this set-up supports: "packed" (6-byte rows), "narrow" (7-byte rows) or "wide".
        The candidates are tried on a pilot range: a form whose escape list or exponent window overflows there is
        passed over (a later overflow of a chunk is retried by the pipeline)."""
        if getattr(self, "_compact_form", None) is None:
            forms = (["packed"] if self.packed_layout() is not None else []) + \
                    (["narrow"] if all(int(d["column"]) <= 4095 and int(d["row"]) <= 4095 for d in self.devices) else [])
            chosen = "wide"
            for form in forms:
                try:
                    r = self.run(0, pilot, seed=seed, capacities=[pilot * 64] * self.D, compact=form)
                    chosen = form
                    if self._hits_per_particle is None:      # the same pilot sizes the buffers (hits_per_particle)
                        self._hits_per_particle = np.array(r.totals["hits"], dtype=np.float64) / pilot
                        self._rows_per_particle = np.array(r.totals["reserved"], dtype=np.float64) / pilot
                    break
                except PtbError:
                    continue
            self._compact_form = chosen
        return self._compact_form

    def narrow_ok(self, seed=12345, pilot=32768) -> bool:
        """Does a form with 4-bit counts and back-to-back rows (narrow or packed) fit this set-up?"""
        return self.compact_form(seed, pilot) != "wide"

    def alloc_tracks(self, n):
        D = self.D
        f8 = lambda m: torch.empty(m, dtype=torch.float64, device=self.device)  # noqa: E731
        return {"energy": f8(n * D), "angle_x": f8(n * D), "angle_y": f8(n * D), "x": f8(n * D), "y": f8(n * D),
                "z": f8(n * D), "time_stamp": torch.empty(n * D, dtype=torch.int64, device=self.device),
                "eloss": f8(D * n), "accepted": torch.empty(n, dtype=torch.uint8, device=self.device)}

    def workspace_bytes(self, n, capacities, flags, scratch_rows=None) -> int:
        caps = (C.c_uint64 * self.D)(*[int(c) for c in capacities])
        rows = (C.c_uint64 * self.D)(*[int(c) for c in scratch_rows]) if scratch_rows is not None else None
        return int(self.lib.ptb_workspace_bytes(self._h, int(n), caps, rows, int(flags)))

    def new_bases(self) -> torch.Tensor:
        return torch.zeros(_BASES_WORDS, dtype=torch.int64, device=self.device)

    def hits_per_particle(self, seed=12345, pilot=32768) -> np.ndarray:
        """Hits per particle and plane measured on a small pilot range (sizes the slabs of large runs)."""
        if self._hits_per_particle is None:
            r = self.run(0, pilot, seed=seed, capacities=[pilot * 64] * self.D)
            self._hits_per_particle = np.array(r.totals["hits"], dtype=np.float64) / pilot
            self._rows_per_particle = np.array(r.totals["reserved"], dtype=np.float64) / pilot
        return self._hits_per_particle

    @staticmethod
    def _with_margin(n, per_particle):
        return [int(n * h * 1.05 + 8 * np.sqrt(n * max(h, 1e-3) * 30) + 4096) for h in per_particle]

    def default_capacities(self, n: int):
        """Slab rows per plane for n particles: the pilot's hits per particle plus a margin."""
        return self._with_margin(n, self.hits_per_particle())

    def default_scratch_rows(self, n: int):
        """Scratch rows per plane for n particles (PtbRun.scratch_rows): the pilot's reserved rows plus a margin."""
        self.hits_per_particle()
        return self._with_margin(n, self._rows_per_particle)

    # -- one call of ptb_simulate --------------------------------------------------------------------
    def launch(self, first, n, *, seed=0, injected: Injected | None = None, flags=N.PTB_DIGITISE, slabs=None,
               tracks=None, bases_in=None, bases_out=None, totals=None, workspace=None, scratch_rows=None,
               compact=None):
        """Enqueue one ptb_simulate call on the current stream (no synchronisation)."""
        run = N.PtbRun()
        run.first, run.n, run.seed, run.flags = int(first), int(n), int(seed) & (2**64 - 1), int(flags)
        inj_struct = None
        if injected is not None:
            if injected.first != first or injected.n != n:
                raise ValueError("injected tables do not cover the requested particle range")
            inj_struct = injected.struct()
            run.injected = C.pointer(inj_struct)
        if tracks is not None:
            for k in ("energy", "angle_x", "angle_y", "x", "y", "z", "time_stamp", "eloss", "accepted"):
                if tracks.get(k) is not None:
                    setattr(run.tracks, k, tracks[k].data_ptr())
        if slabs is not None:
            for d, s in enumerate(slabs):
                run.slabs[d] = N.PtbHitSlab(s["event"].data_ptr(), s["col"].data_ptr(), s["row"].data_ptr(),
                                            s["charge"].data_ptr(),
                                            s["charge64"].data_ptr() if "charge64" in s else None,
                                            int(s["event"].numel()))
        if compact is not None:
            form = compact.get("form", "narrow" if compact.get("narrow") else "wide")
            if form == "wide":
                run.compact.hits = compact["hits"].data_ptr()
            else:
                run.compact.escapes = compact["escapes"].data_ptr()
                run.compact.escape_capacity = int(compact["escapes"].numel())
                if form == "narrow":
                    run.flags |= N.PTB_COMPACT_NARROW
                    run.compact.charge, run.compact.pos_lo = compact["charge"].data_ptr(), compact["pos_lo"].data_ptr()
                    run.compact.pos_hi = compact["pos_hi"].data_ptr()
                else:
                    run.flags |= N.PTB_COMPACT_PACKED
                    run.compact.word0, run.compact.word1 = compact["word0"].data_ptr(), compact["word1"].data_ptr()
                    run.compact.word2 = compact["word2"].data_ptr()
            run.compact.counts, run.compact.accepted = compact["counts"].data_ptr(), compact["accepted"].data_ptr()
            for d, v in enumerate(compact["plane_first"]):
                run.compact.plane_first[d] = int(v)
        run.bases_in = bases_in.data_ptr() if bases_in is not None else None
        run.bases_out = bases_out.data_ptr() if bases_out is not None else None
        run.totals = totals.data_ptr()
        run.workspace = workspace.data_ptr()
        run.workspace_bytes = workspace.numel()
        if scratch_rows is not None:
            for d, r in enumerate(scratch_rows):
                run.scratch_rows[d] = int(r)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            self._check(self.lib.ptb_simulate(self._h, C.byref(run), C.c_void_p(stream)), "ptb_simulate")

    @staticmethod
    def decode_totals(t: torch.Tensor, D: int) -> dict:
        a = t.cpu().numpy()
        u = a.view(np.uint64)
        return {"hits": [int(v) for v in u[:D]], "accepted": int(u[32]), "time_sum": int(a[33]), "pairs": int(u[34]),
                "inside": int(u[35]), "pixels": int(u[36]), "error": int(u[37] & 0xFFFFFFFF), "escapes": int(u[37] >> 32),
                "reserved": [int(v) for v in u[38:38 + D]]}

    def run(self, first, n, *, seed=0, injected=None, digitise=True, write_tracks=False, tracks_in=None,
            keep_charge64=False, capacities=None, bases_in=None, zero_radius=False, scratch_rows=None,
            compact=False) -> RunResult:
        """Simulate particles [first, first+n), synchronise and return trimmed device buffers.

        Retries with exact slab / scratch sizes when the device reports an overflow (its totals stay exact).
        """
        first, n = int(first), int(n)
        flags = N.PTB_RECYCLE_SLABS   # every call of run() gets fresh slabs
        if digitise:
            flags |= N.PTB_DIGITISE
        if tracks_in is not None:
            flags |= N.PTB_TRACKS_FROM_TABLE
        if write_tracks:
            flags |= N.PTB_WRITE_TRACKS
        if keep_charge64:
            flags |= N.PTB_KEEP_CHARGE64
        if zero_radius:
            flags |= N.PTB_ZERO_RADIUS
        if compact:
            flags |= N.PTB_COMPACT
        if capacities is None:
            capacities = [4 * n + 4096] * self.D if digitise else [0] * self.D
        for attempt in range(8):
            slabs = self.alloc_slabs(capacities, keep_charge64) if (digitise and not compact) else None
            cmp_out = self.alloc_compact(n, capacities, form=compact if isinstance(compact, str) else "wide") if compact else None
            if compact and scratch_rows is None:
                scratch_rows = [N.default_scratch_rows(int(c)) for c in capacities]
            tracks = tracks_in if tracks_in is not None else (self.alloc_tracks(n) if write_tracks else None)
            totals = torch.zeros(_TOTALS_WORDS, dtype=torch.int64, device=self.device)
            bases_out = self.new_bases()
            ws = torch.empty(self.workspace_bytes(n, capacities, flags, scratch_rows), dtype=torch.uint8,
                             device=self.device)
            self.launch(first, n, seed=seed, injected=injected, flags=flags, slabs=slabs, tracks=tracks,
                        bases_in=bases_in, bases_out=bases_out, totals=totals, workspace=ws, scratch_rows=scratch_rows,
                        compact=cmp_out)
            torch.cuda.synchronize(self.device)
            tot = self.decode_totals(totals, self.D)
            err = tot["error"]
            if err & ~(N.PTB_DEV_BOX_TOO_LARGE | N.PTB_DEV_SLAB_OVERFLOW | N.PTB_DEV_SCRATCH_OVERFLOW |
                       N.PTB_DEV_COUNT_OVERFLOW | N.PTB_DEV_PACK_OVERFLOW):
                raise PtbError(f"device reported an internal consistency failure (error bits {err:#x})")
            if err & N.PTB_DEV_BOX_TOO_LARGE:
                raise PtbError("a cluster bounding box exceeded 2^20 pixels (check pitch / geometry)")
            if err & (N.PTB_DEV_SLAB_OVERFLOW | N.PTB_DEV_SCRATCH_OVERFLOW):
                capacities = [max(int(h), 1) for h in tot["hits"]]
                scratch_rows = [max(int(r), 1) for r in tot["reserved"]]
                continue
            if err & N.PTB_DEV_PACK_OVERFLOW:
                raise PtbError("the row packer met a charge below its plane's threshold (internal inconsistency)")
            if err & N.PTB_DEV_COUNT_OVERFLOW:
                raise PtbError("a particle left more than 255 rows on one plane (or the narrow form ran out of escape "
                               "entries): run this range without compact")
            if digitise and not compact:
                slabs = [{k: v[:tot["hits"][d]] for k, v in s.items()} for d, s in enumerate(slabs)]
            ev0 = int(bases_in[0].item()) if bases_in is not None else 0
            return RunResult(first, n, tot, slabs or [], tracks, bases_out, cmp_out, ev0)
        raise PtbError("ptb_simulate kept overflowing its buffers")

    def timing(self, on: bool):
        """Bracket the k_tracks / k_digitise launches with CUDA events on their stream (see timing_read)."""
        self._check(self.lib.ptb_timing_enable(self._h, int(on)), "ptb_timing_enable")

    def timing_read(self):
        """(k_tracks ms, k_digitise ms, calls) summed since the last read; waits for the recorded events."""
        a, b, n = C.c_double(), C.c_double(), C.c_uint64()
        self._check(self.lib.ptb_timing_read(self._h, C.byref(a), C.byref(b), C.byref(n)), "ptb_timing_read")
        return float(a.value), float(b.value), int(n.value)

    def timing_read_stages(self):
        """([k_tracks, k_digitise, k_digitise_big + scans, k_gather] ms, calls) summed since the last read."""
        st, n = (C.c_double * 4)(), C.c_uint64()
        self._check(self.lib.ptb_timing_read_stages(self._h, C.byref(st), C.byref(n)), "ptb_timing_read_stages")
        return [float(v) for v in st], int(n.value)

    # -- helpers -------------------------------------------------------------------------------------
    def prefix(self, first, n, *, seed=0, injected: Injected | None = None):
        """(accepted particles, sum of time gaps) of a particle range."""
        out = torch.zeros(2, dtype=torch.int64, device=self.device)
        inj = C.pointer(injected.struct()) if injected is not None else None
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            self._check(self.lib.ptb_prefix(self._h, int(first), int(n), int(seed) & (2**64 - 1), inj, out.data_ptr(),
                                            C.c_void_p(stream)), "ptb_prefix")
        a = out.cpu().numpy()
        return int(a[0]), int(a[1])

    def digest(self, slab: dict, n: int, row_base: int, out: torch.Tensor):
        """Add the checksums of rows [0, n) of a slab to out (7 int64 on the device); see ptb_digest in include/ptb.h."""
        if n:
            s = N.PtbHitSlab(slab["event"].data_ptr(), slab["col"].data_ptr(), slab["row"].data_ptr(),
                             slab["charge"].data_ptr(), None, int(slab["event"].numel()))
            stream = torch.cuda.current_stream(self.device).cuda_stream
            with torch.cuda.device(self.device):
                rc = self.lib.ptb_digest(C.byref(s), int(n), int(row_base) & (2**64 - 1), out.data_ptr(), C.c_void_p(stream))
            if rc != 0:
                raise PtbError(f"ptb_digest failed ({rc})")

    def pack_hits(self, slab: dict, n: int | None = None) -> torch.Tensor:
        """SoA slab -> device tensor of packed 18-byte reference records (uint8 [n*18])."""
        n = int(slab["event"].numel() if n is None else n)
        out = torch.empty(n * 18, dtype=torch.uint8, device=self.device)
        if n:
            s = N.PtbHitSlab(slab["event"].data_ptr(), slab["col"].data_ptr(), slab["row"].data_ptr(),
                             slab["charge"].data_ptr(), None, int(slab["event"].numel()))
            stream = torch.cuda.current_stream(self.device).cuda_stream
            with torch.cuda.device(self.device):
                rc = self.lib.ptb_pack_hits(C.byref(s), n, out.data_ptr(), C.c_void_p(stream))
            if rc != 0:
                raise PtbError(f"ptb_pack_hits failed ({rc})")
        return out


_ROW_KEYS = {"wide": ("hits",), "narrow": ("charge", "pos_lo", "pos_hi"), "packed": ("word0", "word1", "word2")}


def _form_of(c: dict) -> str:
    return c.get("form") or ("narrow" if c.get("narrow") else "wide")


def plane_of(c: dict, plane: int) -> dict:
    """One plane's arrays of a host-side PtbCompactOut (RunResult.compact_host / HostChunk.compact)."""
    form = _form_of(c)
    out = {k: c[k][plane] for k in _ROW_KEYS[form]}
    out.update(form=form, narrow=form != "wide", counts=c["counts"][plane], accepted=c["accepted"], plane=plane,
               escapes=c.get("escapes"), layout=c.get("layout"),
               row_base=int(c["row_base"][plane]) if c.get("row_base") is not None else 0)
    return out


def _expand_args(c: dict):
    """(form, rows, leading pointer arguments of the plane's expansion call) for one plane of a host-side PtbCompactOut."""
    form = _form_of(c)
    keys = _ROW_KEYS[form]
    rows = int(c[keys[0]].shape[0])
    counts = np.ascontiguousarray(c["counts"], dtype=np.uint8)
    args = [c[k].ctypes.data if rows else None for k in keys] + [counts.ctypes.data]
    keep = [counts]                                      # arrays that must outlive the call
    if form != "wide":
        esc = c.get("escapes")
        esc = np.zeros(0, np.uint64) if esc is None else np.ascontiguousarray(esc, dtype=np.uint64)
        keep.append(esc)
        args += [esc.ctypes.data if len(esc) else None, len(esc), int(c.get("plane", 0))]
        if form == "packed":
            args += [C.byref(c["layout"]), int(c.get("row_base", 0))]
    args.append(c["accepted"].ctypes.data)
    return form, rows, args, keep


_EXPAND = {"wide": "ptb_expand_hits", "narrow": "ptb_expand_hits_narrow", "packed": "ptb_expand_hits_packed"}


def expand_compact(lib, c: dict, n: int, event_base: int, out: np.ndarray | None = None, threads: int = 0) -> np.ndarray:
    """One plane of a PtbCompactOut (host arrays, see plane_of) -> the reference's 18-byte records (``ptb_expand_hits``
    / ``ptb_expand_hits_narrow`` / ``ptb_expand_hits_packed``).

    wide: hits uint64 rows, counts uint8 [n]; narrow: charge f32, pos_lo u16, pos_hi u8 rows; packed: word0..2 u16 rows
    and the row layout; narrow and packed: counts [(n+1)//2] nibble pairs, escapes uint64 (the call's entries for counts
    of 15 and more) and the plane's index; accepted: uint32 [(n+31)//32]; out: optional destination (any writable buffer
    of that many records, e.g. a slice of a memory-mapped HDF5 dataset)."""
    form, rows, args, _keep = _expand_args(c)
    if out is None:
        out = np.empty(rows, dtype=HITS_DTYPE)
    if out.shape[0] != rows or out.dtype.itemsize != 18 or not out.flags["C_CONTIGUOUS"]:
        raise ValueError("destination must be a contiguous array of as many 18-byte records as the plane has rows")
    got = getattr(lib, _EXPAND[form])(*args, int(n), int(event_base), rows, out.ctypes.data if rows else None, int(threads))
    if got != rows:
        raise PtbError(f"ptb_expand_hits: counts add up to something else than {rows} rows ({got})")
    return out


def expand_compact_to_fd(lib, c: dict, n: int, event_base: int, fd: int, file_offset: int, threads: int = 0) -> int:
    """expand_compact with a file as destination (``ptb_expand_hits*_fd``): the plane's records are written to descriptor
    `fd` from byte `file_offset` on, block by block.  Returns the number of rows."""
    form, rows, args, _keep = _expand_args(c)
    got = getattr(lib, _EXPAND[form] + "_fd")(*args, int(n), int(event_base), rows, int(fd), int(file_offset), int(threads))
    if got != rows:
        raise PtbError(f"ptb_expand_hits_fd: {'a write failed' if got == -4 else 'counts and rows disagree'} ({got}, {rows} rows)")
    return rows
