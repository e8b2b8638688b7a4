"""Error behaviour of the C ABI: bad arguments come back as negative codes with a message, nothing is thrown or
silently ignored."""
import ctypes as C

import pytest

from pytestbeam_b200 import config as cfg

pytestmark = pytest.mark.gpu


def _sim():
    from pytestbeam_b200.engine import Simulator
    beam, devices, materials, names = cfg.flatten(cfg.c1(100))
    return Simulator(beam, devices, materials)


def test_invalid_calls_are_refused_with_a_message():
    import torch
    from pytestbeam_b200 import native as N
    from pytestbeam_b200.engine import PtbError, _TOTALS_WORDS
    sim = _sim()
    lib = sim.lib
    run = N.PtbRun()
    assert lib.ptb_simulate(sim._h, C.byref(run), None) == -1               # no totals
    totals = torch.zeros(_TOTALS_WORDS, dtype=torch.int64, device=sim.device)
    run.totals = totals.data_ptr()
    run.n = 0
    assert lib.ptb_simulate(sim._h, C.byref(run), None) == -1
    assert b"empty" in lib.ptb_last_error(sim._h)
    run.n, run.flags = 1000, N.PTB_DIGITISE
    assert lib.ptb_simulate(sim._h, C.byref(run), None) == -1
    assert b"slab" in lib.ptb_last_error(sim._h)
    slabs = sim.alloc_slabs([4096] * sim.D)
    for d, s in enumerate(slabs):
        run.slabs[d] = N.PtbHitSlab(s["event"].data_ptr(), s["col"].data_ptr(), s["row"].data_ptr(),
                                    s["charge"].data_ptr(), None, 4096)
    assert lib.ptb_simulate(sim._h, C.byref(run), None) == -1
    assert b"workspace" in lib.ptb_last_error(sim._h)
    run.flags = N.PTB_TRACKS_FROM_TABLE
    assert lib.ptb_simulate(sim._h, C.byref(run), None) == -1
    run.flags = N.PTB_DIGITISE | N.PTB_TRACKS_FROM_TABLE | N.PTB_WRITE_TRACKS
    assert lib.ptb_simulate(sim._h, C.byref(run), None) == -1
    assert b"exclusive" in lib.ptb_last_error(sim._h)
    assert lib.ptb_simulate(None, C.byref(run), None) == -1
    assert lib.ptb_prefix(sim._h, 0, 10, 1, None, None, None) == -1
    with pytest.raises(PtbError):
        sim.launch(0, 1000, flags=N.PTB_DIGITISE, slabs=slabs, totals=totals,
                   workspace=torch.empty(16, dtype=torch.uint8, device=sim.device))
    torch.cuda.synchronize()


def test_unsupported_options_and_geometry():
    from pytestbeam_b200.engine import Simulator
    beam, devices, materials, names = cfg.flatten(cfg.c1(100))
    with pytest.raises(ValueError, match="variants"):
        Simulator(beam, devices, materials, pool_entries=77)
    bad = [dict(d) for d in devices]
    bad[3]["z_position"] = bad[2]["z_position"]
    with pytest.raises(ValueError, match="z_position"):
        Simulator(beam, bad, materials)
    bad = [dict(d) for d in devices]
    bad[1]["column"] = 70000
    with pytest.raises(ValueError, match="geometry"):
        Simulator(beam, bad, materials)
    b2 = dict(beam, particle_rate=0)
    with pytest.raises(ValueError, match="particle_rate"):
        Simulator(b2, devices, materials)


def test_injected_tables_must_be_complete():
    from pytestbeam_b200.engine import Injected, PtbError
    from tests import helpers as H
    n = 600
    setup, run = H.oracle_run("c1", n, 1, 2)
    from pytestbeam_b200.engine import Simulator
    sim = Simulator(*setup[:3])
    v = run.variates
    inj = Injected.from_numpy(sim.device, 0, n, zstart=v.zstart, gap=v.gap, zkick=v.zkick, eloss=v.eloss,
                              accepted=v.accepted)                   # digitisation normals missing
    with pytest.raises(PtbError, match="incomplete"):
        sim.run(0, n, injected=inj)
    with pytest.raises(ValueError):
        sim.run(10, n, injected=inj, digitise=False, write_tracks=True)     # tables do not cover that range
