"""Child process of tests/test_emu.py -- TEST INFRASTRUCTURE ONLY.
Loads tests/emu/libroms_b200_emu.so (the kernel SOURCES of roms_b200/csrc built with g++, see tests/emu/include/cuda_runtime.h)
in place of the CUDA library and runs the per-kernel parity protocol of tests/test_gpu_parity.py against the oracle.
Runs in its own process because the switches it needs (no CUDA graph, no programmatic launch, step3d_t layout) are read once
per process by the library.   usage: emu_worker.py [driver|tiles|eos] APP Lm Mm N NSTEPS [v8|v6|v4]"""
import os
import sys

os.environ["ROMS_B200_NO_GRAPH"] = "1"
os.environ["ROMS_B200_NO_PDL"] = "1"
if "v4" in sys.argv:                          # the column-march fallback (closed W/E walls, N < 4): k_step3d_t4.cu
    os.environ["ROMS_B200_STEP3D_T_V4"] = "1"
elif "v6" in sys.argv:                        # the round-1 warp-specialised kernel k_step3d_t6.cu (fallback for shapes the production
    os.environ["ROMS_B200_S3T_V8"] = "0"      # kernel declines): named barriers and the warp vote emulated
# default ("v8"): the production dispatch -> k_step3d_t8.cu (mbarriers, TMA boxes and tensor memory emulated)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np  # noqa: E402
import oracle_lib as ol  # noqa: E402
import roms_b200 as rb  # noqa: E402

rb.lib.Lib._inst = rb.lib.Lib(path=os.path.join(HERE, "emu", os.environ.get("EMU_WORKER_LIB", "libroms_b200_emu.so")))
from parity_common import GPU_PHASE, TRANSCENDENTAL, make_pair, push, diff_fields, run_phase_gpu  # noqa: E402


def main():
    app, Lm, Mm, N, nsteps = (int(x) for x in sys.argv[1:6])
    o, ctx = make_pair(app, Lm, Mm, N)
    not_bitwise = set()
    for step in range(nsteps):
        for ph in ol.PHASES:
            emu = ph in GPU_PHASE or ph in ("vmix", "step2d_loop")
            if ph == "bulk_flux" and app != ol.BENCHMARK:
                emu = False
            if emu:
                push(o, ctx)
                indx1 = run_phase_gpu(o, ctx, ph)
            o.phase(ph)
            if ph == "diag":
                push(o, ctx)
                d, ref = ctx.diag_full(o.stepping()["nstp"]), o.diag_full()
                assert np.array_equal(d, ref), ("diag", step, d, ref)
            if not emu:
                continue
            bitwise = ph not in TRANSCENDENTAL and not (ph == "pre_step3d" and app == ol.BENCHMARK)
            bad = []
            for n, (dmax, scale, eq) in diff_fields(o, ctx).items():
                if not eq:
                    not_bitwise.add((ph, n))
                if (bitwise and not eq) or (not bitwise and dmax > 1e-12 * max(scale, 1e-300)):
                    bad.append((n, dmax, scale))
            assert not bad, "step %d phase %s: fields differ from the oracle: %s" % (step, ph, bad)
            if ph == "step2d_loop":
                assert indx1 == o.stepping()["indx1"]
    ctx.close()
    print("EMU-PARITY-OK app=%d %dx%dx%d steps=%d ; within tolerance but not bit-identical: %s" % (app, Lm, Mm, N, nsteps, sorted(not_bitwise)))


def driver():
    """ROMS_initialize / ROMS_run of the C++ host driver on the emulated kernels vs the oracle: with glibc on both sides the
    whole loop is bit-identical, including KPP, bulk fluxes and the analytical initial state."""
    app, Lm, Mm, N, nsteps = (int(x) for x in sys.argv[2:7])
    o = ol.Oracle(app, Lm, Mm, N)
    o.initial()
    o.phase("begin")
    d = rb.Driver(rb.default_config(app, Lm, Mm, N))
    assert d.nfast == o.dims()["nfast"]
    skip = {"xr", "yr", "lonr", "latr"}        # set only on Istr-1..Iend+1 by the reference; the product fills the whole row
    bad = [n for n in rb.FIELD_NAMES if n not in skip and not np.array_equal(o.get(n), d.ctx.download(n))]
    assert not bad, ("start state", bad)
    for ph in ol.PHASES[1:]:
        o.phase(ph)
    o.step(nsteps - 1)
    d.run(nsteps)                              # device-resident: forcing evaluated by set_data_kernel, diag launched in every step
    bad = [n for n in rb.FIELD_NAMES if n not in skip and not np.array_equal(o.get(n), d.ctx.download(n))]
    assert not bad, ("after %d steps" % nsteps, bad)
    # host forcing (set_data on the host, upload, diag read back every step): the diag returned is that of the last step's start
    diag = d.run(2, host_forcing=True)
    o.step(1)
    for ph in ol.PHASES[:4]:
        o.phase(ph)
    ref = o.diag_full()
    assert np.array_equal(diag, ref[:3]) and np.array_equal(d.ctx.diag_last(), ref), (diag, d.ctx.diag_last(), ref)
    for ph in ol.PHASES[4:]:
        o.phase(ph)
    bad = [n for n in ("zeta", "ubar", "vbar", "u", "v", "t") if not np.array_equal(o.get(n), d.ctx.download(n))]
    assert not bad, ("host forcing", bad)
    # output snapshot (data path only: the emulation is synchronous)
    names = ["zeta", "u", "t"]
    ref_s = {n: d.ctx.download(n) for n in names}
    views = d.ctx.snapshot_begin(names)
    try:
        d.ctx.snapshot_begin(names)
        raise AssertionError("second snapshot_begin accepted while one is pending")
    except RuntimeError:
        pass
    d.run(1)
    d.ctx.snapshot_end()
    assert all(np.array_equal(views[n], ref_s[n]) for n in names)
    # blow-up: diag.F:512-542 sets exit_flag=1 and the driver stops (main3d.F:362)
    u = d.ctx.download("u")
    u[u.size // 3] = 1.0e3                      # an interior point of time level 1 ...
    u[u.size // 2 + u.size // 3] = 1.0e3        # ... and of level 2 (diag reads level nstp)
    d.ctx.upload("u", u)
    try:
        d.run(1, host_forcing=True)
        raise AssertionError("blow-up not detected")
    except RuntimeError:
        pass
    d.finalize()
    print("EMU-DRIVER-OK app=%d %dx%dx%d steps=%d" % (app, Lm, Mm, N, nsteps))


def tiles():
    """Tiling invariance on the emulated kernels (the reference's acceptance criterion, ROMS/Bin/verify.sh): NtileI x NtileJ ranks,
    one host thread and one mirror each, halo swaps as in-process copies (tests/emu/emu_rt.cpp), against ONE tile."""
    import threading
    app, Lm, Mm, N, nsteps, nti, ntj = (int(x) for x in sys.argv[2:9])
    one = rb.Driver(rb.default_config(app, Lm, Mm, N))
    one.run(nsteps)
    one.ctx._bounds = one.bounds()
    cfg = rb.default_config(app, Lm, Mm, N)
    cfg.NtileI, cfg.NtileJ = nti, ntj
    world = nti * ntj
    ds = [rb.Driver(cfg, tile=r) for r in range(world)]
    for r, d in enumerate(ds):
        d.comm_init(r, world, bytes(128))
        d.ctx._bounds = d.bounds()
    errs, diags = [], [None] * world

    def work(r, fn):
        try:
            fn(r)
        except Exception as e:      # noqa: BLE001
            errs.append((r, repr(e)))

    def par(fn):
        th = [threading.Thread(target=work, args=(r, fn)) for r in range(world)]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errs, errs

    par(lambda r: ds[r].run(nsteps))
    fields = [("zeta", 1, 1, 1), ("zeta", 2, 1, 1), ("ubar", 1, 1, 1), ("vbar", 2, 1, 1), ("u", 1, 1, N), ("u", 2, 1, N), ("v", 1, 1, N),
              ("v", 2, 1, N), ("t", 1, 1, N), ("t", 2, 1, N), ("t", 1, 2, N), ("t", 2, 2, N), ("wvel", 1, 1, N + 1), ("Akv", 1, 1, N + 1),
              ("W", 1, 1, N + 1), ("Huon", 1, 1, N), ("rho", 1, 1, N)]
    bad = []
    for n, l, m, nk in fields:
        ref = one.ctx.download_interior(n, l, m, nk)
        for d in ds:
            b = d.ctx._bounds
            if not np.array_equal(d.ctx.download_interior(n, l, m, nk), ref[:, b.Jstr - 1:b.Jend, b.Istr - 1:b.Iend]):
                bad.append((n, l, m, (b.Itile, b.Jtile)))
    assert not bad, bad

    def last(r):
        ds[r].run(1, host_forcing=True)
        diags[r] = ds[r].ctx.diag_last()
    par(last)
    one.run(1, host_forcing=True)
    ref = one.ctx.diag_last()
    for dg in diags:
        assert np.allclose(dg[:3], ref[:3], rtol=1e-13, atol=0.0) and np.array_equal(dg[3:], ref[3:]), (dg, ref)
    print("EMU-TILES-OK app=%d %dx%dx%d steps=%d tiles=%dx%d" % (app, Lm, Mm, N, nsteps, nti, ntj))


def restart():
    """PERFECT_RESTART: the field list of roms_b200_restart_fields + the stepping integers written after 5 steps, uploaded into a
    freshly initialised context, must continue bit-identically to the uninterrupted run."""
    print("EMU-RESTART-OK", perfect_restart_check(int(sys.argv[2]), tuple(int(x) for x in sys.argv[3:6])))


def perfect_restart_check(app, grid, use_snapshot=False):
    import ctypes as C
    L = rb.lib.Lib.get().L
    L.roms_b200_restart_fields.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.roms_b200_restart_finish.argtypes = [C.c_void_p]
    d = rb.Driver(rb.default_config(app, *grid))
    d.run(5)
    ids = (C.c_int * 32)()
    n = L.roms_b200_restart_fields(d.ctx.h, ids, 32)
    assert n >= 15
    names = [rb.FIELD_NAMES[ids[q]] for q in range(n)]
    if use_snapshot:                           # the asynchronous output path: the loop keeps running while the record drains
        views = d.ctx.snapshot_begin(names)
        st, tm = d.ctx.get_stepping()
        d.run(4)
        d.ctx.snapshot_end()
        snap = {nm: views[nm].copy() for nm in names}
    else:
        snap = {nm: d.ctx.download(nm) for nm in names}
        st, tm = d.ctx.get_stepping()
        d.run(4)
    ref = {nm: d.ctx.download(nm) for nm in rb.FIELD_NAMES}
    d2 = rb.Driver(rb.default_config(app, *grid))
    for nm in names:
        d2.ctx.upload(nm, snap[nm])
    d2.ctx.set_stepping(st["iic"], st["ntfirst"], st["nstp"], st["nnew"], st["nrhs"], st["indx1"], tm)
    assert L.roms_b200_restart_finish(d2.ctx.h) == 0
    d2.run(4)
    bad = [nm for nm in ("zeta", "ubar", "vbar", "u", "v", "t", "ru", "rv", "Akv", "Akt", "W", "rho", "Zt_avg1", "Hz")
           if not np.array_equal(ref[nm], d2.ctx.download(nm))]
    assert not bad, bad
    d.finalize(); d2.finalize()
    return names


def eos():
    """rho_eos_kernel (emulated) against the reference's own check values, rho_eos.F:21-29."""
    from parity_common import eos_check_state, eos_check_compare, push
    o, ctx = make_pair(ol.BENCHMARK, 32, 16, 10)
    o.phase("begin")
    N, nj, ni, nrhs = eos_check_state(o)
    push(o, ctx)
    ctx.call("rho_eos", nrhs); ctx.sync()
    print("EMU-EOS-OK", eos_check_compare(ctx.download, N, nj, ni))


if __name__ == "__main__":
    if sys.argv[1] == "eos":
        eos()
    elif sys.argv[1] == "restart":
        restart()
    elif sys.argv[1] == "tiles":
        tiles()
    elif sys.argv[1] == "driver":
        driver()
    else:
        main()
