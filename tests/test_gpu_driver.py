"""GPU tests of the host driver surface (ROMS_initialize / ROMS_run / ROMS_finalize mirror) and a
negative control proving the parity harness detects a wrong kernel result."""
import numpy as np
import pytest

import oracle_lib as ol
import roms_b200 as rb
from parity_common import make_pair, push, diff_fields, PROGNOSTIC

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("app,Lm,Mm,N", [(ol.UPWELLING, 0, 0, 0), (ol.BENCHMARK, 96, 40, 30)])
def test_roms_initialize_matches_oracle_start_state(app, Lm, Mm, N):
    """Product start-up (host 2-D grid + device 3-D initial state) vs oracle initial()+first set_data/post_initial.
    exp/tanh/cos on the device differ from glibc by ulps -> 1e-13 relative."""
    o = ol.Oracle(app, Lm, Mm, N)
    o.initial()
    o.phase("begin")
    d = rb.Driver(rb.default_config(app, Lm, Mm, N))
    assert d.nfast == o.dims()["nfast"]
    bad = []
    skip = {"xr", "yr", "lonr", "latr"}        # set only on Istr-1..Iend+1 by the reference; product fills the whole row
    for n in rb.FIELD_NAMES:
        a, g = o.get(n), d.ctx.download(n)
        if n in skip:
            continue
        scale = float(np.max(np.abs(a)))
        # t is exp/tanh of z_r (ulp-level libm differences); bvf differentiates density -> amplified
        tol = 1e-9 if n in ("bvf",) else 1e-13
        if float(np.max(np.abs(a - g))) > tol * max(scale, 1e-300):
            bad.append((n, float(np.max(np.abs(a - g))), scale))
    assert not bad, bad
    # and 10 steps from there stay within 1e-10 of the oracle on the prognostic fields
    for ph in ol.PHASES[1:]:
        o.phase(ph)
    o.step(9)
    d.run(10)
    for n in PROGNOSTIC:
        a, g = o.get(n), d.ctx.download(n)
        rel = float(np.max(np.abs(a - g))) / max(float(a.max() - a.min()), 1e-300)
        assert rel <= 1e-10, (n, rel)
    # host-forcing path (per-step H2D of set_data's fields + D2H diag) gives the same answer as device forcing
    # (diag runs inside the step at the reference's place, after rho_eos: the values returned are those of the last step's
    # start state, the line the reference prints for that step)
    diag = d.run(2, host_forcing=True)
    o.step(1)
    for ph in ol.PHASES[:4]:
        o.phase(ph)
    ref = o.diag_full()
    np.testing.assert_allclose(diag, ref[:3], rtol=1e-11)
    full = d.ctx.diag_last()
    np.testing.assert_allclose(full[3:7], ref[3:7], rtol=1e-8)          # Courant numbers
    np.testing.assert_array_equal(full[7:10], ref[7:10])                # and where
    np.testing.assert_allclose(full[10:12], ref[10:12], rtol=1e-9)
    assert full[12] == 0.0
    for ph in ol.PHASES[4:]:
        o.phase(ph)
    # device-resident run: same diag, launched inside every step, read once at the end
    diag2 = d.run(1)
    for ph in ol.PHASES[:4]:
        o.phase(ph)
    np.testing.assert_allclose(diag2, o.diag_full()[:3], rtol=1e-11)
    d.finalize()


def test_output_snapshot_does_not_disturb_the_time_loop():
    """roms_b200_snapshot_begin/end (the output path): the snapshot holds the state of the instant it was requested although the
    loop keeps stepping while it drains, and the loop's results are those of a run without snapshots."""
    cfg = rb.default_config(ol.BENCHMARK, 96, 40, 30)
    names = ["zeta", "ubar", "vbar", "u", "v", "t"]
    d = rb.Driver(cfg)
    d.run(3)
    ref = {n: d.ctx.download(n) for n in names}
    views = d.ctx.snapshot_begin(names)
    with pytest.raises(RuntimeError):
        d.ctx.snapshot_begin(names)            # one snapshot in flight at a time
    d.run(2)
    d.ctx.snapshot_end()
    for n in names:
        assert np.array_equal(views[n], ref[n]), n
    d2 = rb.Driver(cfg)
    d2.run(3)
    d2.run(2)
    for n in names:
        assert np.array_equal(d.ctx.download(n), d2.ctx.download(n)), n
    views = d.ctx.snapshot_begin(["zeta"])     # buffers are reused, a smaller request works after a larger one
    d.ctx.snapshot_end()
    assert np.array_equal(views["zeta"], d.ctx.download("zeta"))
    d.finalize(); d2.finalize()


@pytest.mark.parametrize("app,grid", [(ol.BENCHMARK, (96, 40, 30)), (ol.UPWELLING, (0, 0, 0))])
def test_perfect_restart_through_the_snapshot_path(app, grid):
    """The PERFECT_RESTART record (roms_b200_restart_fields: the reference's list, Utility/wrt_rst.F:178-900) taken with the
    asynchronous snapshot API while the loop keeps stepping, read into a fresh context: the continuation is bit-identical to the
    uninterrupted run (zeta, ubar, vbar, u, v, t, ru, rv, Akv, Akt, W, rho, Zt_avg1, Hz)."""
    import ctypes as C
    L = rb.lib.Lib.get().L
    L.roms_b200_restart_fields.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.roms_b200_restart_finish.argtypes = [C.c_void_p]
    d = rb.Driver(rb.default_config(app, *grid))
    d.run(5)
    ids = (C.c_int * 32)()
    n = L.roms_b200_restart_fields(d.ctx.h, ids, 32)
    names = [rb.FIELD_NAMES[ids[q]] for q in range(n)]
    assert {"zeta", "rzeta", "ubar", "rubar", "vbar", "rvbar", "u", "ru", "v", "rv", "t", "rho", "Akv", "Akt"} <= set(names)
    views = d.ctx.snapshot_begin(names)
    st, tm = d.ctx.get_stepping()
    d.run(4)
    d.ctx.snapshot_end()
    snap = {nm: views[nm].copy() for nm in names}
    check = ("zeta", "ubar", "vbar", "u", "v", "t", "ru", "rv", "Akv", "Akt", "W", "rho", "Zt_avg1", "Hz")
    ref = {nm: d.ctx.download(nm) for nm in check}
    d2 = rb.Driver(rb.default_config(app, *grid))
    for nm in names:
        d2.ctx.upload(nm, snap[nm])
    d2.ctx.set_stepping(st["iic"], st["ntfirst"], st["nstp"], st["nnew"], st["nrhs"], st["indx1"], tm)
    assert L.roms_b200_restart_finish(d2.ctx.h) == 0
    d2.run(4)
    bad = [nm for nm in check if not np.array_equal(ref[nm], d2.ctx.download(nm))]
    assert not bad, bad
    d.finalize(); d2.finalize()


def test_two_stream_step_is_bit_identical_to_the_serial_order():
    """roms_b200_main3d runs independent branches of a step on a second stream (tracer branch beside the momentum branch and the
    fast loop; mass fluxes / omega beside density / surface fluxes / KPP).  A missing dependency would show up as a difference
    from the serial launch order (ROMS_B200_ONE_STREAM=1, read once per process -> child process), also on a grid large enough
    for kernels of both streams to be resident together."""
    import os
    import subprocess
    import sys
    import tempfile
    code = ("import sys, numpy as np; sys.path.insert(0, %r)\n"
            "import roms_b200 as rb\n"
            "for app, g in ((rb.APP_BENCHMARK, (512, 64, 30)), (rb.APP_UPWELLING, (0, 0, 0))):\n"
            "    d = rb.Driver(rb.default_config(app, *g))\n"
            "    d.run(12); d.run(3, host_forcing=True); d.ctx.sync()\n"
            "    np.savez(sys.argv[1] + '_%%d.npz' %% app, **{n: d.ctx.download(n) for n in ('zeta', 'ubar', 'vbar', 'u', 'v', 't', 'W', 'Akv', 'wvel')})\n"
            "    d.finalize()\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as tmp:
        for tag, env in (("two", {}), ("one", {"ROMS_B200_ONE_STREAM": "1"})):
            r = subprocess.run([sys.executable, "-c", code, os.path.join(tmp, tag)], capture_output=True, text=True, timeout=600,
                               env=dict(os.environ, **env))
            assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
        for app in (rb.APP_BENCHMARK, rb.APP_UPWELLING):
            a, b = np.load(os.path.join(tmp, "two_%d.npz" % app)), np.load(os.path.join(tmp, "one_%d.npz" % app))
            bad = [n for n in a.files if not np.array_equal(a[n], b[n])]
            assert not bad, (app, bad)


def test_negative_control_detects_missing_kernel():
    """If a kernel is NOT run the comparison must fail: guards against a vacuous harness."""
    o, ctx = make_pair(ol.UPWELLING)
    o.step(2)
    for ph in ol.PHASES[:18]:
        o.phase(ph)
    push(o, ctx)
    o.phase("step3d_t")            # oracle advances, GPU does not
    res = diff_fields(o, ctx, ["t"])
    assert not res["t"][2] and res["t"][0] > 0.0
    s = o.stepping()
    # now run it on pre-kernel state: re-push is impossible (oracle moved on), so rebuild
    o2, ctx2 = make_pair(ol.UPWELLING)
    o2.step(2)
    for ph in ol.PHASES[:18]:
        o2.phase(ph)
    push(o2, ctx2)
    ctx2.call("step3d_t", s["nrhs"], s["nstp"], s["nnew"]); ctx2.sync()
    o2.phase("step3d_t")
    assert diff_fields(o2, ctx2, ["t"])["t"][2]
    ctx.close(); ctx2.close()
