"""CPU tests (no GPU): the oracle against the invariants that pin it, the host logic of the product
(index contract, S-coordinate, fast-time weights, tile neighbours, symbol table of the C ABI), and the
N>1 halo plan on a 2-process gloo group.  The reference ships no golden vectors: the
pins are the reference's own acceptance criterion -- results do not depend on the tiling -- plus
conservation laws and self-generated regression values (tests/golden/)."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as ol
import roms_b200 as rb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROG = ["zeta", "ubar", "vbar", "u", "v", "t"]


def run(app, tiles=(1, 1), steps=0, threads=1, **kw):
    o = ol.Oracle(app, NtileI=tiles[0], NtileJ=tiles[1], **kw)
    if threads > 1:
        o.set_threads(threads)
    o.initial()
    if steps:
        o.step(steps)
    return o


# ---------------------------------------------------------------- oracle pins
@pytest.mark.parametrize("app,kw,steps", [(ol.UPWELLING, {}, 12), (ol.BENCHMARK, dict(Lm=48, Mm=24, N=30), 6)])
def test_oracle_tiling_invariance(app, kw, steps):
    """ROMS/Bin/verify.sh + check_nc.sh:35-43: 1x1, 2x2, 3x3 tilings give bit-identical output."""
    a = run(app, (1, 1), steps, **kw)
    others = []
    for tiles, th in (((2, 2), 1), ((3, 3), 1), ((4, 2), 4)):
        b = run(app, tiles, steps, threads=th, **kw)
        others.append((tiles, b))
        for n in PROG + ["Huon", "Hvom", "W", "wvel", "ru", "rv", "Akv", "Akt", "Zt_avg1", "DU_avg2", "rufrc"]:
            assert np.array_equal(a.get(n), b.get(n)), (tiles, n)
    # diag of the next step: sums differ by round-off between tilings (mp_reduce order); maxima and the MAXLOC location do not
    for o in [a] + [b for _, b in others]:
        for ph in ("begin", "set_massflux", "rho_eos", "diag"):
            o.phase(ph)
    da = a.diag_full()
    for tiles, b in others:
        db = b.diag_full()
        np.testing.assert_allclose(da[:3], db[:3], rtol=1e-13)
        assert np.array_equal(da[3:], db[3:]), (tiles, da, db)


def test_oracle_fast_time_filter():
    """set_weights.F: nfast = 42 for NDTFAST=30 and 29 for NDTFAST=20; weights normalised."""
    for app, kw, nf in ((ol.UPWELLING, {}, 42), (ol.BENCHMARK, dict(Lm=32, Mm=16, N=30), 29)):
        o = run(app, **kw)
        d = o.dims()
        assert d["nfast"] == nf
        w1, w2 = o.vec("weight1"), o.vec("weight2")
        assert abs(w1[1:nf + 1].sum() - 1.0) < 1e-14 and abs(w2[1:nf + 1].sum() - 1.0) < 1e-14
        # first moment of the primary weights = ndtfast (centred on the new baroclinic time)
        assert abs((w1[1:nf + 1] * np.arange(1, nf + 1)).sum() / d["ndtfast"] - 1.0) < 1e-12


def test_oracle_conservation_benchmark():
    """Closed/periodic channel: NET_VOLUME constant to round-off; uniform S=35 stays 35 (the artificial
    continuity term of pre_step3d.F:827-850 exists for exactly that); fluid stays bounded."""
    o = run(ol.BENCHMARK, Lm=64, Mm=32, N=30)
    o.step(1)
    v0 = o.scalars()["volume"]
    o.step(40)
    sc = o.scalars()
    assert abs(sc["volume"] / v0 - 1.0) < 1e-13
    d = o.dims()
    S = o.shaped("t").reshape(2, 3, d["N"], d["UBj"] + 1, -1)[1, :2]
    assert np.max(np.abs(S[:, :, 1:d["Mm"] + 1, 3:3 + d["Lm"]] - 35.0)) < 1e-11
    assert np.all(np.isfinite(o.get("u"))) and np.max(np.abs(o.get("u"))) < 2.0


def test_oracle_wvelocity_and_courant():
    """wvelocity.F: in a resting ocean w = 0; in the spun-up channel the surface value equals the free-surface tendency
    term (omega(N)=0) and the Courant search of diag.F returns a location inside the grid whose components add up."""
    o = run(ol.BENCHMARK, Lm=48, Mm=24, N=30)
    for ph in ol.PHASES[:9]:          # ... omega, wvelocity of the first step: fluid at rest
        o.phase(ph)
    assert np.max(np.abs(o.get("wvel"))) == 0.0
    for ph in ol.PHASES[9:]:
        o.phase(ph)
    o.step(5)
    for ph in ("begin", "set_massflux", "rho_eos", "diag"):
        o.phase(ph)
    d = dict(zip(ol.Oracle.DIAG_KEYS, o.diag_full()))
    dd = o.dims()
    assert d["max_C"] > 0 and abs(d["max_Cu"] + d["max_Cv"] + d["max_Cw"] - d["max_C"]) <= 1e-15 * d["max_C"] * 4
    assert 1 <= d["max_Ci"] <= dd["Lm"] and 1 <= d["max_Cj"] <= dd["Mm"] and 1 <= d["max_Ck"] <= dd["N"]
    assert 0 < d["maxspeed"] < 2.0 and 20.0 < d["maxrho"] < 60.0 and d["exit_flag"] == 0
    w = o.shaped("wvel")
    assert np.all(np.isfinite(w)) and 0 < np.max(np.abs(w)) < 1e-2
    # E-W periodic images and the closed-wall rows of bc_w3d
    Lm, Mm = dd["Lm"], dd["Mm"]
    assert np.array_equal(w[:, :, 2 + Lm + 1], w[:, :, 2 + 1]) and np.array_equal(w[:, 0, :], w[:, 1, :]) and np.array_equal(w[:, Mm + 1, :], w[:, Mm, :])


def test_oracle_eos_matches_the_reference_check_values():
    """The only known-answer vector the reference ships for this path (rho_eos.F:21-29): T=3, S=35.5, Z=-5000 m ->
    den, den1, alpha, beta to the 14 digits printed there.  Pins the nonlinear equation of state of the oracle."""
    from parity_common import eos_check_state, eos_check_compare
    o = run(ol.BENCHMARK, Lm=32, Mm=16, N=10)
    o.phase("begin")
    N, nj, ni, nrhs = eos_check_state(o)
    o.phase("rho_eos")
    eos_check_compare(o.get, N, nj, ni)


def test_oracle_regression_pins():
    """Self-generated regression values (NOT reference output): tests/golden/make_golden.py."""
    with open(os.path.join(ROOT, "tests", "golden", "oracle_pins.json")) as f:
        pins = json.load(f)
    for case in pins["cases"]:
        o = run(case["app"], Lm=case["Lm"], Mm=case["Mm"], N=case["N"])
        o.step(case["steps"])
        for ph in ("begin", "set_massflux", "rho_eos", "diag"):
            o.phase(ph)
        sc = o.scalars()
        np.testing.assert_allclose([sc["avgke"], sc["avgpe"], sc["volume"]], case["diag"], rtol=1e-11)
        for n, val in case["checksums"].items():
            np.testing.assert_allclose(float(np.sum(o.get(n) ** 2)), val, rtol=1e-10)


# ---------------------------------------------------------------- product host logic
TILE_KEYS = ("Istr Iend Jstr Jend IstrR IendR JstrR JendR IstrU JstrV IstrP IendP JstrP JendP IstrT IendT JstrT JendT IstrB IendB "
             "JstrB JendB IstrM JstrM Istrm3 Istrm2 Istrm1 IstrUm2 IstrUm1 Iendp1 Iendp2 Iendp2i Iendp3 Jstrm3 Jstrm2 Jstrm1 JstrVm2 "
             "JstrVm1 Jendp1 Jendp2 Jendp2i Jendp3 Western_Edge Eastern_Edge Southern_Edge Northern_Edge").split()


@pytest.mark.parametrize("Lm,Mm,ti,tj", [(41, 80, 1, 1), (512, 64, 2, 2), (2048, 256, 4, 2), (100, 37, 3, 3)])
def test_tile_bounds_match_oracle(Lm, Mm, ti, tj):
    """roms_b200_tile_bounds (product host code) vs the oracle's get_bounds restatement, every integer."""
    o = ol.Oracle(ol.BENCHMARK, Lm, Mm, 30, NtileI=ti, NtileJ=tj)
    L = ol.lib()
    L.orc_get_tile.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    for tile in range(ti * tj):
        buf = (C.c_int * len(TILE_KEYS))()
        L.orc_get_tile(o.h, tile, buf)
        b = rb.tile_bounds(Lm, Mm, 30, NtileI=ti, NtileJ=tj, tile=tile).asdict()
        assert [b[k] for k in TILE_KEYS] == list(buf), tile
    # serial allocation bounds (get_bounds.F:258-269, mod_param.F:1633-1636)
    d = o.dims()
    b = rb.tile_bounds(Lm, Mm, 30, NtileI=ti, NtileJ=tj, tile=0)
    assert (b.LBi, b.UBi, b.LBj, b.UBj) == (d["LBi"], d["UBi"], d["LBj"], d["UBj"])
    # distributed mirror: halo 3 around the tile, closed walls keep the global edge
    b = rb.tile_bounds(Lm, Mm, 30, NtileI=ti, NtileJ=tj, tile=ti * tj - 1, distributed=3)
    assert b.LBi == b.Istr - 3 and b.UBi == b.Iend + 3 and (b.LBj == (0 if tj == 1 else b.Jstr - 3))


def test_host_scoord_and_weights_match_oracle():
    L = rb.Lib.get().L
    L.roms_b200_host_scoord.argtypes = [C.c_int, C.c_double, C.c_double] + [C.c_void_p] * 4
    L.roms_b200_host_weights.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    for app, kw, th_s in ((ol.UPWELLING, {}, 3.0), (ol.BENCHMARK, dict(Lm=32, Mm=16, N=30), 0.0)):
        o = run(app, **kw)
        d = o.dims()
        a = [np.zeros(d["N"] + 1) for _ in range(4)]
        L.roms_b200_host_scoord(d["N"], th_s, 0.0, *[x.ctypes.data for x in a])
        for got, name in zip(a, ("sc_r", "Cs_r", "sc_w", "Cs_w")):
            assert np.array_equal(got, o.vec(name)), name
        w1, w2 = np.zeros(2 * d["ndtfast"] + 4), np.zeros(2 * d["ndtfast"] + 4)
        assert L.roms_b200_host_weights(d["ndtfast"], w1.ctypes.data, w2.ctypes.data) == d["nfast"]
        n = 2 * d["ndtfast"] + 2
        assert np.array_equal(w1[:n], o.vec("weight1")) and np.array_equal(w2[:n], o.vec("weight2"))


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "roms_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(roms_b200_\w+)\s*\(", hdr)))
    out = subprocess.run(["nm", "-D", "--defined-only", rb.library_path()], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (roms_b200_\w+)", out))
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    assert len(declared) > 45
    # the product never links the oracle
    ldd = subprocess.run(["ldd", rb.library_path()], capture_output=True, text=True).stdout
    assert "oracle" not in ldd


def test_fortran_shim_covers_the_kernel_entry_points():
    """roms_b200/fortran/roms_b200_mod.F90 (not compilable here: no Fortran compiler) must carry an ISO_C_BINDING interface for
    every entry point a Fortran ROMS calls; only the C++ driver surface and the bench/test helpers may be missing, and it must
    not name a symbol the header does not declare."""
    hdr = open(os.path.join(ROOT, "include", "roms_b200.h")).read()
    shim = open(os.path.join(ROOT, "roms_b200", "fortran", "roms_b200_mod.F90")).read()
    declared = set(re.findall(r"\b(roms_b200_\w+)\s*\(", hdr))
    bound = set(re.findall(r"name='(roms_b200_\w+)'", shim))
    host_only = {"roms_b200_ROMS_initialize", "roms_b200_ROMS_run", "roms_b200_ROMS_finalize", "roms_b200_default_config",
                 "roms_b200_driver_bounds", "roms_b200_driver_ctx", "roms_b200_driver_nfast", "roms_b200_host_scoord",
                 "roms_b200_host_weights", "roms_b200_tile_bounds", "roms_b200_tile_neighbors", "roms_b200_halo_plan",
                 "roms_b200_ana_initial", "roms_b200_ini_fields", "roms_b200_fill", "roms_b200_device_ptr", "roms_b200_flush_l2",
                 "roms_b200_launch_count", "roms_b200_time_step3d_t", "roms_b200_timer_start", "roms_b200_timer_stop"}
    assert bound <= declared, sorted(bound - declared)
    assert declared - bound <= host_only, sorted(declared - bound - host_only)


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly (no CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        rb.Context(rb.tile_bounds(32, 16, 8), rb.Params())
    with pytest.raises(RuntimeError):
        rb.Driver(rb.default_config(rb.APP_UPWELLING))


def test_tile_neighbours():
    L = rb.Lib.get().L
    L.roms_b200_tile_neighbors.argtypes = [C.POINTER(rb.Bounds), C.c_void_p]

    def nb(ti, tj, tile):
        b = rb.tile_bounds(512, 64, 30, NtileI=ti, NtileJ=tj, tile=tile, distributed=3)
        out = (C.c_int * 4)()
        assert L.roms_b200_tile_neighbors(C.byref(b), out) == 0
        return list(out)
    assert nb(1, 1, 0) == [-1, -1, -1, -1]
    assert nb(2, 1, 0) == [1, 1, -1, -1] and nb(2, 1, 1) == [0, 0, -1, -1]          # periodic pair
    assert nb(4, 2, 0) == [3, 1, -1, 4] and nb(4, 2, 7) == [6, 4, 3, -1] and nb(4, 2, 5) == [4, 6, 1, -1]
    assert nb(1, 2, 1) == [-1, -1, 0, -1]


def test_halo_plan_gloo_world2():
    """N>1 host logic on CPU: two gloo ranks run the two-phase (W/E then S/N, width 3, periodic E-W) exchange
    plan on numpy arrays with the strip index ranges of k_halo.cu and must reproduce the global field in
    every ghost cell, corners included."""
    script = os.path.join(ROOT, "tests", "halo_plan_worker.py")
    for tiles in ("2 1", "1 2"):
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                            "127.0.0.1", "--master-port", "29533", script] + tiles.split(), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert "HALO_OK" in r.stdout


def test_halo_plan_eight_neighbours():
    """The single-phase eight-neighbour plan of the NVLink mailbox transport (roms_b200_halo_plan), emulated on numpy for every
    tile of several tilings: after ONE exchange every ghost cell of every tile (corners included) must hold the global field,
    the block a tile sends towards d must have the shape of what the neighbour expects from the opposite side, and nothing
    outside the halo frame may be touched (mp_exchange semantics, Utility/mp_exchange.F:290-773, E-W periodic / N-S closed)."""
    L = rb.Lib.get().L
    L.roms_b200_halo_plan.argtypes = [C.POINTER(rb.Bounds), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    Lm, Mm = 48, 24
    opp = [1, 0, 3, 2, 7, 6, 5, 4]

    def G(i, j):                       # global analytic field, periodic in i
        return ((i - 1) % Lm + 1) + 1000.0 * j

    for nti, ntj, w in ((2, 1, 3), (1, 2, 3), (2, 2, 3), (4, 2, 3), (4, 1, 6), (3, 3, 3), (2, 2, 6), (4, 2, 6)):   # w=6: deep-halo mirror
        tiles = []
        for t in range(nti * ntj):
            b = rb.tile_bounds(Lm, Mm, 4, NtileI=nti, NtileJ=ntj, tile=t, distributed=w)
            r8, snd, rcv = (C.c_int * 8)(), (C.c_int * 32)(), (C.c_int * 32)()
            assert L.roms_b200_halo_plan(C.byref(b), w, r8, snd, rcv) == 0
            ni, nj = b.UBi - b.LBi + 1, b.UBj - b.LBj + 1
            A = np.full((nj, ni), np.nan)
            j0 = b.Jstr - (1 if b.Southern_Edge else 0)
            j1 = b.Jend + (1 if b.Northern_Edge else 0)
            for j in range(j0, j1 + 1):                    # interior + physical wall rows are "computed" locally
                for i in range(b.Istr, b.Iend + 1):
                    A[j - b.LBj, i - b.LBi] = G(i, j)
            if nti == 1:                                   # single tile in i: periodic images are local (kernels' st())
                for i in list(range(b.LBi, b.Istr)) + list(range(b.Iend + 1, b.UBi + 1)):
                    A[:, i - b.LBi] = A[:, ((i - 1) % Lm + 1) - b.LBi]
            tiles.append((b, list(r8), np.array(snd).reshape(8, 4), np.array(rcv).reshape(8, 4), A))
        new = [t[4].copy() for t in tiles]
        for t, (b, r8, snd, rcv, A) in enumerate(tiles):
            for d in range(8):
                if r8[d] < 0:
                    continue
                nbq = tiles[r8[d]]
                assert nbq[1][opp[d]] == t, (nti, ntj, t, d)                  # neighbour relation is symmetric
                s, r = snd[d], nbq[3][opp[d]]
                assert (s[1] - s[0], s[3] - s[2]) == (r[1] - r[0], r[3] - r[2]), (nti, ntj, t, d)
                blk = A[s[2] - b.LBj:s[3] - b.LBj + 1, s[0] - b.LBi:s[1] - b.LBi + 1]
                nb_b = nbq[0]
                new[r8[d]][r[2] - nb_b.LBj:r[3] - nb_b.LBj + 1, r[0] - nb_b.LBi:r[1] - nb_b.LBi + 1] = blk
        for t, (b, r8, snd, rcv, A) in enumerate(tiles):
            An = new[t]
            for j in range(max(b.LBj, 0), min(b.UBj, Mm + 1) + 1):           # every physical row of the mirror, ghosts included
                for i in range(b.Istr - w, b.Iend + w + 1):
                    assert An[j - b.LBj, i - b.LBi] == G(i, j), (nti, ntj, t, i, j, An[j - b.LBj, i - b.LBi])
