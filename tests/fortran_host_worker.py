"""Child process of tests/test_fortran_host.py -- TEST INFRASTRUCTURE ONLY.

Plays the part of the reference's Fortran MPI host against the C ABI (on the CPU emulation of the kernel sources, tests/emu):
  * NtileI x NtileJ ranks (host threads), one context (device mirror, halo 6) per rank;
  * every model array is a separate HOST array with the reference's own tile bounds (LBi:UBi,LBj:UBj) and NghostPoints = 2
    (Utility/get_bounds.F:193-212), column-major with i fastest -- what c_loc(OCEAN(ng)%t) etc. would hand over;
  * the arrays are registered once, uploaded through their bounds, the mirror's wider halo is refreshed by roms_b200_exchange_field;
  * the time step is driven routine by routine through the `_tile` entry points with the reference's argument lists
    (the calls a wrapper X(ng,tile) would make), in the order of main3d.F:216-1148; the physics without a `_tile` export in the
    hot-path list (rho_eos, set_vbc, bulk_flux, vertical mixing, wvelocity, t3dmix2, uv3dmix2) through their plain entry points;
  * after NSTEPS the prognostic arrays are downloaded into the host arrays and compared with the oracle (one tile, whole domain)
    bit for bit on every point the reference keeps current in such an array (interior + its two ghost points).
It also checks that a `_tile` call with an array that is not the registered one, or with wrong bounds, is refused.
usage: fortran_host_worker.py APP Lm Mm N NSTEPS NtileI NtileJ"""
import ctypes as C
import os
import sys
import threading

os.environ["ROMS_B200_NO_GRAPH"] = "1"
os.environ["ROMS_B200_NO_PDL"] = "1"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np  # noqa: E402
import oracle_lib as ol  # noqa: E402
import roms_b200 as rb  # noqa: E402

rb.lib.Lib._inst = rb.lib.Lib(path=os.path.join(HERE, "emu", os.environ.get("EMU_WORKER_LIB", "libroms_b200_emu.so")))
from parity_common import FORCING_FIELDS, PROGNOSTIC, make_params  # noqa: E402

ci, vp = C.c_int, C.c_void_p


class Rank:
    """One MPI rank of the reference: its tile, its host arrays, its device mirror."""

    def __init__(self, o, app, Lm, Mm, N, nti, ntj, tile):
        self.L = L = rb.lib.Lib.get().L
        self.tile, self.nti, self.ntj = tile, nti, ntj
        d = o.dims()
        self.g = (d["LBi"], d["UBi"], d["LBj"], d["UBj"])               # bounds of the oracle's whole-domain arrays
        self.b = rb.tile_bounds(Lm, Mm, N, d["NT"], d["NAT"], nti, ntj, tile, distributed=6)
        self.ctx = rb.Context(self.b, make_params(o))
        self.ctx.set_scoord(o.vec("sc_r"), o.vec("Cs_r"), o.vec("sc_w"), o.vec("Cs_w"))
        self.ctx.set_weights(d["nfast"], o.vec("weight1"), o.vec("weight2"))
        self.nfast, self.N = d["nfast"], N
        hb = (ci * 4)()
        L.roms_b200_mpi_array_bounds.argtypes = [ci] * 8 + [vp]
        assert L.roms_b200_mpi_array_bounds(Lm, Mm, nti, ntj, tile, 1, 0, 2, hb) == 0
        self.hb = tuple(hb)                                              # LBi, UBi, LBj, UBj of the HOST arrays (halo 2)
        self.host = {}
        for f in (L.roms_b200_register_field, L.roms_b200_upload_bounds, L.roms_b200_download_bounds):
            f.argtypes = [vp, ci, vp, ci, ci, ci, ci]
        for f in (L.roms_b200_upload_registered, L.roms_b200_download_registered, L.roms_b200_exchange_field):
            f.argtypes = [vp, ci]
        L.roms_b200_set_fast_step.argtypes = [vp, ci, ci]
        L.roms_b200_fast_loop_begin.argtypes = [vp]
        L.roms_b200_comm_init.argtypes = [vp, ci, ci, C.c_char_p]

    def tile_slice(self, glob):
        """The tile's (LBi:UBi,LBj:UBj) box of a whole-domain array (planes, j, i)."""
        gLBi, gUBi, gLBj, gUBj = self.g
        LBi, UBi, LBj, UBj = self.hb
        a = glob.reshape(-1, gUBj - gLBj + 1, gUBi - gLBi + 1)
        return np.ascontiguousarray(a[:, LBj - gLBj:UBj - gLBj + 1, LBi - gLBi:UBi - gLBi + 1])

    def allocate(self, o):
        for n in rb.FIELD_NAMES:                                         # ROMS_allocate_arrays + initial: host arrays, halo 2
            self.host[n] = self.tile_slice(o.get(n))
            assert self.L.roms_b200_register_field(self.ctx.h, self.ctx.fid(n), self.host[n].ctypes.data, *self.hb) == 0

    def upload(self, names):
        for n in names:
            assert self.L.roms_b200_upload_registered(self.ctx.h, self.ctx.fid(n)) == 0, n

    def exchange(self, names):
        for n in names:
            assert self.L.roms_b200_exchange_field(self.ctx.h, self.ctx.fid(n)) == 0, n

    def download(self, names):
        for n in names:
            assert self.L.roms_b200_download_registered(self.ctx.h, self.ctx.fid(n)) == 0, n

    # ---- the `_tile` calls, argument lists as in the reference (arrays by name = c_loc of the host array)
    def tile_call(self, routine, ints, arrays, model=None, ubk=None, expect=0):
        b = self.b
        head = [1, self.tile] + ([model] if model is not None else []) + list(self.hb) + ([ubk] if ubk is not None else []) + \
               [b.Istr - 3, b.Iend + 3, b.Jstr - 3, b.Jend + 3]
        ptrs = [None if a is None else (a if isinstance(a, int) else self.host[a].ctypes.data) for a in arrays]
        fn = getattr(self.L, "roms_b200_" + routine)
        fn.argtypes = [vp] + [ci] * (len(head) + len(ints)) + [vp] * len(ptrs)
        rc = fn(self.ctx.h, *head, *ints, *ptrs)
        assert (rc == 0) == (expect == 0), (routine, rc)
        return rc

    def step(self, s, bench):
        """One baroclinic step, main3d.F:216-1148, stepping indices from the dict `s` (the oracle's mod_stepping)."""
        c, nrhs, nstp, nnew = self.ctx, s["nrhs"], s["nstp"], s["nnew"]
        opt = (lambda n: n) if bench else (lambda n: None)               # cpp-optional arguments: absent in UPWELLING
        c.set_stepping(s["iic"], s["ntfirst"], nstp, nnew, nrhs, s["indx1"], 0.0)
        self.tile_call("set_massflux_tile", [nrhs], ["u", "v", "Hz", "om_v", "on_u", "Huon", "Hvom"], model=1)
        c.call("rho_eos", nrhs)
        if bench:
            c.call("bulk_flux", nrhs)
        c.call("set_vbc", nrhs)
        c.call("lmd_vmix", nstp) if bench else c.call("ana_vmix")
        self.tile_call("omega_tile", [], ["Huon", "Hvom", "z_w", "W"], model=1)
        c.call("wvelocity", nstp)
        self.tile_call("set_zeta_tile", [], ["Zt_avg1", "zeta"])
        self.tile_call("pre_step3d_tile", [nrhs, nstp, nnew],
                       ["pm", "pn", "Hz", "Huon", "Hvom", "z_r", "z_w", "btflx", "bustr", "bvstr", "stflx", "sustr", "svstr", opt("srflx"), "Akt",
                        "Akv", opt("ghats"), "W", "ru", "rv", "t", "u", "v"])
        self.tile_call("prsgrd32_tile", [nrhs], ["om_v", "on_u", "Hz", "z_r", "z_w", "rho", "ru", "rv"])
        c.call("t3dmix2", nrhs, nstp, nnew)
        self.tile_call("rhs3d_tile_tile", [nrhs],
                       ["Hz", "Huon", "Hvom", opt("dmde"), opt("dndx"), "fomn", "om_u", "om_v", "on_u", "on_v", "pm", "pn", "bustr", "bvstr", "sustr",
                        "svstr", "u", "v", "W", "rufrc", "rvfrc", "ru", "rv"])
        c.call("uv3dmix2", nrhs, nnew)
        # ---- main3d.F:810-918: LF-AM3 fast loop
        assert self.L.roms_b200_fast_loop_begin(c.h) == 0
        arrays2d = ["fomn", "h", "om_u", "om_v", "on_u", "on_v", "omn", "pm", "pn", opt("dndx"), opt("dmde"), "pmon_r", "pnom_r", "pmon_p", "pnom_p",
                    "om_r", "on_r", "om_p", "on_p", "visc2_p", "visc2_r", opt("rhoA"), opt("rhoS"), "DU_avg1", "DU_avg2", "DV_avg1", "DV_avg2",
                    "Zt_avg1", "rufrc", "rvfrc", "ru", "rv", "rubar", "rvbar", "rzeta", "ubar", "vbar", "zeta"]
        indx1, pred = s["indx1"], False
        kstp = knew = krhs = iif = 1
        for my_iif in range(1, self.nfast + 2):
            next_indx1 = 3 - indx1
            if not pred:
                pred, iif = True, my_iif
                kstp = indx1 if iif == 1 else 3 - indx1
                knew, krhs = 3, indx1
            self.L.roms_b200_set_fast_step(c.h, iif, 1)
            self.tile_call("step2d_tile", [krhs, kstp, knew, nstp, nnew], arrays2d, ubk=self.N)
            if pred:
                pred = False
                knew = next_indx1
                kstp, krhs = 3 - knew, 3
                if iif < self.nfast + 1:
                    indx1 = next_indx1
            if iif < self.nfast + 1:
                self.L.roms_b200_set_fast_step(c.h, iif, 0)
                self.tile_call("step2d_tile", [krhs, kstp, knew, nstp, nnew], arrays2d, ubk=self.N)
        self.tile_call("set_depth_tile", [nstp, nnew], ["h", "Zt_avg1", "Hz", "z_r", "z_w"], model=1)
        self.tile_call("step3d_uv_tile", [nrhs, nstp, nnew],
                       ["om_v", "on_u", "pm", "pn", "Hz", "z_r", "z_w", "Akv", "DU_avg1", "DV_avg1", "DU_avg2", "DV_avg2", "ru", "rv", "u", "v", "ubar",
                        "vbar", "Huon", "Hvom"])
        self.tile_call("omega_tile", [], ["Huon", "Hvom", "z_w", "W"], model=1)
        self.tile_call("step3d_t_tile", [nrhs, nstp, nnew],
                       ["omn", "om_u", "om_v", "on_u", "on_v", "pm", "pn", "Hz", "Huon", "Hvom", "z_r", "Akt", "W", "t"])
        c.sync()
        return indx1


def main():
    app, Lm, Mm, N, nsteps, nti, ntj = (int(x) for x in sys.argv[1:8])
    world = nti * ntj
    bench = (app == ol.BENCHMARK)
    o = ol.Oracle(app, Lm, Mm, N)
    o.initial()
    o.phase("begin")
    ranks = [Rank(o, app, Lm, Mm, N, nti, ntj, r) for r in range(world)]
    for r in ranks:
        assert r.L.roms_b200_comm_init(r.ctx.h, r.tile, world, bytes(128)) == 0
        r.allocate(o)
    errs = []

    def par(fn):
        def work(r):
            try:
                fn(r)
            except BaseException as e:      # noqa: BLE001
                errs.append((r.tile, repr(e)))
        th = [threading.Thread(target=work, args=(r,)) for r in ranks]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errs, errs

    par(lambda r: (r.upload(rb.FIELD_NAMES), r.exchange(rb.FIELD_NAMES)))
    for step in range(nsteps):
        if step > 0:
            o.phase("begin")                                            # set_data of this step on the "host": new forcing arrays
            for r in ranks:
                for n in FORCING_FIELDS:
                    r.host[n][...] = r.tile_slice(o.get(n))
            par(lambda r: (r.upload(FORCING_FIELDS), r.exchange(FORCING_FIELDS)))
        s = o.stepping()
        res = {}
        par(lambda r: res.__setitem__(r.tile, r.step(s, bench)))
        for ph in ol.PHASES[1:]:
            o.phase(ph)
        assert all(v == o.stepping()["indx1"] for v in res.values()), (res, o.stepping())
    # ---- output: download the prognostic arrays into the host arrays, compare every point they hold with the oracle
    par(lambda r: r.download(PROGNOSTIC))
    # (the reference keeps interior + NghostPoints = 2 ghost points current; the padding column Im+2 = Lm+3 of an even Lm,
    #  mod_param.F:1633-1636, which the host array of the eastern tile also holds, is never written by it)
    bad = []
    for r in ranks:
        LBi, UBi, LBj, UBj = r.hb
        i0, i1 = max(LBi, r.b.Istr - 2) - LBi, min(UBi, r.b.Iend + 2) - LBi
        j0, j1 = max(LBj, r.b.Jstr - 2) - LBj, min(UBj, r.b.Jend + 2) - LBj
        for n in PROGNOSTIC:
            ref, got = r.tile_slice(o.get(n))[:, j0:j1 + 1, i0:i1 + 1], r.host[n][:, j0:j1 + 1, i0:i1 + 1]
            if not np.array_equal(ref, got):
                bad.append((r.tile, n, float(np.max(np.abs(ref - got)))))
    assert not bad, bad
    # ---- error behaviour of the `_tile` entry points: a foreign array, wrong bounds, a wrong tile number are refused
    r = ranks[0]
    other = np.zeros_like(r.host["zeta"])
    assert r.tile_call("set_zeta_tile", [], ["Zt_avg1", other.ctypes.data], expect=1) != 0
    keep = r.hb
    r.hb = (keep[0] + 1,) + keep[1:]
    assert r.tile_call("set_zeta_tile", [], ["Zt_avg1", "zeta"], expect=1) != 0
    r.hb = keep
    r.tile += 1
    assert r.tile_call("set_zeta_tile", [], ["Zt_avg1", "zeta"], expect=1) != 0
    r.tile -= 1
    for r in ranks:
        r.ctx.close()
    print("FORTRAN-HOST-OK app=%d %dx%dx%d steps=%d tiles=%dx%d host bounds of tile 0: %s, mirror bounds: %s"
          % (app, Lm, Mm, N, nsteps, nti, ntj, ranks[0].hb, (ranks[0].b.LBi, ranks[0].b.UBi, ranks[0].b.LBj, ranks[0].b.UBj)))


if __name__ == "__main__":
    main()
