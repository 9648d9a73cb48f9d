"""ctypes binding of libroms_b200.so (include/roms_b200.h).

Fails loudly if the CUDA library is missing: there is no CPU fallback.
"""
import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
APP_UPWELLING, APP_BENCHMARK = 0, 1


def library_path():
    return os.path.join(_HERE, "libroms_b200.so")


def _field_names():
    hdr = open(os.path.join(ROOT, "include", "roms_b200.h")).read()
    block = hdr[hdr.index("#define ROMS_B200_FIELDS(X)"):hdr.index("enum roms_b200_field")]
    return re.findall(r"X\((\w+),", block)


FIELD_NAMES = _field_names()

_BOUNDS_INTS = ("Lm Mm N NT NAT LBi UBi LBj UBj Istr Iend Jstr Jend IstrR IendR JstrR JendR IstrU JstrV "
                "IstrP IendP JstrP JendP IstrT IendT JstrT JendT IstrB IendB JstrB JendB IstrM JstrM "
                "Istrm3 Istrm2 Istrm1 IstrUm2 IstrUm1 Iendp1 Iendp2 Iendp2i Iendp3 "
                "Jstrm3 Jstrm2 Jstrm1 JstrVm2 JstrVm1 Jendp1 Jendp2 Jendp2i Jendp3 "
                "Western_Edge Eastern_Edge Southern_Edge Northern_Edge EWperiodic NSperiodic "
                "NtileI NtileJ Itile Jtile").split()


class Bounds(C.Structure):
    _fields_ = [(n, C.c_int) for n in _BOUNDS_INTS]

    def asdict(self):
        return {n: getattr(self, n) for n in _BOUNDS_INTS}


class Params(C.Structure):
    _fields_ = [("app", C.c_int), ("dt", C.c_double), ("dtfast", C.c_double), ("ndtfast", C.c_int), ("nfast", C.c_int),
                ("rho0", C.c_double), ("g", C.c_double), ("gamma2", C.c_double), ("hc", C.c_double),
                ("R0", C.c_double), ("T0", C.c_double), ("S0", C.c_double), ("Tcoef", C.c_double), ("Scoef", C.c_double),
                ("Akt_bak", C.c_double * 2), ("Akv_bak", C.c_double),
                ("blk_ZQ", C.c_double), ("blk_ZT", C.c_double), ("blk_ZW", C.c_double), ("dstart", C.c_double)]


class Lib:
    """Loaded libroms_b200.so with typed prototypes."""
    _inst = None

    def __init__(self, path=None):
        # `path` is for tests/emu only (a host build of the kernel sources used to check arithmetic without a GPU);
        # the product always loads libroms_b200.so and has no CPU path
        path = path or library_path()
        if not os.path.exists(path):
            raise RuntimeError("roms_b200: %s not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)" % path)
        L = self.L = C.CDLL(path)
        vp, ci, cd = C.c_void_p, C.c_int, C.c_double
        L.roms_b200_tile_bounds.argtypes = [ci] * 11 + [C.POINTER(Bounds)]
        L.roms_b200_create.argtypes = [C.POINTER(Bounds), C.POINTER(Params), ci, C.POINTER(vp)]
        L.roms_b200_destroy.argtypes = [vp]
        L.roms_b200_set_scoord.argtypes = [vp] + [vp] * 4
        L.roms_b200_set_weights.argtypes = [vp, ci, vp, vp]
        L.roms_b200_field_id.argtypes = [C.c_char_p]
        L.roms_b200_field_size.argtypes = [vp, ci]
        L.roms_b200_field_size.restype = C.c_long
        L.roms_b200_upload.argtypes = [vp, ci, vp]
        L.roms_b200_download.argtypes = [vp, ci, vp]
        L.roms_b200_device_ptr.argtypes = [vp, ci]
        L.roms_b200_device_ptr.restype = vp
        L.roms_b200_sync.argtypes = [vp]
        L.roms_b200_launch_count.argtypes = [vp]
        L.roms_b200_launch_count.restype = C.c_long
        sig = {
            "set_massflux": [ci], "rho_eos": [ci], "omega": [], "wvelocity": [ci], "set_zeta": [], "set_depth": [], "bulk_flux": [ci],
            "set_vbc": [ci], "ana_vmix": [], "lmd_vmix": [ci], "pre_step3d": [ci] * 5, "prsgrd": [ci], "t3dmix2": [ci] * 3,
            "rhs3d_tile": [ci], "uv3dmix2": [ci] * 2, "rhs3d": [ci] * 5, "step2d": [ci] * 9, "step3d_uv": [ci] * 5,
            "step3d_t": [ci] * 3, "set_data": [cd],
        }
        for name, args in sig.items():
            getattr(L, "roms_b200_" + name).argtypes = [vp] + args
        L.roms_b200_diag.argtypes = [vp, ci, vp]
        L.roms_b200_diag_full.argtypes = [vp, ci, vp]
        L.roms_b200_diag_last.argtypes = [vp, vp]
        L.roms_b200_diag_begin.argtypes = [vp, ci]
        L.roms_b200_diag_end.argtypes = [vp, vp]
        L.roms_b200_step2d_loop.argtypes = [vp, ci, ci, ci, ci, C.POINTER(ci)]
        L.roms_b200_main3d.argtypes = [vp, ci, ci, ci]
        L.roms_b200_get_stepping.argtypes = [vp, vp, C.POINTER(cd)]
        L.roms_b200_set_stepping.argtypes = [vp, vp, cd]
        L.roms_b200_time_step3d_t.argtypes = [vp, ci, ci, ci, ci, C.POINTER(C.c_float)]

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = Lib()
        return cls._inst


def tile_bounds(Lm, Mm, N, NT=2, NAT=2, NtileI=1, NtileJ=1, tile=0, EWperiodic=1, NSperiodic=0, distributed=0):
    b = Bounds()
    rc = Lib.get().L.roms_b200_tile_bounds(Lm, Mm, N, NT, NAT, NtileI, NtileJ, tile, EWperiodic, NSperiodic, distributed,
                                           C.byref(b))
    if rc:
        raise ValueError("roms_b200_tile_bounds failed")
    return b


class Context:
    """Device mirror + kernel entry points for one tile on one GPU."""

    def __init__(self, bounds, params, device=0):
        self.lib = Lib.get()
        self.L = self.lib.L
        self.bounds, self.params = bounds, params
        h = C.c_void_p()
        rc = self.L.roms_b200_create(C.byref(bounds), C.byref(params), device, C.byref(h))
        if rc:
            raise RuntimeError("roms_b200_create failed (rc=%d): a CUDA device is required" % rc)
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.roms_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _chk(self, rc, what):
        if rc:
            raise RuntimeError("roms_b200_%s failed rc=%d" % (what, rc))

    def fid(self, name):
        f = self.L.roms_b200_field_id(name.encode())
        if f < 0:
            raise KeyError(name)
        return f

    def size(self, name):
        return self.L.roms_b200_field_size(self.h, self.fid(name))

    def upload(self, name, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64).ravel()
        assert a.size == self.size(name), (name, a.size, self.size(name))
        self._chk(self.L.roms_b200_upload(self.h, self.fid(name), a.ctypes.data), "upload")

    def download(self, name, out=None):
        if out is None:
            out = np.empty(self.size(name))
        self._chk(self.L.roms_b200_download(self.h, self.fid(name), out.ctypes.data), "download")
        return out

    def set_scoord(self, sc_r, Cs_r, sc_w, Cs_w):
        """sc_r, Cs_r: N values (levels 1:N) -- vectors indexed by level with an unused element 0 are accepted too; sc_w, Cs_w:
        N+1 values (levels 0:N), as SCALARS(ng) holds them (mod_scalars.F:1950-1968)."""
        N = self.bounds.N
        sc_r, Cs_r = (np.asarray(a, dtype=np.float64)[-N:] for a in (sc_r, Cs_r))
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (sc_r, Cs_r, sc_w, Cs_w)]
        assert arrs[0].size == N and arrs[1].size == N and arrs[2].size == N + 1 and arrs[3].size == N + 1
        self._chk(self.L.roms_b200_set_scoord(self.h, *[a.ctypes.data for a in arrs]), "set_scoord")

    def set_weights(self, nfast, w1, w2):
        w1 = np.ascontiguousarray(w1, dtype=np.float64)
        w2 = np.ascontiguousarray(w2, dtype=np.float64)
        self._chk(self.L.roms_b200_set_weights(self.h, nfast, w1.ctypes.data, w2.ctypes.data), "set_weights")

    def call(self, name, *args):
        self._chk(getattr(self.L, "roms_b200_" + name)(self.h, *args), name)

    def sync(self):
        self._chk(self.L.roms_b200_sync(self.h), "sync")

    def diag(self, nstp):
        out = np.zeros(3)
        self._chk(self.L.roms_b200_diag(self.h, nstp, out.ctypes.data), "diag")
        return out

    def diag_full(self, nstp):
        """avgke avgpe volume max_C max_Cu max_Cv max_Cw max_Ci max_Cj max_Ck maxspeed maxrho exit_flag (diag.F)."""
        out = np.zeros(13)
        self._chk(self.L.roms_b200_diag_full(self.h, nstp, out.ctypes.data), "diag_full")
        return out

    def diag_last(self):
        out = np.zeros(13)
        self._chk(self.L.roms_b200_diag_last(self.h, out.ctypes.data), "diag_last")
        return out

    def step2d_loop(self, nstp, nnew, iic, ntfirst, indx1):
        x = C.c_int(indx1)
        self._chk(self.L.roms_b200_step2d_loop(self.h, nstp, nnew, iic, ntfirst, C.byref(x)), "step2d_loop")
        return x.value

    def main3d(self, nsteps, analytic_forcing=1, with_diag=0):
        self._chk(self.L.roms_b200_main3d(self.h, nsteps, analytic_forcing, with_diag), "main3d")

    def get_stepping(self):
        a = (C.c_int * 6)()
        t = C.c_double()
        self.L.roms_b200_get_stepping(self.h, a, C.byref(t))
        return dict(zip(["iic", "ntfirst", "nstp", "nnew", "nrhs", "indx1"], list(a))), t.value

    def set_stepping(self, iic, ntfirst, nstp, nnew, nrhs, indx1, time):
        a = (C.c_int * 6)(iic, ntfirst, nstp, nnew, nrhs, indx1)
        self.L.roms_b200_set_stepping(self.h, a, time)

    def launches(self):
        return self.L.roms_b200_launch_count(self.h)

    def snapshot_begin(self, names):
        """Start an asynchronous snapshot of the named fields into pinned host buffers; returns {name: numpy view}.  The views
        are valid after snapshot_end()."""
        self.L.roms_b200_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
        self.L.roms_b200_snapshot_begin.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]
        self.L.roms_b200_snapshot_end.argtypes = [C.c_void_p]
        pool = self.__dict__.setdefault("_snap_pinned", {})
        ids = (C.c_int * len(names))(*[self.fid(n) for n in names])
        ptrs = (C.c_void_p * len(names))()
        views = {}
        for q, n in enumerate(names):
            if n not in pool:
                p = C.c_void_p()
                self._chk(self.L.roms_b200_host_alloc(self.size(n) * 8, C.byref(p)), "host_alloc")
                pool[n] = p
            ptrs[q] = pool[n]
            views[n] = np.ctypeslib.as_array(C.cast(pool[n], C.POINTER(C.c_double)), shape=(self.size(n),))
        self._chk(self.L.roms_b200_snapshot_begin(self.h, len(names), ids, ptrs), "snapshot_begin")
        return views

    def snapshot_end(self):
        self._chk(self.L.roms_b200_snapshot_end(self.h), "snapshot_end")

    def time_step3d_t(self, nrhs, nstp, nnew, reps):
        ms = C.c_float()
        self._chk(self.L.roms_b200_time_step3d_t(self.h, nrhs, nstp, nnew, reps, C.byref(ms)), "time_step3d_t")
        return ms.value


class Config(C.Structure):
    _fields_ = [("app", C.c_int), ("Lm", C.c_int), ("Mm", C.c_int), ("N", C.c_int), ("NT", C.c_int), ("NAT", C.c_int),
                ("NtileI", C.c_int), ("NtileJ", C.c_int), ("dt", C.c_double), ("ndtfast", C.c_int),
                ("theta_s", C.c_double), ("theta_b", C.c_double), ("Tcline", C.c_double),
                ("rho0", C.c_double), ("g", C.c_double), ("gamma2", C.c_double), ("rdrg", C.c_double), ("rdrg2", C.c_double),
                ("Akt_bak", C.c_double * 2), ("Akv_bak", C.c_double), ("tnu2", C.c_double * 2), ("visc2", C.c_double),
                ("R0", C.c_double), ("T0", C.c_double), ("S0", C.c_double), ("Tcoef", C.c_double), ("Scoef", C.c_double),
                ("blk_ZQ", C.c_double), ("blk_ZT", C.c_double), ("blk_ZW", C.c_double), ("lmd_Jwt", C.c_int)]


def default_config(app, Lm=0, Mm=0, N=0):
    L = Lib.get().L
    L.roms_b200_default_config.argtypes = [C.c_int] * 4 + [C.POINTER(Config)]
    L.roms_b200_default_config.restype = None
    c = Config()
    L.roms_b200_default_config(app, Lm, Mm, N, C.byref(c))
    return c


class Driver:
    """Mirror of the reference driver surface (Drivers/nl_roms.h): ROMS_initialize / ROMS_run / ROMS_finalize."""

    def __init__(self, cfg, tile=0, distributed=0, device=0):
        self.L = L = Lib.get().L
        L.roms_b200_ROMS_initialize.argtypes = [C.POINTER(Config), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.roms_b200_ROMS_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.roms_b200_ROMS_finalize.argtypes = [C.c_void_p]
        L.roms_b200_driver_ctx.argtypes = [C.c_void_p]
        L.roms_b200_driver_ctx.restype = C.c_void_p
        L.roms_b200_driver_nfast.argtypes = [C.c_void_p]
        L.roms_b200_timer_start.argtypes = [C.c_void_p]
        L.roms_b200_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.roms_b200_flush_l2.argtypes = [C.c_void_p, C.c_int]
        self.cfg = cfg
        d = C.c_void_p()
        rc = L.roms_b200_ROMS_initialize(C.byref(cfg), tile, distributed, device, C.byref(d))
        if rc:
            raise RuntimeError("ROMS_initialize failed rc=%d (a CUDA device is required; no CPU fallback)" % rc)
        self.d = d
        # borrow the context for field access without owning it
        self.ctx = Context.__new__(Context)
        self.ctx.lib = Lib.get()
        self.ctx.L = L
        self.ctx.h = C.c_void_p(L.roms_b200_driver_ctx(d))
        self.ctx.close = lambda: None
        self.nfast = L.roms_b200_driver_nfast(d)

    def run(self, nsteps, host_forcing=False):
        diag = np.zeros(3)
        rc = self.L.roms_b200_ROMS_run(self.d, nsteps, 1 if host_forcing else 0, diag.ctypes.data)
        if rc:
            raise RuntimeError("ROMS_run failed rc=%d" % rc)
        return diag

    def timer_start(self):
        self.L.roms_b200_timer_start(self.ctx.h)

    def timer_stop(self):
        ms = C.c_float()
        self.L.roms_b200_timer_stop(self.ctx.h, C.byref(ms))
        return ms.value

    def flush_l2(self, mbytes=256):
        self.L.roms_b200_flush_l2(self.ctx.h, mbytes)

    def finalize(self):
        if getattr(self, "d", None):
            self.ctx.h = None
            self.L.roms_b200_ROMS_finalize(self.d)
            self.d = None

    def __del__(self):
        self.finalize()


def comm_unique_id():
    L = Lib.get().L
    buf = C.create_string_buffer(128)
    L.roms_b200_comm_unique_id.argtypes = [C.c_char_p]
    if L.roms_b200_comm_unique_id(buf):
        raise RuntimeError("ncclGetUniqueId failed")
    return buf.raw


def _driver_comm_init(self, rank, nranks, id128):
    L = self.L
    L.roms_b200_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
    rc = L.roms_b200_comm_init(self.ctx.h, rank, nranks, id128)
    if rc:
        raise RuntimeError("roms_b200_comm_init failed rc=%d" % rc)


def _driver_p2p_handle(self):
    L = self.L
    buf = C.create_string_buffer(64)
    L.roms_b200_p2p_handle.argtypes = [C.c_void_p, C.c_char_p]
    rc = L.roms_b200_p2p_handle(self.ctx.h, buf)
    if rc:
        raise RuntimeError("roms_b200_p2p_handle failed rc=%d" % rc)
    return buf.raw


def _driver_p2p_connect(self, handles, nranks):
    L = self.L
    L.roms_b200_p2p_connect.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    rc = L.roms_b200_p2p_connect(self.ctx.h, handles, nranks)
    if rc:
        raise RuntimeError("roms_b200_p2p_connect failed rc=%d" % rc)


def _driver_bounds(self):
    b = Bounds()
    self.L.roms_b200_driver_bounds.argtypes = [C.c_void_p, C.POINTER(Bounds)]
    self.L.roms_b200_driver_bounds(self.d, C.byref(b))
    return b


def field_shapes(N, NT=2, NAT=2):
    """name -> (kLB, nk, nl, nm) from the X-macro table of include/roms_b200.h."""
    hdr = open(os.path.join(ROOT, "include", "roms_b200.h")).read()
    block = hdr[hdr.index("#define ROMS_B200_FIELDS(X)"):hdr.index("enum roms_b200_field")]
    sym = {"N": N, "Np1": N + 1, "NT": NT, "NAT": NAT}
    out = {}
    for name, klb, nk, nl, nm in re.findall(r"X\((\w+),(-?\w+),(\w+),(\w+),(\w+)\)", block):
        out[name] = (int(klb),) + tuple(sym[x] if x in sym else int(x) for x in (nk, nl, nm))
    return out


def _ctx_download_interior(self, name, l=1, m=1, nk=None):
    """Interior of volume (l,m) of a 3-D field, or of plane l of an (i,j,level) field."""
    b = self._bounds
    klb, fnk, fnl, fnm = field_shapes(b.N, b.NT, b.NAT)[name]
    if fnl == 1 and fnm == 1 and fnk <= 3 and name in ("zeta", "ubar", "vbar", "rzeta", "rubar", "rvbar", "diff2", "stflx", "btflx", "stflux", "btflux"):
        plane0, nplanes = l - 1, 1
    else:
        plane0, nplanes = fnk * ((l - 1) + fnl * (m - 1)), fnk
    wi, wj = b.Iend - b.Istr + 1, b.Jend - b.Jstr + 1
    out = np.empty((nplanes, wj, wi))
    self.L.roms_b200_download_interior.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    self._chk(self.L.roms_b200_download_interior(self.h, self.fid(name), plane0, nplanes, out.ctypes.data), "download_interior")
    return out


Driver.comm_init = _driver_comm_init
Driver.p2p_handle = _driver_p2p_handle
Driver.p2p_connect = _driver_p2p_connect
Driver.bounds = _driver_bounds
Context.download_interior = _ctx_download_interior
