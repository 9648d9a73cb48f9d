"""ctypes binding of oracle/liboracle.so (the CPU restatement) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None
_FAST = None

UPWELLING, BENCHMARK = 0, 1


def build():
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "oracle"), "all"])


def fast_lib_path():
    """The timed CPU-baseline build (bench.py only, never a checker): -O3, FMA contraction; built with -march=native on the box that
    runs the benchmark when a compiler is there (make native), else the x86-64-v3 build shipped with the snapshot."""
    native = os.path.join(ROOT, "oracle", "_native", "liboracle_fast.so")
    try:
        subprocess.run(["make", "-s", "-j8", "-C", os.path.join(ROOT, "oracle"), "native"], check=True, timeout=300,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except Exception:
        pass
    if os.path.exists(native):
        return native, "-O3 -march=native -ffp-contract=fast"
    path = os.path.join(ROOT, "oracle", "liboracle_fast.so")
    if not os.path.exists(path):
        build()
    return path, "-O3 -march=x86-64-v3 -ffp-contract=fast"


def _bind(path):
    if True:
        L = C.CDLL(path)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int] * 6
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_dt.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.orc_initial.argtypes = [C.c_void_p]
        L.orc_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.orc_step.argtypes = [C.c_void_p, C.c_int]
        L.orc_phase.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_field_size.restype = C.c_long
        L.orc_field_size.argtypes = [C.c_void_p, C.c_char_p]
        for f in (L.orc_get, L.orc_set, L.orc_get_vec):
            f.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        for f in (L.orc_get_dims, L.orc_get_stepping, L.orc_set_stepping, L.orc_get_scalars, L.orc_get_ksbl, L.orc_get_diag):
            f.argtypes = [C.c_void_p, C.c_void_p]
    return L


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = _bind(path)
    return _LIB


def fast_lib():
    """(library, compiler flags) of the timed CPU-baseline build."""
    global _FAST
    if _FAST is None:
        path, flags = fast_lib_path()
        _FAST = (_bind(path), flags)
    return _FAST


PHASES = ["begin", "set_massflux", "rho_eos", "diag", "bulk_flux", "set_vbc", "vmix", "omega", "wvelocity", "set_zeta",
          "pre_step3d", "prsgrd", "t3dmix2", "rhs3d_tile", "uv3dmix2", "step2d_loop", "set_depth", "step3d_uv",
          "omega2", "step3d_t", "end"]

FIELDS_2D = ["h", "f", "fomn", "pm", "pn", "om_r", "on_r", "om_u", "on_u", "om_v", "on_v", "om_p", "on_p", "pmon_r",
             "pnom_r", "pmon_u", "pnom_u", "pmon_v", "pnom_v", "pmon_p", "pnom_p", "omn", "dndx", "dmde", "lonr", "latr",
             "xr", "yr", "angler", "rdrag", "rdrag2", "visc2_r", "visc2_p", "hsbl", "Jwtype", "Zt_avg1", "DU_avg1",
             "DU_avg2", "DV_avg1", "DV_avg2", "rufrc", "rvfrc", "rhoA", "rhoS", "alpha", "beta", "sustr", "svstr",
             "bustr", "bvstr", "srflx", "Uwind", "Vwind", "Tair", "Pair", "Hair", "cloud", "rain", "lrflx", "lhflx",
             "shflx"]
FIELDS_ND = ["Hz", "z_r", "z_w", "Huon", "Hvom", "diff2", "Akv", "bvf", "Akt", "ghats", "zeta", "ubar", "vbar", "rzeta",
             "rubar", "rvbar", "rho", "pden", "W", "wvel", "u", "v", "ru", "rv", "t", "stflx", "btflx", "stflux", "btflux"]
ALL_FIELDS = FIELDS_2D + FIELDS_ND


class Oracle:
    def __init__(self, app, Lm=0, Mm=0, N=0, NtileI=1, NtileJ=1, dt=None, ndtfast=None, fast=False):
        self.L = fast_lib()[0] if fast else lib()
        self.h = self.L.orc_create(app, Lm, Mm, N, NtileI, NtileJ)
        if dt is not None:
            self.L.orc_set_dt(self.h, dt, ndtfast)
        self.app = app

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    def set_threads(self, n):
        self.L.orc_set_threads(self.h, n)

    def initial(self):
        self.L.orc_initial(self.h)

    def step(self, n=1):
        self.L.orc_step(self.h, n)

    def phase(self, name):
        if self.L.orc_phase(self.h, name.encode()):
            raise ValueError(name)

    def dims(self):
        a = (C.c_int * 11)()
        self.L.orc_get_dims(self.h, a)
        k = ["LBi", "UBi", "LBj", "UBj", "N", "NT", "NAT", "Lm", "Mm", "nfast", "ndtfast"]
        return dict(zip(k, list(a)))

    def stepping(self):
        a = (C.c_int * 11)()
        self.L.orc_get_stepping(self.h, a)
        k = ["iic", "ntfirst", "nstp", "nnew", "nrhs", "kstp", "knew", "krhs", "indx1", "iif", "predictor"]
        return dict(zip(k, list(a)))

    def scalars(self):
        a = (C.c_double * 8)()
        self.L.orc_get_scalars(self.h, a)
        k = ["dt", "dtfast", "hc", "time", "tdays", "avgke", "avgpe", "volume"]
        return dict(zip(k, list(a)))

    DIAG_KEYS = ["avgke", "avgpe", "volume", "max_C", "max_Cu", "max_Cv", "max_Cw", "max_Ci", "max_Cj", "max_Ck", "maxspeed",
                 "maxrho", "exit_flag"]

    def diag_full(self):
        """Everything diag.F reports after the last `diag` phase, in the order of roms_b200_diag_full."""
        a = np.zeros(13)
        self.L.orc_get_diag(self.h, a.ctypes.data)
        return a

    def vec(self, name):
        buf = np.zeros(1024)
        n = self.L.orc_get_vec(self.h, name.encode(), buf.ctypes.data)
        assert n >= 0, name
        return buf[:n].copy()

    def get(self, name):
        n = self.L.orc_field_size(self.h, name.encode())
        assert n >= 0, name
        out = np.empty(n)
        self.L.orc_get(self.h, name.encode(), out.ctypes.data)
        return out

    def set(self, name, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64).ravel()
        assert arr.size == self.L.orc_field_size(self.h, name.encode()), name
        self.L.orc_set(self.h, name.encode(), arr.ctypes.data)

    def shaped(self, name):
        """Field as a numpy array indexed [..., k, j - LBj, i - LBi]."""
        d = self.dims()
        ni, nj = d["UBi"] - d["LBi"] + 1, d["UBj"] - d["LBj"] + 1
        a = self.get(name)
        return a.reshape(-1, nj, ni)

    def snapshot(self, names=ALL_FIELDS):
        return {n: self.get(n) for n in names}
