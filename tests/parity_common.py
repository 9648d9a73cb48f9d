"""Shared plumbing of the parity tests: build an oracle Model and a roms_b200
Context with identical configuration, move state between them, compare fields."""
import ctypes as C

import numpy as np

import oracle_lib as ol
import roms_b200 as rb

# oracle phase -> (C-ABI entry point, argument builder from the oracle's stepping dict)
GPU_PHASE = {
    "set_massflux": ("set_massflux", lambda s: (s["nrhs"],)),
    "rho_eos": ("rho_eos", lambda s: (s["nrhs"],)),
    "bulk_flux": ("bulk_flux", lambda s: (s["nrhs"],)),
    "set_vbc": ("set_vbc", lambda s: (s["nrhs"],)),
    "omega": ("omega", lambda s: ()),
    "wvelocity": ("wvelocity", lambda s: (s["nstp"],)),
    "set_zeta": ("set_zeta", lambda s: ()),
    "pre_step3d": ("pre_step3d", lambda s: (s["nrhs"], s["nstp"], s["nnew"], s["iic"], s["ntfirst"])),
    "prsgrd": ("prsgrd", lambda s: (s["nrhs"],)),
    "t3dmix2": ("t3dmix2", lambda s: (s["nrhs"], s["nstp"], s["nnew"])),
    "rhs3d_tile": ("rhs3d_tile", lambda s: (s["nrhs"],)),
    "uv3dmix2": ("uv3dmix2", lambda s: (s["nrhs"], s["nnew"])),
    "set_depth": ("set_depth", lambda s: ()),
    "step3d_uv": ("step3d_uv", lambda s: (s["nrhs"], s["nstp"], s["nnew"], s["iic"], s["ntfirst"])),
    "omega2": ("omega", lambda s: ()),
    "step3d_t": ("step3d_t", lambda s: (s["nrhs"], s["nstp"], s["nnew"])),
}
# kernels that call exp/log/pow: device libm vs glibc differ by a few ulp
TRANSCENDENTAL = {"vmix", "bulk_flux"}
FORCING_FIELDS = ["srflx", "sustr", "svstr", "stflux", "btflux", "cloud", "Tair", "Hair", "Pair", "rain", "Uwind", "Vwind"]
PROGNOSTIC = ["zeta", "ubar", "vbar", "u", "v", "t"]


def prognostic_errors(get_ref, get_got):
    """Max-norm differences of the prognostic fields, relative.  Scalars (zeta, t) are scaled by the field's own range; the two
    components of a velocity vector ((ubar, vbar), (u, v)) by the larger of the two ranges -- a zonal flow has a meridional
    component whose own range is orders of magnitude smaller than the velocity scale the error lives on.  Returns
    ({field: error / scale}, {field: error / own range})."""
    ref = {n: get_ref(n) for n in PROGNOSTIC}
    rng = {n: float(ref[n].max() - ref[n].min()) for n in PROGNOSTIC}
    scale = dict(rng)
    for a, b in (("ubar", "vbar"), ("u", "v")):
        scale[a] = scale[b] = max(rng[a], rng[b])
    err = {n: float(np.max(np.abs(ref[n] - get_got(n)))) for n in PROGNOSTIC}
    return ({n: err[n] / max(scale[n], 1e-300) for n in PROGNOSTIC}, {n: err[n] / max(rng[n], 1e-300) for n in PROGNOSTIC})


def make_params(o):
    d, sc, c = o.dims(), o.scalars(), None
    p = rb.Params()
    p.app = o.app
    p.dt, p.dtfast, p.ndtfast, p.nfast = sc["dt"], sc["dtfast"], d["ndtfast"], d["nfast"]
    p.rho0, p.g, p.gamma2, p.hc = 1025.0, 9.81, 1.0, sc["hc"]
    if o.app == ol.UPWELLING:
        p.R0, p.T0, p.S0, p.Tcoef, p.Scoef = 1027.0, 14.0, 35.0, 1.7e-4, 0.0
        p.Akt_bak[0] = p.Akt_bak[1] = 1.0e-6
        p.Akv_bak = 1.0e-5
    else:
        p.R0, p.T0, p.S0, p.Tcoef, p.Scoef = 1027.0, 10.0, 35.0, 1.7e-4, 7.6e-4
        p.Akt_bak[0] = p.Akt_bak[1] = 1.0e-5
        p.Akv_bak = 1.0e-4
    p.blk_ZQ = p.blk_ZT = p.blk_ZW = 10.0
    p.dstart = 0.0
    return p


def make_pair(app, Lm=0, Mm=0, N=0, device=0, dt=None, ndtfast=None):
    """Oracle (1x1 tiling) + GPU context for the same configuration; oracle initialised."""
    o = ol.Oracle(app, Lm, Mm, N, dt=dt, ndtfast=ndtfast)
    o.initial()
    d = o.dims()
    b = rb.tile_bounds(d["Lm"], d["Mm"], d["N"], d["NT"], d["NAT"])
    assert (b.LBi, b.UBi, b.LBj, b.UBj) == (d["LBi"], d["UBi"], d["LBj"], d["UBj"])
    ctx = rb.Context(b, make_params(o), device)
    ctx.set_scoord(o.vec("sc_r"), o.vec("Cs_r"), o.vec("sc_w"), o.vec("Cs_w"))
    ctx.set_weights(d["nfast"], o.vec("weight1"), o.vec("weight2"))
    return o, ctx


def push(o, ctx, names=None):
    for n in (names or rb.FIELD_NAMES):
        ctx.upload(n, o.get(n))


def diff_fields(o, ctx, names=None):
    """max |gpu-oracle| and the scale max|oracle| per field."""
    out = {}
    for n in (names or rb.FIELD_NAMES):
        a, g = o.get(n), ctx.download(n)
        out[n] = (float(np.max(np.abs(a - g))) if a.size else 0.0, float(np.max(np.abs(a))) if a.size else 0.0,
                  bool(np.array_equal(a, g)))
    return out


def run_phase_gpu(o, ctx, ph):
    s = o.stepping()
    if ph == "vmix":
        if o.app == ol.BENCHMARK:
            ctx.call("lmd_vmix", s["nstp"])
        else:
            ctx.call("ana_vmix")
    elif ph == "step2d_loop":
        return ctx.step2d_loop(s["nstp"], s["nnew"], s["iic"], s["ntfirst"], s["indx1"])
    else:
        name, argf = GPU_PHASE[ph]
        ctx.call(name, *argf(s))
    ctx.sync()
    return None


# The one known-answer vector the reference holds for this path: the header of ROMS/Nonlinear/rho_eos.F:21-29,
# "Check Values: (T=3 C, S=35.5 PSU, Z=-5000 m)" of the Jackett & McDougall (1995) equation of state.
EOS_CHECK = {"den": 1050.3639165364, "den1": 1028.2845117925, "alpha": 2.1014611551470e-04, "beta": 7.2575037309946e-04}


def eos_check_state(o):
    """Put T=3, S=35.5, z_r=-5000 into the surface level of the oracle's state (alpha, beta are evaluated there,
    rho_eos.F:426-470) and return the indices of an interior point plus the array shapes."""
    d = o.dims()
    N, ni, nj = d["N"], d["UBi"] - d["LBi"] + 1, d["UBj"] - d["LBj"] + 1
    nrhs = o.stepping()["nrhs"]
    t = o.get("t").reshape(2, 3, N, nj, ni)
    t[0, nrhs - 1, N - 1] = 3.0
    t[1, nrhs - 1, N - 1] = 35.5
    o.set("t", t)
    zr = o.get("z_r").reshape(N, nj, ni)
    zr[N - 1] = -5000.0
    o.set("z_r", zr)
    return N, nj, ni, nrhs


def eos_check_compare(get, N, nj, ni):
    """`get(name)` returns a field after rho_eos ran on the state above; compare with the reference's printed digits."""
    j, i = nj // 2, ni // 2
    got = {"den": get("rho").reshape(N, nj, ni)[N - 1, j, i] + 1000.0, "den1": get("pden").reshape(N, nj, ni)[N - 1, j, i] + 1000.0,
           "alpha": get("alpha").reshape(nj, ni)[j, i], "beta": get("beta").reshape(nj, ni)[j, i]}
    for k, ref in EOS_CHECK.items():
        assert abs(got[k] / ref - 1.0) < 5e-14, (k, got[k], ref)      # the header prints 14 significant digits
    return got
