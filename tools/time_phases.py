"""Warm-cache CUDA-event timing of every kernel entry point of one baroclinic step (BENCHMARK option set).
Each phase is repeated `reps` times back to back on the state after a few steps (results are discarded:
this is a timing tool, not a model run).  usage: time_phases.py [Lm Mm N] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import roms_b200 as rb

a = sys.argv[1:]
Lm, Mm, N = (int(x) for x in a[:3]) if len(a) >= 3 else (512, 64, 30)
reps = int(a[3]) if len(a) > 3 else 20
cfg = rb.default_config(rb.APP_BENCHMARK, Lm, Mm, N)
d = rb.Driver(cfg)
d.run(4)
ctx = d.ctx
st, _ = ctx.get_stepping()
iic, ntf, nstp, nnew, nrhs, indx1 = (st[k] for k in ("iic", "ntfirst", "nstp", "nnew", "nrhs", "indx1"))
phases = [("set_massflux", (nrhs,)), ("rho_eos", (nrhs,)), ("bulk_flux", (nrhs,)), ("set_vbc", (nrhs,)), ("lmd_vmix", (nstp,)),
          ("omega", ()), ("wvelocity", (nstp,)), ("set_zeta", ()), ("pre_step3d", (nrhs, nstp, nnew, iic, ntf)), ("prsgrd", (nrhs,)),
          ("t3dmix2", (nrhs, nstp, nnew)), ("rhs3d_tile", (nrhs,)), ("uv3dmix2", (nrhs, nnew)),
          ("set_depth", ()), ("step3d_uv", (nrhs, nstp, nnew, iic, ntf)), ("step3d_t", (nrhs, nstp, nnew))]
# algorithmic doubles per interior cell per call: every 3-D array read once and written once (column "min")
ALG = {"set_massflux": 5, "rho_eos": 8, "lmd_vmix": 14, "omega": 4, "wvelocity": 6, "pre_step3d": 27, "prsgrd": 6, "t3dmix2": 8,
       "rhs3d_tile": 10, "uv3dmix2": 8, "set_depth": 3, "step3d_uv": 12, "step3d_t": 12}
try:
    import json
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0
cells = Lm * Mm * N
tot = 0.0
print("%dx%dx%d, %d reps each (%s); algorithmic GB/s = doubles/cell x 8 B x cells / time, fraction of %.0f GB/s"
      % (Lm, Mm, N, reps, "warm L2" if cells < 4e6 else "3-D state >> L2", PEAK))
for name, args in phases:
    ctx.call(name, *args); ctx.sync()
    d.timer_start()
    for _ in range(reps):
        ctx.call(name, *args)
    ms = d.timer_stop() / reps
    tot += ms * (2 if name == "omega" else 1)
    if name in ALG:
        gbs = ALG[name] * 8.0 * cells / (ms * 1e-3) / 1e9
        print("  %-14s %9.1f us   %2d doubles/cell  %7.0f GB/s  %.2f" % (name, 1e3 * ms, ALG[name], gbs, gbs / PEAK))
    else:
        print("  %-14s %9.1f us" % (name, 1e3 * ms))
ctx.diag(nstp)
d.timer_start()
for _ in range(reps):
    ctx.main3d(0)            # no-op; keeps the call pattern
    ctx.L.roms_b200_diag_begin(ctx.h, nstp)
ms = d.timer_stop() / reps
tot += ms
print("  %-14s %8.1f us  (two kernels + asynchronous D2H)" % ("diag", 1e3 * ms))
ctx.step2d_loop(nstp, nnew, iic, ntf, indx1); ctx.sync()
d.timer_start()
for _ in range(reps):
    ctx.step2d_loop(nstp, nnew, iic, ntf, indx1)
ms = d.timer_stop() / reps
nsub = 2 * d.nfast + 1
print("  %-14s %8.1f us  (%d sub-steps, %.2f us each)" % ("step2d_loop", 1e3 * ms, nsub, 1e3 * ms / nsub))
print("  sum %.1f us" % (1e3 * (tot + ms)))
d.finalize()
