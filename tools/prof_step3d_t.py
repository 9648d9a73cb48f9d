"""Run a few step3d_t launches on a grid >> L2 so ncu can capture the kernel (use under gpurun + ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import roms_b200 as rb

Lm, Mm, N = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (1024, 512, 50)))
cfg = rb.default_config(rb.APP_BENCHMARK, Lm, Mm, N)
cfg.dt, cfg.ndtfast = 20.0, 20
d = rb.Driver(cfg)
d.run(3)
st, _ = d.ctx.get_stepping()
for reps in (3, 10):
    ms = d.ctx.time_step3d_t(st["nrhs"], st["nstp"], st["nnew"], reps)
cells = Lm * Mm * N
print("step3d_t %dx%dx%d: %.4f ms/launch, %.1f GB/s algorithmic (96 B/cell), %.2f Gcell/s" % (Lm, Mm, N, ms, 96.0 * cells / ms / 1e6, cells / ms / 1e6))
d.finalize()
