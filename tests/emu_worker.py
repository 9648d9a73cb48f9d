"""Child process of tests/test_emu.py -- TEST INFRASTRUCTURE ONLY.
Loads tests/emu/libroms_b200_emu.so (the kernel SOURCES of roms_b200/csrc built with g++, see tests/emu/include/cuda_runtime.h)
in place of the CUDA library and runs the per-kernel parity protocol of tests/test_gpu_parity.py against the oracle.
Runs in its own process because the switches it needs (no CUDA graph, no programmatic launch, column step3d_t) are read once
per process by the library.   usage: emu_worker.py APP Lm Mm N NSTEPS"""
import os
import sys

os.environ["ROMS_B200_NO_GRAPH"] = "1"
os.environ["ROMS_B200_NO_PDL"] = "1"
os.environ["ROMS_B200_STEP3D_T_V1"] = "1"
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


if __name__ == "__main__":
    main()
