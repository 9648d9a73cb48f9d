"""The boundary as the reference's Fortran MPI host would use it, on the CPU emulation of the kernel sources (tests/emu): host
arrays with the reference's own tile bounds and NghostPoints = 2, registered and moved through roms_b200_*_bounds, the time step
driven routine by routine through the `_tile` entry points with the reference's argument lists (include/roms_b200.h), several
tiles as host threads -- bit-identical to the oracle on everything the reference keeps current in those arrays.  See
tests/fortran_host_worker.py for what the stand-in host does."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu_lib():
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(HERE, "emu")])
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(os.path.dirname(HERE), "oracle"), "liboracle.so"])


# app, Lm, Mm, N, steps, NtileI, NtileJ: BENCHMARK option set on 2x2 tiles (all cpp-optional arguments present: dndx/dmde, rhoA/rhoS,
# srflx, ghats), UPWELLING on 2x1 tiles (those arguments absent -> null pointers), three steps = all AB start-up forms
@pytest.mark.parametrize("app,Lm,Mm,N,steps,nti,ntj", [(1, 48, 24, 10, 2, 2, 2), (0, 40, 24, 8, 3, 2, 1)])
def test_fortran_mpi_host_drives_the_tile_entry_points(emu_lib, app, Lm, Mm, N, steps, nti, ntj):
    r = subprocess.run([sys.executable, os.path.join(HERE, "fortran_host_worker.py")] + [str(x) for x in (app, Lm, Mm, N, steps, nti, ntj)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "FORTRAN-HOST-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
