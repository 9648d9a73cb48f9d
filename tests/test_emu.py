"""CPU tests of the kernel SOURCES: roms_b200/csrc/*.cu built with g++ against a stand-in CUDA runtime (tests/emu/) and
run through the same per-kernel parity protocol as tests/test_gpu_parity.py -- push the oracle's state, run ONE kernel entry
point through the C ABI, compare every field of the mirror with the oracle.  This pins loop bounds, stencil indices and
operation order of a kernel before it ever reaches a GPU (the build container has none).  It says nothing about speed, about
races between blocks, or about the halo transport (k_halo.cu is not built; the multi-tile test replaces it by in-process
copies), which only the GPU runs cover.  The production step3d_t (k_step3d_t8.cu) IS built: its PTX helpers have host alternates.
The emulation library is test infrastructure: the product never loads it."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu_lib():
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(HERE, "emu")])
    subprocess.check_call(["make", "-s", "-C", os.path.join(os.path.dirname(HERE), "oracle")])
    return os.path.join(HERE, "emu", "libroms_b200_emu.so")


# app, Lm, Mm, N, steps, step3d_t kernel, emulated SM count: UPWELLING (linear EOS, ana_vmix, t3dmix2_s) on a small channel;
# BENCHMARK (UNESCO EOS, KPP, bulk fluxes, geopotential mixing, curvilinear terms) on ragged grids: Lm not a multiple of 16,
# fewer rows than a chunk, N = 30.  "v8" = the production step3d_t (k_step3d_t8.cu: loader / producer / consumer warps as fibers,
# mbarriers with transaction counts, TMA boxes with out-of-bound fill and the 16-byte start rule, tensor-memory lanes emulated):
# with 1-3 "SMs" a persistent CTA walks several work items (ring and slot recycling across items), with 148 every CTA gets one.
# N = 9 / 50: odd level count / four level pairs per producer warp; N = 64: the slots do not fit -> falls back to "v6", the
# round-1 warp-specialised kernel (k_step3d_t6.cu: named barriers, warp vote), which is also run on its own; "v4" = the
# shuffle-based column march (the fallback for closed W/E walls and N < 4).
CASES = [(0, 24, 10, 8, 3, "v8", 3), (0, 0, 0, 0, 2, "v8", 148),                       # UPWELLING small and as shipped (41x80x16)
         (1, 20, 6, 8, 3, "v8", 2), (1, 33, 5, 9, 3, "v8", 2), (1, 70, 9, 30, 3, "v8", 148), (1, 70, 9, 30, 2, "v8", 1),
         (1, 45, 7, 50, 2, "v8", 2), (1, 40, 6, 64, 2, "v8", 148),                          # the ragged shapes of the GPU tests
         (1, 70, 9, 30, 2, "v6", 2), (1, 45, 7, 50, 1, "v6", 2), (1, 33, 9, 10, 2, "v4", 148)]


@pytest.mark.parametrize("app,Lm,Mm,N,steps,s3t,nsm", CASES)
def test_kernel_sources_match_oracle_on_cpu(emu_lib, app, Lm, Mm, N, steps, s3t, nsm):
    r = subprocess.run([sys.executable, os.path.join(HERE, "emu_worker.py")] + [str(x) for x in (app, Lm, Mm, N, steps)] + [s3t],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ, EMU_SM_COUNT=str(nsm)))
    assert r.returncode == 0 and "EMU-PARITY-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


# app, Lm, Mm, N, steps, NtileI, NtileJ: the tilings of the scaling run (2x1, 2x2, 4x2) and an N-S split, against ONE tile
@pytest.mark.parametrize("app,Lm,Mm,N,steps,nti,ntj", [(1, 48, 24, 10, 2, 2, 1), (1, 48, 24, 10, 2, 2, 2), (1, 64, 32, 8, 2, 4, 2),
                                                       (1, 48, 36, 10, 2, 1, 2), (0, 40, 24, 8, 3, 2, 2)])
def test_tiling_invariance_on_emulated_kernels(emu_lib, app, Lm, Mm, N, steps, nti, ntj):
    """ROMS/Bin/verify.sh: results must not depend on the tiling.  Every rank is a host thread with its own mirror, the halo
    swaps of k_halo.cu are in-process copies along roms_b200_halo_plan (tests/emu/emu_rt.cpp): checks the distributed LOGIC
    (tile bounds, redundant evaluation on halos, deep-halo predictor, what is swapped when, diag's gather) bit for bit,
    including wvel, Akv, W, Huon, rho and the 13 diag outputs."""
    r = subprocess.run([sys.executable, os.path.join(HERE, "emu_worker.py"), "tiles"] + [str(x) for x in (app, Lm, Mm, N, steps, nti, ntj)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "EMU-TILES-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("seed,s3t,extra", [(1, "v8", {"ROMS_B200_S3T_JCH": "3"}), (2, "v8", {"ROMS_B200_S3T_SLOTS": "2"}),
                                            (3, "v8", {"ROMS_B200_S3T_TMEM": "0"}), (4, "v6", {})])
def test_step3d_t_synchronisation_under_random_schedules(emu_lib, seed, s3t, extra):
    """The pipelined step3d_t (TMA ring and slots handed over through full/ready/empty mbarriers; named barriers in the round-1
    kernel) with the threads of a block resumed in a pseudo-random order at every scheduling pass, with short chunks (many work
    items per CTA), the minimum number of slots, and the shared-memory CF/DC variant: results must not depend on the schedule."""
    r = subprocess.run([sys.executable, os.path.join(HERE, "emu_worker.py"), "1", "70", "9", "30", "2", s3t], capture_output=True, text=True,
                       timeout=900, env=dict(os.environ, EMU_SM_COUNT="1", EMU_SCHED_SEED=str(seed), **extra))
    assert r.returncode == 0 and "EMU-PARITY-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("case,seed", [((1, 70, 9, 30, 2), 5), ((1, 33, 5, 9, 2), 11), ((0, 24, 10, 8, 3), 12), ((1, 130, 9, 8, 2), 7)])
def test_persistent_fast_loop_neighbour_flags_under_random_block_schedules(emu_lib, case, seed):
    """ROMS_B200_S2_PERSIST=1: the 2*nfast+1 sub-steps of a baroclinic step in ONE kernel, every block waiting for the sub-step
    counters of the blocks of the adjacent tiles (k_step2d.cu).  All blocks run as fibers of one scheduler that leaves a random
    half of the blocks out in every pass, so blocks drift apart by whole sub-steps unless the flags hold them together; 33 wide:
    the last tile of a periodic row is narrower than the stencil reach (neighbourhood of two tiles).  Bit-identical to the oracle.
    Negative control: with the waits skipped (EMU_S2P_NOWAIT) the same schedule must fail."""
    args = [sys.executable, os.path.join(HERE, "emu_worker.py")] + [str(x) for x in case] + ["v8"]
    env = dict(os.environ, EMU_SM_COUNT="148", EMU_SCHED_SEED=str(seed), ROMS_B200_S2_PERSIST="1")
    r = subprocess.run(args, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "EMU-PARITY-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    if seed == 5:
        r = subprocess.run(args, capture_output=True, text=True, timeout=900, env=dict(env, EMU_S2P_NOWAIT="1"))
        assert r.returncode != 0 and "step2d_loop" in r.stderr, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("case", [(1, 70, 19, 30, 2), (0, 24, 10, 8, 2)])
def test_level_major_block_order_is_a_pure_renumbering(emu_lib, case):
    """common.cuh level_major(): the level-parallel kernels renumber their blocks (bands of tiles, level by level) on grids of more
    than LM_G block tiles; EMU_LM_G=3 makes the small test grids take that path (bands of 3 tiles, a ragged last band)."""
    r = subprocess.run([sys.executable, os.path.join(HERE, "emu_worker.py")] + [str(x) for x in case] + ["v8"],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ, EMU_SM_COUNT="148", EMU_LM_G="3"))
    assert r.returncode == 0 and "EMU-PARITY-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("case", [(1, 70, 19, 30, 3), (0, 24, 10, 8, 3), (1, 45, 7, 50, 2)])
def test_marching_kernels_whole_column_mode(emu_lib, case):
    """The level-marching kernels (t3dmix2_geo, pre_step3d tracers and momentum, uv3dmix2, rhs3d) split the column into chunks on
    small grids -- what the other emulation cases run -- and take the WHOLE column per block on grids with enough tiles: then
    rufrc/rvfrc are summed in registers and the sum kernels are not launched.  *_FILL=0 forces that mode on the small test grids."""
    env = dict(os.environ, EMU_SM_COUNT="148", ROMS_B200_T3DMIX_FILL="0", ROMS_B200_PRE3D_FILL="0", ROMS_B200_PRE3DUV_FILL="0",
               ROMS_B200_UVMIX_FILL="0", ROMS_B200_RHS3D_FILL="0")
    r = subprocess.run([sys.executable, os.path.join(HERE, "emu_worker.py")] + [str(x) for x in case] + ["v8"],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "EMU-PARITY-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_emulated_rho_eos_matches_the_reference_check_values(emu_lib):
    r = subprocess.run([sys.executable, os.path.join(HERE, "emu_worker.py"), "eos"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "EMU-EOS-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_host_driver_on_emulated_kernels(emu_lib):
    """ROMS_initialize / ROMS_run (C++ host driver, roms_b200/csrc/host_driver.cpp) over the emulated kernels: start state, the
    device-resident loop, the host-forcing loop with diag read back every step and the blow-up stop -- all bit-identical to
    the oracle (glibc on both sides)."""
    r = subprocess.run([sys.executable, os.path.join(HERE, "emu_worker.py"), "driver", "1", "33", "9", "10", "3"], capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0 and "EMU-DRIVER-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("app,grid", [(1, (33, 9, 10)), (0, (24, 10, 8))])
def test_perfect_restart_field_list(emu_lib, app, grid):
    """roms_b200_restart_fields (the reference's PERFECT_RESTART record, Utility/wrt_rst.F:178-900, for both option sets) written
    after 5 steps and read into a fresh context continues bit-identically to the uninterrupted run."""
    r = subprocess.run([sys.executable, os.path.join(HERE, "emu_worker.py"), "restart", str(app)] + [str(x) for x in grid], capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0 and "EMU-RESTART-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_kernel_sources_memory_safe_under_asan(emu_lib):
    """Same protocol with the emulation built under AddressSanitizer: every mirror field is its own heap block and every
    shared-memory tile its own static array, so a stencil index outside a field or a tile aborts the worker."""
    asan = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan not available")
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(HERE, "emu"), "ASAN=1"])
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0", EMU_WORKER_LIB="libroms_b200_emu_asan.so", EMU_SM_COUNT="2",
               EMU_TEAM="threads")          # AddressSanitizer does not follow swapcontext: OS-thread teams
    r = subprocess.run([sys.executable, os.path.join(HERE, "emu_worker.py"), "1", "33", "9", "10", "2", "v8"], capture_output=True, text=True,
                       timeout=900, env=env)
    assert r.returncode == 0 and "EMU-PARITY-OK" in r.stdout and "AddressSanitizer" not in r.stderr, r.stdout[-2000:] + r.stderr[-4000:]
    r = subprocess.run([sys.executable, os.path.join(HERE, "emu_worker.py"), "1", "33", "9", "10", "1", "v6"], capture_output=True, text=True,
                       timeout=900, env=dict(env, ROMS_B200_T3DMIX_FILL="0", ROMS_B200_PRE3D_FILL="0", ROMS_B200_UVMIX_FILL="0", ROMS_B200_RHS3D_FILL="0"))
    assert r.returncode == 0 and "EMU-PARITY-OK" in r.stdout and "AddressSanitizer" not in r.stderr, r.stdout[-2000:] + r.stderr[-4000:]


def test_emulation_library_is_not_the_product(emu_lib):
    """The product binding loads roms_b200/libroms_b200.so only; the emulation library lives under tests/."""
    import roms_b200 as rb
    assert os.path.basename(rb.library_path()) == "libroms_b200.so" and "tests" not in rb.library_path()
    assert os.path.dirname(emu_lib).endswith(os.path.join("tests", "emu"))
