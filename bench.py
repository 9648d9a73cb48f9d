#!/usr/bin/env python
"""bench.py -- baroclinic-step throughput of the roms_b200 main3d path.

    python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference --gpus N ...          # the reference's own CPU algorithm (oracle port, -O3 build)

Metric (BASELINE.json): 3-D cell-updates/s of the baroclinic step = Lm*Mm*N*K / time(K steps of main3d),
workload BENCHMARK1 (512x64x30, full main3d loop: EOS, KPP, bulk fluxes, 59 barotropic sub-steps, ...).
Prints ONE JSON line (rank 0).  The line is self-checking:
  N = 1: "parity"               -- zeta,ubar,vbar,u,v,T,S after --parity-steps steps against the oracle (the CHECKER build of oracle/,
                                   -O2 -ffp-contract=off), max-norm relative to the field's range, bar 1e-10 (BASELINE.json);
  N > 1: "tiling_bit_identical" -- every rank also integrates the WHOLE grid as one tile on its own GPU through the same call
                                   sequence and compares its tile's interior bit for bit (the reference's acceptance criterion,
                                   ROMS/Bin/verify.sh:12-14, check_nc.sh:35-43).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "3D cell-updates/sec (baroclinic step)"
UNIT = "cell-updates/s"
WORKLOADS = {"BENCHMARK1": (512, 64, 30), "BENCHMARK2": (1024, 128, 30), "BENCHMARK3": (2048, 256, 30)}
PROGNOSTIC = ["zeta", "ubar", "vbar", "u", "v", "t"]


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.p, self.index = [], None, index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _oracle_tiling(cores):
    nti, ntj = 1, 1
    while nti * ntj < cores:            # tiles = host threads, like the reference's NtileI*NtileJ = cores
        if nti <= ntj * 4:
            nti *= 2
        else:
            ntj *= 2
    return nti, ntj


def cpu_baseline(workload, nsteps, warm=1, world=1):
    """The oracle (CPU restatement of the reference algorithm, kind="port": the Fortran cannot be built in this image) on this
    box's host cores, compiled like the reference's own build (-O3, contraction on, -march=native when the box has a compiler)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    Lm, Mm, N = grid_for(workload, world)[:3]
    cores = os.cpu_count() or 1
    nti, ntj = _oracle_tiling(cores)
    flags = ol.fast_lib()[1]
    o = ol.Oracle(ol.BENCHMARK, Lm, Mm, N, NtileI=nti, NtileJ=ntj, fast=True)
    o.set_threads(cores)
    o.initial()
    o.step(warm)
    t0 = time.perf_counter()
    o.step(nsteps)
    dt = time.perf_counter() - t0
    return {"value": Lm * Mm * N * nsteps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d baroclinic steps on the %dx%dx%d grid, %dx%d tiles on %d host threads, %.2f s; oracle (C++ restatement of the "
                      "Fortran, which cannot be built in this image) compiled %s" % (nsteps, Lm, Mm, N, nti, ntj, cores, dt, flags)}, dt / nsteps


def parity_vs_oracle(rb, workload, nsteps):
    """zeta,ubar,vbar,u,v,t after nsteps baroclinic steps from the analytical start state: this library (device-resident loop)
    against the checker build of the oracle."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_lib as ol
    Lm, Mm, N = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    nti, ntj = _oracle_tiling(cores)
    o = ol.Oracle(ol.BENCHMARK, Lm, Mm, N, NtileI=nti, NtileJ=ntj)      # tiling-invariant bit for bit (tests/test_cpu.py)
    o.set_threads(cores)
    o.initial()
    o.step(nsteps)
    d = rb.Driver(rb.default_config(rb.APP_BENCHMARK, Lm, Mm, N), device=0)
    d.run(nsteps)
    d.ctx.sync()
    from parity_common import prognostic_errors
    rel, own = prognostic_errors(o.get, d.ctx.download)
    d.finalize()
    worst = max(rel.values())
    return {"against": "oracle (checker build -O2 -ffp-contract=off, %dx%d tiles)" % (nti, ntj), "workload": "%s %dx%dx%d" % (workload, Lm, Mm, N),
            "steps": nsteps, "max_rel": worst, "tolerance": 1e-10, "ok": bool(worst <= 1e-10), "per_field": rel,
            "scale": "max-norm of the difference / range of the field; velocity components / range of the larger component of their vector",
            "per_field_over_own_range": own}


def run_reference(args, rank, world):
    if rank != 0:
        return
    world = world if world > 1 else max(1, args.gpus)      # the grid of OUR arm at this GPU count
    cb, sps = cpu_baseline(args.workload, args.steps, max(1, min(args.warmup, 2)), world)
    line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (analytical BENCHMARK grid/initial state/forcing, random-free)", "impl": "reference",
            "config": workload_config(args.workload, world),
            "cpu_baseline": cb, "gpu_launches": 0,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


TILINGS = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}     # BENCHMARK2 = 2x2 on 4 GPUs, 4x2 on 8 (BASELINE.json configs)


def grid_for(workload, world):
    """Global grid and tiling of a run on `world` GPUs: the BENCHMARK1 series is weak-scaled (every GPU holds one 512x64x30 tile,
    N=4 is exactly BENCHMARK2 1024x128x30 on 2x2 tiles); any other workload keeps its grid (strong-scaling style)."""
    nti, ntj = TILINGS[world]
    Lm, Mm, N = WORKLOADS[workload]
    if workload == "BENCHMARK1":
        Lm, Mm = Lm * nti, Mm * ntj
    return Lm, Mm, N, nti, ntj


def workload_config(workload, world):
    """The same `config` object for both arms (the driver compares them)."""
    Lm, Mm, N, nti, ntj = grid_for(workload, world)
    name = workload if world == 1 else "%s-sized tile per GPU" % workload if workload == "BENCHMARK1" else workload
    return {"workload": "%s: %dx%dx%d grid, full main3d loop (rho_eos, diag every step, bulk_flux, KPP, omega, wvelocity, 2*nfast+1 step2d "
                        "sub-steps, rhs3d, step3d_uv, step3d_t)" % (name, Lm, Mm, N),
            "grid": "%dx%dx%d" % (Lm, Mm, N), "gpu_tiles": "%dx%d" % (nti, ntj)}


def ncu_traffic():
    """DRAM bytes per launch of the graded kernel: dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from an
    `ncu --set full` capture (ncu cannot run inside the timed bench); written by tools/ncu_traffic.py together with the commit
    the capture was taken at."""
    try:
        with open(os.path.join(ROOT, "profiles", "step3d_t_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def _time_step3d_t(rb, Lm, Mm, N, reps):
    cfg = rb.default_config(rb.APP_BENCHMARK, Lm, Mm, N)
    cfg.dt, cfg.ndtfast = 20.0, 20          # smaller DT for the finer synthetic grid (arithmetic unchanged)
    d = rb.Driver(cfg)
    d.run(3)                                # non-degenerate state (upwind branches active)
    st, _ = d.ctx.get_stepping()
    d.ctx.time_step3d_t(st["nrhs"], st["nstp"], st["nnew"], 3)
    ms = d.ctx.time_step3d_t(st["nrhs"], st["nstp"], st["nnew"], reps)     # CUDA events on the launch stream
    d.finalize()
    return ms


def roofline_step3d_t(rb, peak, peak_kind, basin_tile=True):
    """step3d_t (the graded kernel) on grids whose working set is >> L2 (126 MB), so every launch streams from HBM:
    96 B algorithmic per cell per call (NT=2).  Primary: the BENCHMARK3 grid (N=30); also a quarter
    of and a whole 1024x2048x50 tile of the 4096^2 x 50 basin (also, also2).  The timing loop re-applies the kernel to its own output (same traffic, arithmetic not meaningful)."""
    out = None
    traffic = ncu_traffic()
    grids = ((2048, 256, 30), (1024, 512, 50), (1024, 2048, 50))     # the last: one tile of the 4096^2 x 50 basin on 4x2 GPUs
    for (Lm, Mm, N) in (grids if basin_tile else grids[:2]):
        try:
            ms = _time_step3d_t(rb, Lm, Mm, N, 20 if Mm < 2048 else 8)
        except Exception as e:                 # (the basin tile needs ~30 GB of HBM)
            if out is not None:
                out["also2"] = {"grid": "%dx%dx%d" % (Lm, Mm, N), "error": str(e)[:200]}
                continue
            raise
        cells = Lm * Mm * N
        achieved = 96.0 * cells / (ms * 1e-3) / 1e9
        tr = traffic.get("%dx%dx%d" % (Lm, Mm, N), {})
        r = {"bound": "hbm", "kernel": "step3d_t_v8_kernel (TMA-staged j x k tiles, mbarrier pipeline, CF/DC in tensor memory)", "achieved": achieved,
             "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_kind": peak_kind, "traffic": tr.get("dram_bytes"),
             "traffic_source": tr.get("source"),
             "grid": "%dx%dx%d (t working set %.2f GB >> L2)" % (Lm, Mm, N, 6 * cells * 8 / 1e9),
             "ms_per_launch": ms, "cell_updates_per_s": cells / (ms * 1e-3), "algorithmic_bytes_per_cell": 96,
             "algorithmic_bytes_per_launch": 96 * cells}
        if out is None:
            out = r
        else:
            out["also" if "also" not in out else "also2"] = r
    return out


def run_ours(args, rank, world):
    import roms_b200 as rb
    if world > 1:
        return run_ours_multi(args, rank, world)
    Lm, Mm, N = WORKLOADS[args.workload]
    cfg = rb.default_config(rb.APP_BENCHMARK, Lm, Mm, N)
    d = rb.Driver(cfg, device=0)
    cells = Lm * Mm * N
    # ---- device-resident throughput (inputs already in HBM, forcing evaluated on the device)
    clk = ClockSampler(0)
    clk.start()
    d.run(max(args.warmup, 3))
    l0 = d.ctx.launches()
    d.ctx.sync()
    d.timer_start()
    d.run(args.steps)
    ms = d.timer_stop()
    clocks = clk.stop()
    launches = d.ctx.launches() - l0
    value = cells * args.steps / (ms * 1e-3)
    # ---- end to end through the driver surface with HOST buffers: per step the host evaluates
    # set_data (ana_srflux), uploads it (H2D), runs main3d, and reads the diag scalars back (D2H)
    d.run(2, host_forcing=True)
    d.ctx.sync()
    t0 = time.perf_counter()
    diag = d.run(args.steps, host_forcing=True)
    d.ctx.sync()
    e2e_s = time.perf_counter() - t0
    b = rb.tile_bounds(Lm, Mm, N)
    ni, nj = b.UBi - b.LBi + 1, b.UBj - b.LBj + 1
    e2e = {"value": cells * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": ni * nj * 8, "d2h_bytes_per_step": (3 * (b.Iend - b.Istr + 1) + 9) * 8,
           "ms_per_step": 1e3 * e2e_s / args.steps, "last_diag": {"avgke": diag[0], "avgpe": diag[1], "volume": diag[2]}}
    nfast = d.nfast
    d.finalize()
    peak, peak_kind = measured_peak()
    roof = roofline_step3d_t(rb, peak, peak_kind) if not args.no_roofline else None
    cb = parity = None
    if not args.no_cpu:
        cb, _ = cpu_baseline(args.workload, args.cpu_steps)
        parity = parity_vs_oracle(rb, args.workload, args.parity_steps)
    cfgd = workload_config(args.workload, 1)
    notes = {"nfast": nfast, "l2": "state 0.3 GB per step > 126 MB L2, no explicit flush", "fmad": "false (parity build)"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (analytical BENCHMARK grid/initial state/forcing, random-free)",
            "config": cfgd, "notes": notes, "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cb,
            "parity": parity}
    print(json.dumps(line))


TILING_FIELDS = [("zeta", 1, 1), ("zeta", 2, 1), ("ubar", 1, 1), ("ubar", 2, 1), ("vbar", 1, 1), ("vbar", 2, 1), ("u", 1, 1), ("u", 2, 1),
                 ("v", 1, 1), ("v", 2, 1), ("t", 1, 1), ("t", 2, 1), ("t", 1, 2), ("t", 2, 2)]


def _make_distributed(rb, dist, torch, cfg, rank, world, local):
    """One tile per rank: NCCL id broadcast + all-gather of the CUDA IPC handles of the NVLink mailboxes (what MPI_Bcast /
    MPI_Allgather do in a Fortran host)."""
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.frombuffer(bytearray(rb.comm_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(idt, 0)
    d = rb.Driver(cfg, tile=rank, device=local)
    d.comm_init(rank, world, bytes(idt.cpu().numpy().tobytes()))
    hnd = torch.frombuffer(bytearray(d.p2p_handle()), dtype=torch.uint8).cuda()
    allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(allh, hnd)
    d.p2p_connect(b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh), world)
    dist.barrier()
    return d


def _tiling_check(rb, dist, torch, d, gcfg, local, sequence):
    """Tiling invariance on the benchmarked grid: this rank integrates the whole grid as ONE tile on its own GPU through the same
    call sequence the tiled run went through, then compares its tile's interior of every prognostic field bit for bit."""
    import numpy as np
    one = rb.Driver(gcfg, device=local)
    for n, host in sequence:
        one.run(n, host_forcing=host)
    one.ctx.sync()
    one.ctx._bounds = one.bounds()
    b = d.bounds()
    d.ctx._bounds = b
    ok, worst = True, 0.0
    for name, l, m in TILING_FIELDS:
        ref = one.ctx.download_interior(name, l, m)[:, b.Jstr - 1:b.Jend, b.Istr - 1:b.Iend]
        got = d.ctx.download_interior(name, l, m)
        if not np.array_equal(ref, got):
            ok = False
            worst = max(worst, float(np.max(np.abs(ref - got))))
    one.finalize()
    flag = torch.tensor([1.0 if ok else 0.0, -worst], dtype=torch.float64, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag[0] > 0.5), float(-flag[1])


def run_ours_multi(args, rank, world):
    """Weak scaling: every GPU holds one BENCHMARK1-sized tile (512x64x30); the global grid grows with N (N=4 is exactly
    BENCHMARK2 1024x128x30 on 2x2 tiles).  Halo swaps: NVLink peer mailboxes inside the library (k_halo.cu)."""
    import torch
    import torch.distributed as dist
    import roms_b200 as rb
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gLm, gMm, N, nti, ntj = grid_for(args.workload, world)

    def timed(d, fn):
        d.ctx.sync(); dist.barrier(); torch.cuda.synchronize()
        d.timer_start(); t0 = time.perf_counter()
        fn()
        ms = d.timer_stop(); d.ctx.sync(); wall = time.perf_counter() - t0
        dist.barrier(); torch.cuda.synchronize()
        tt = torch.tensor([ms * 1e-3, wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt[0]), float(tt[1])

    def one_grid(gLm, gMm, N, steps, warm, check):
        cfg = rb.default_config(rb.APP_BENCHMARK, gLm, gMm, N)
        gcfg = rb.default_config(rb.APP_BENCHMARK, gLm, gMm, N)
        cfg.NtileI, cfg.NtileJ = nti, ntj
        d = _make_distributed(rb, dist, torch, cfg, rank, world, local)
        clk = ClockSampler(local)
        if rank == 0:
            clk.start()
        d.run(warm)
        l0 = d.ctx.launches()
        dev_s, _ = timed(d, lambda: d.run(steps))
        clocks = clk.stop() if rank == 0 else None
        launches = d.ctx.launches() - l0
        d.run(2, host_forcing=True)
        _, e2e_s = timed(d, lambda: d.run(steps, host_forcing=True))
        b = d.bounds()
        ident, worst = (None, None)
        if check:
            ident, worst = _tiling_check(rb, dist, torch, d, gcfg, local, [(warm, False), (steps, False), (2, True), (steps, True)])
        d.finalize()
        dist.barrier()
        return {"dev_s": dev_s, "e2e_s": e2e_s, "launches": launches, "clocks": clocks, "bounds": b, "ident": ident, "worst": worst,
                "cells": gLm * gMm * N}

    warm = max(args.warmup, 3)
    r = one_grid(gLm, gMm, N, args.steps, warm, not args.no_check)
    also = None
    if world == 8 and args.workload == "BENCHMARK1" and not args.no_also:
        # BASELINE.json config 4: BENCHMARK3 (2048x256x30) on 4x2 tiles of 512x128, and the same tile alone on one GPU
        b3 = one_grid(2048, 256, 30, args.steps, warm, not args.no_check)
        t1 = None
        if rank == 0:
            s = rb.Driver(rb.default_config(rb.APP_BENCHMARK, 512, 128, 30), device=local)
            s.run(warm); s.ctx.sync(); s.timer_start(); s.run(args.steps); t1 = s.timer_stop() * 1e-3; s.finalize()
        dist.barrier()
        also = {"workload": "BENCHMARK3 2048x256x30 full main3d loop, 4x2 tiles of 512x128 (one per GPU)", "ms_per_step": 1e3 * b3["dev_s"] / args.steps,
                "value": b3["cells"] * args.steps / b3["dev_s"], "unit": UNIT, "tiling_bit_identical": b3["ident"],
                "one_gpu_512x128x30_ms_per_step": (1e3 * t1 / args.steps) if t1 else None,
                "weak_scaling_efficiency_vs_one_tile": (t1 / b3["dev_s"]) if t1 else None}
    roof = None
    if rank == 0 and not args.no_roofline:
        peak, peak_kind = measured_peak()
        roof = roofline_step3d_t(rb, peak, peak_kind, basin_tile=False)      # single-GPU kernel measurement on rank 0's GPU (the N=1 line also carries the 30 GB basin tile)
    dist.barrier()
    if rank == 0:
        b = r["bounds"]
        ni, nj = b.UBi - b.LBi + 1, b.UBj - b.LBj + 1
        cells = r["cells"]
        cfgd = workload_config(args.workload, world)
        notes = {"tiles": "%dx%d tiles of %dx%d, one per GPU (the N=1 line of this series is BENCHMARK1 512x64x30)" % (nti, ntj, gLm // nti, gMm // ntj),
                "halo": ("NVLink peer mailboxes (CUDA IPC): one kernel per exchange writes the strips and corner blocks of all 8 neighbours "
                         "into their mailboxes as flag-in-data messages (8-byte words carrying the sequence number, no fences or flags) and polls/unpacks "
                         "what arrives, strips split into 256-element work items over the blocks, the kernel chained to the sub-steps around it by "
                         "programmatic dependent launch; mirror halo 6, deep-halo fast loop (one 3-field swap per barotropic sub-step pair), 33 swaps "
                         "per step; NCCL only for diag's all-reduce; fast loop replayed as a CUDA graph") if os.environ.get("ROMS_B200_HALO_NCCL") is None
                        else "NCCL send/recv (pack kernel, grouped send/recv, unpack kernel), W/E then S/N",
                "l2": "state 0.3 GB per GPU per step > 126 MB L2, no explicit flush", "fmad": "false (parity build)"}
        line = {"metric": METRIC, "value": cells * args.steps / r["dev_s"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": 1e3 * r["dev_s"] / args.steps, "higher_is_better": True,
                "scaling": "weak" if args.workload == "BENCHMARK1" else "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic (analytical BENCHMARK grid/initial state/forcing, random-free)",
                "config": cfgd, "notes": notes, "clocks": r["clocks"],
                "e2e": {"value": cells * args.steps / r["e2e_s"], "unit": UNIT, "h2d_bytes_per_step": ni * nj * 8 * world,
                        "d2h_bytes_per_step": (3 * (b.Iend - b.Istr + 1) + 9) * 8 * world, "ms_per_step": 1e3 * r["e2e_s"] / args.steps},
                "gpu_launches": r["launches"] * world, "roofline": roof, "cpu_baseline": None,
                "tiling_bit_identical": r["ident"],
                "tiling_check": {"against": "the whole grid as ONE tile on each rank's own GPU, same call sequence", "fields": "zeta,ubar,vbar,u,v,t (both time levels), tile interiors",
                                 "steps": warm + 2 * args.steps + 2, "max_abs_diff": r["worst"]},
                "also": also}
        print(json.dumps(line))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="BENCHMARK1", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-steps", type=int, default=20)
    ap.add_argument("--parity-steps", type=int, default=100)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--no-also", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
