#!/usr/bin/env python
"""bench.py -- baroclinic-step throughput of the roms_b200 main3d path.

    python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference --gpus N ...          # the reference's own CPU algorithm (oracle port)

Metric (BASELINE.json): 3-D cell-updates/s of the baroclinic step = Lm*Mm*N*K / time(K steps of main3d),
workload BENCHMARK1 (512x64x30, full main3d loop: EOS, KPP, bulk fluxes, 59 barotropic sub-steps, ...).
Prints ONE JSON line (rank 0).
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


def cpu_baseline(workload, nsteps, warm=1):
    """The oracle (CPU restatement of the reference algorithm, kind="port") on this box's host cores."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    Lm, Mm, N = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    nti, ntj = 1, 1
    while nti * ntj < cores:            # tiles = host threads, like the reference's NtileI*NtileJ = cores
        if nti <= ntj * 4:
            nti *= 2
        else:
            ntj *= 2
    o = ol.Oracle(ol.BENCHMARK, Lm, Mm, N, NtileI=nti, NtileJ=ntj)
    o.set_threads(cores)
    o.initial()
    o.step(warm)
    t0 = time.perf_counter()
    o.step(nsteps)
    dt = time.perf_counter() - t0
    return {"value": Lm * Mm * N * nsteps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d baroclinic steps of %s (%dx%dx%d), %dx%d tiles on %d threads, %.2f s" % (nsteps, workload, Lm, Mm, N, nti, ntj, cores, dt)}, dt / nsteps


def run_reference(args, rank, world):
    if rank != 0:
        return
    cb, sps = cpu_baseline(args.workload, args.steps, max(1, min(args.warmup, 2)))
    Lm, Mm, N = WORKLOADS[args.workload]
    line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (analytical BENCHMARK grid/initial state/forcing)", "impl": "reference",
            "config": {"workload": "%s %dx%dx%d full main3d loop" % (args.workload, Lm, Mm, N)},
            "cpu_baseline": cb, "gpu_launches": 0,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# DRAM bytes per launch of the graded kernel from `ncu --set full` (profiles/r01_step3d_t_v6_*): dram__bytes_read.sum +
# dram__bytes_write.sum of ONE launch on the same grid (ncu cannot run inside the timed bench)
NCU_TRAFFIC = {(2048, 256, 30): 1.320e9 + 0.248e9, (1024, 512, 50): 2.312e9 + 0.416e9}


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


def roofline_step3d_t(rb, peak, peak_kind):
    """step3d_t (the graded kernel) on grids whose working set is >> L2 (126 MB), so every launch streams from HBM:
    96 B algorithmic per cell per call (NT=2).  Primary: the BENCHMARK3 grid (N=30); also the
    N=50 basin-like tile, where the kernel's shared-memory ring leaves room for only one row in flight."""
    out = None
    for (Lm, Mm, N) in ((2048, 256, 30), (1024, 512, 50)):
        ms = _time_step3d_t(rb, Lm, Mm, N, 20)
        cells = Lm * Mm * N
        achieved = 96.0 * cells / (ms * 1e-3) / 1e9
        r = {"bound": "hbm", "kernel": "step3d_t_v6_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
             "peak_kind": peak_kind, "traffic": NCU_TRAFFIC.get((Lm, Mm, N)),
             "grid": "%dx%dx%d (t working set %.2f GB >> L2)" % (Lm, Mm, N, 6 * cells * 8 / 1e9),
             "ms_per_launch": ms, "cell_updates_per_s": cells / (ms * 1e-3), "algorithmic_bytes_per_cell": 96,
             "algorithmic_bytes_per_launch": 96 * cells}
        if out is None:
            out = r
        else:
            out["also"] = r
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
    ni, nj = cfg.Lm + 6 + (0 if cfg.Lm % 2 else 0), cfg.Mm + 3
    b = rb.tile_bounds(Lm, Mm, N)
    ni, nj = b.UBi - b.LBi + 1, b.UBj - b.LBj + 1
    e2e = {"value": cells * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": ni * nj * 8, "d2h_bytes_per_step": (3 * (b.Iend - b.Istr + 1) + 9) * 8,
           "ms_per_step": 1e3 * e2e_s / args.steps, "last_diag": {"avgke": diag[0], "avgpe": diag[1], "volume": diag[2]}}
    d.finalize()
    peak, peak_kind = measured_peak()
    roof = roofline_step3d_t(rb, peak, peak_kind) if not args.no_roofline else None
    cb = None
    if not args.no_cpu:
        cb, _ = cpu_baseline(args.workload, args.cpu_steps)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (analytical BENCHMARK grid/initial state/forcing, random-free)",
            "config": {"workload": "%s %dx%dx%d full main3d loop (rho_eos, diag every step, bulk_flux, KPP, omega, wvelocity, %d step2d sub-steps, rhs3d, step3d_uv, step3d_t)"
                       % (args.workload, Lm, Mm, N, 2 * d.nfast + 1), "tiles": "1x1", "l2": "state 0.3 GB per step > 126 MB L2, no explicit flush",
                       "fmad": "false (parity build)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cb}
    print(json.dumps(line))


TILINGS = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}     # BENCHMARK2 = 2x2 on 4 GPUs, 4x2 on 8 (BASELINE.json configs)


def run_ours_multi(args, rank, world):
    """Weak scaling: every GPU holds one BENCHMARK1-sized tile (512x64x30); the global grid grows with N
    (N=4 is exactly BENCHMARK2 1024x128x30 on 2x2 tiles).  Halo swaps = NCCL send/recv inside the library."""
    import torch
    import torch.distributed as dist
    import roms_b200 as rb
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nti, ntj = TILINGS[world]
    tLm, tMm, N = WORKLOADS["BENCHMARK1"]
    if args.workload != "BENCHMARK1":                      # e.g. --workload BENCHMARK3 --gpus 8 (strong-scaling style run)
        gLm, gMm, N = WORKLOADS[args.workload]
    else:
        gLm, gMm = tLm * nti, tMm * ntj
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.frombuffer(bytearray(rb.comm_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(idt, 0)
    cfg = rb.default_config(rb.APP_BENCHMARK, gLm, gMm, N)
    cfg.NtileI, cfg.NtileJ = nti, ntj
    d = rb.Driver(cfg, tile=rank, device=local)
    d.comm_init(rank, world, bytes(idt.cpu().numpy().tobytes()))
    # NVLink peer mailboxes: all-gather the CUDA IPC handles (what MPI_Allgather does in a Fortran host)
    hnd = torch.frombuffer(bytearray(d.p2p_handle()), dtype=torch.uint8).cuda()
    allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(allh, hnd)
    d.p2p_connect(b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh), world)
    dist.barrier()
    cells = gLm * gMm * N

    def timed(fn):
        d.ctx.sync(); dist.barrier(); torch.cuda.synchronize()
        d.timer_start(); t0 = time.perf_counter()
        fn()
        ms = d.timer_stop(); d.ctx.sync(); wall = time.perf_counter() - t0
        dist.barrier(); torch.cuda.synchronize()
        tt = torch.tensor([ms * 1e-3, wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt[0]), float(tt[1])

    clk = ClockSampler(local)
    if rank == 0:
        clk.start()
    d.run(max(args.warmup, 3))
    l0 = d.ctx.launches()
    dev_s, _ = timed(lambda: d.run(args.steps))
    clocks = clk.stop() if rank == 0 else None
    launches = d.ctx.launches() - l0
    d.run(2, host_forcing=True)
    _, e2e_s = timed(lambda: d.run(args.steps, host_forcing=True))
    b = d.bounds()
    ni, nj = b.UBi - b.LBi + 1, b.UBj - b.LBj + 1
    d.finalize()
    roof = None
    if rank == 0 and not args.no_roofline:
        peak, peak_kind = measured_peak()
        roof = roofline_step3d_t(rb, peak, peak_kind)      # single-GPU kernel measurement on rank 0's GPU
    dist.barrier()
    if rank == 0:
        line = {"metric": METRIC, "value": cells * args.steps / dev_s, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
                "scaling": "weak" if args.workload == "BENCHMARK1" else "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic (analytical BENCHMARK grid/initial state/forcing, random-free)",
                "config": {"workload": "%dx%dx%d full main3d loop, %dx%d tiles of %dx%d (one per GPU)" % (gLm, gMm, N, nti, ntj, gLm // nti, gMm // ntj),
                           "halo": "NVLink peer mailboxes (CUDA IPC, remote stores from the pack kernel + flags), 2-phase W/E then S/N, width 3, aggregated per kernel; NCCL for the diag all-reduce; fast loop in a CUDA graph" if os.environ.get("ROMS_B200_HALO_NCCL") is None else "NCCL send/recv, 2-phase W/E then S/N, width 3",
                           "l2": "state 0.3 GB per GPU per step > 126 MB L2, no explicit flush", "fmad": "false (parity build)"},
                "clocks": clocks,
                "e2e": {"value": cells * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": ni * nj * 8 * world,
                        "d2h_bytes_per_step": (3 * (b.Iend - b.Istr + 1) + 9) * 8 * world, "ms_per_step": 1e3 * e2e_s / args.steps},
                "gpu_launches": launches * world, "roofline": roof, "cpu_baseline": None}
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
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
