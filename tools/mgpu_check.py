"""Multi-GPU tiling-invariance check (the reference's own acceptance test, ROMS/Bin/verify.sh:
results must not depend on the tiling).  Run under torchrun with N ranks:
every rank integrates its tile; rank 0 also integrates the whole domain on its GPU as ONE tile and
compares the gathered interiors bit-for-bit.
    torchrun --nproc-per-node 2 tools/mgpu_check.py --tiles 2 1 --grid 96 40 30 --steps 6
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import roms_b200 as rb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, nargs=2, default=[2, 1])
    ap.add_argument("--grid", type=int, nargs=3, default=[96, 40, 30])
    ap.add_argument("--app", type=int, default=rb.APP_BENCHMARK)
    ap.add_argument("--steps", type=int, default=6)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    assert world == a.tiles[0] * a.tiles[1]
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.frombuffer(bytearray(rb.comm_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(idt, 0)
    cfg = rb.default_config(a.app, *a.grid)
    cfg.NtileI, cfg.NtileJ = a.tiles
    d = rb.Driver(cfg, tile=rank, device=local)
    d.comm_init(rank, world, bytes(idt.cpu().numpy().tobytes()))
    # NVLink peer mailboxes: all-gather the CUDA IPC handles (what MPI_Allgather does in a Fortran host)
    hnd = torch.frombuffer(bytearray(d.p2p_handle()), dtype=torch.uint8).cuda()
    allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(allh, hnd)
    d.p2p_connect(b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh), world)
    dist.barrier()
    d.run(a.steps)
    d.ctx.sync()
    b = d.bounds()
    d.ctx._bounds = b
    N = cfg.N
    fields = [("zeta", 1, 1, 1), ("zeta", 2, 1, 1), ("u", 1, 1, N), ("u", 2, 1, N), ("v", 1, 1, N), ("v", 2, 1, N),
              ("t", 1, 1, N), ("t", 2, 1, N), ("t", 1, 2, N), ("t", 2, 2, N), ("ubar", 1, 1, 1), ("vbar", 1, 1, 1),
              ("wvel", 1, 1, N + 1), ("Akv", 1, 1, N + 1), ("W", 1, 1, N + 1)]
    mine = {(n, l, m): d.ctx.download_interior(n, l, m, nk) for n, l, m, nk in fields}
    box = (b.Istr, b.Iend, b.Jstr, b.Jend)
    gathered = [None] * world
    dist.all_gather_object(gathered, (box, mine))
    diag = d.run(1, host_forcing=True)
    dfull = d.ctx.diag_last()
    ok = True
    if rank == 0:
        one = rb.default_config(a.app, *a.grid)
        s = rb.Driver(one, device=local)
        s.run(a.steps)
        s.ctx._bounds = s.bounds()
        worst = 0.0
        for n, l, m, nk in fields:
            ref = s.ctx.download_interior(n, l, m, nk)
            for (i0, i1, j0, j1), part in gathered:
                got = part[(n, l, m)]
                exp = ref[:, j0 - 1:j1, i0 - 1:i1]
                if not np.array_equal(got, exp):
                    ok = False
                    worst = max(worst, float(np.max(np.abs(got - exp))))
                    print("MISMATCH", n, l, m, "tile box", (i0, i1, j0, j1), "max abs diff", float(np.max(np.abs(got - exp))), flush=True)
        dg = s.run(1, host_forcing=True)
        sfull = s.ctx.diag_last()
        # diag across tiles: sums agree to round-off (mp_reduce order), maxima and the MAXLOC location exactly
        if not (np.allclose(dfull[:3], sfull[:3], rtol=1e-13, atol=0) and np.array_equal(dfull[3:], sfull[3:])):
            ok = False
            print("DIAG MISMATCH multi", dfull, "single", sfull, flush=True)
        print("tiling %dx%d vs 1x1 after %d steps: %s ; diag multi %s single %s" % (a.tiles[0], a.tiles[1], a.steps,
              "BIT-IDENTICAL" if ok else "DIFFERENT (max %.3e)" % worst, diag, dg), flush=True)
        s.finalize()
    d.finalize()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
