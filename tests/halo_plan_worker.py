"""Worker of test_halo_plan_gloo_world2: the exchange plan of roms_b200/csrc/k_halo.cu on numpy + gloo."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import roms_b200 as rb

ti, tj = int(sys.argv[1]), int(sys.argv[2])
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
Lm, Mm, w = 24, 12, 3
b = rb.tile_bounds(Lm, Mm, 4, NtileI=ti, NtileJ=tj, tile=rank, distributed=w)
nb = (C.c_int * 4)()
L = rb.Lib.get().L
L.roms_b200_tile_neighbors.argtypes = [C.POINTER(rb.Bounds), C.c_void_p]
L.roms_b200_tile_neighbors(C.byref(b), nb)
W, E, S, N = list(nb)
ni, nj = b.UBi - b.LBi + 1, b.UBj - b.LBj + 1


def G(i, j):                       # global analytic field, periodic in i
    iw = (i - 1) % Lm + 1
    return iw + 1000.0 * j


A = np.full((nj, ni), np.nan)
for j in range(max(b.LBj, 0), min(b.UBj, Mm + 1) + 1):          # physical rows incl. wall rows
    for i in range(b.Istr, b.Iend + 1):
        if b.Jstr - (1 if b.Southern_Edge else 0) <= j <= b.Jend + (1 if b.Northern_Edge else 0):
            A[j - b.LBj, i - b.LBi] = G(i, j)
if ti == 1:                        # single tile in i: periodic images are local (kernels' st())
    for i in list(range(b.LBi, b.Istr)) + list(range(b.Iend + 1, b.UBi + 1)):
        A[:, i - b.LBi] = A[:, ((i - 1) % Lm + 1) - b.LBi]


def swap(lo, hi, send_lo, send_hi):
    recv_lo, recv_hi = np.empty_like(send_lo), np.empty_like(send_hi)
    ops = []
    if lo >= 0:
        ops += [dist.P2POp(dist.isend, torch.from_numpy(send_lo.copy()), lo), dist.P2POp(dist.irecv, torch.from_numpy(recv_lo), lo)]
    if hi >= 0:
        ops += [dist.P2POp(dist.isend, torch.from_numpy(send_hi.copy()), hi), dist.P2POp(dist.irecv, torch.from_numpy(recv_hi), hi)]
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    if lo >= 0 and lo == hi:       # two tiles on a periodic axis: same peer both ways -> first message is its LOW strip
        recv_lo, recv_hi = recv_hi, recv_lo
    return recv_lo, recv_hi


# phase 0: W/E strips over all rows
ci0, ci1 = b.Istr - b.LBi, b.Iend - b.LBi
rl, rh = swap(W, E, A[:, ci0:ci0 + w], A[:, ci1 - w + 1:ci1 + 1])
if W >= 0:
    A[:, ci0 - w:ci0] = rl
if E >= 0:
    A[:, ci1 + 1:ci1 + 1 + w] = rh
# phase 1: S/N strips over the full i-range (so corners travel)
cj0, cj1 = b.Jstr - b.LBj, b.Jend - b.LBj
rl, rh = swap(S, N, A[cj0:cj0 + w, :], A[cj1 - w + 1:cj1 + 1, :])
if S >= 0:
    A[cj0 - w:cj0, :] = rl
if N >= 0:
    A[cj1 + 1:cj1 + 1 + w, :] = rh
bad = 0
for j in range(max(b.LBj, 1), min(b.UBj, Mm) + 1):
    for i in range(b.LBi, b.UBi + 1):
        if b.UBi - i >= 0 and i - b.LBi >= 0 and abs(i - (b.Istr + b.Iend) / 2) <= (b.Iend - b.Istr) / 2 + w:
            if not A[j - b.LBj, i - b.LBi] == G(i, j):
                bad += 1
t = torch.tensor([bad])
dist.all_reduce(t)
if rank == 0:
    print("HALO_OK" if int(t) == 0 else "HALO_BAD %d" % int(t))
dist.destroy_process_group()
sys.exit(0 if int(t) == 0 else 1)
