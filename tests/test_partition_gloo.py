"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): slab ownership is a disjoint cover of every
native vector, and the zero-and-sum gather used by the drivers reassembles full arrays on all ranks."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, dim, nx, ny, nz, order, mg, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    pkg = load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = nx * ny * (nz if dim == 3 else 1)
    nlay = nz if dim == 3 else ny
    slab = pkg.partition.slab_range_mg(nlay, rank, world, (nx, ny) if dim == 3 else (nx,)) if mg else None
    darcy, cells = pkg.partition.owned_ranges(dim, nx, ny, nz, rank, world, order=order, slab=slab)
    plane = n // nlay
    if order == 0:
        nv = (nx + 1) * ny * (nz if dim == 3 else 1) + (nx * (ny + 1) * nz if dim == 3 else 0) + plane * (nlay + 1)
        np_ = n
    else:   # RT1 node grids: 2n+1 closed nodes along the component, 2n open nodes across; 2^dim L2 dofs per cell
        k = (2 * nz) if dim == 3 else 1
        nv = (2 * nx + 1) * 2 * ny * k + 2 * nx * (2 * ny + 1) * k + (4 * nx * ny * (2 * nz + 1) if dim == 3 else 0)
        np_ = n * 2 ** dim
    ref_x = np.arange(nv + np_, dtype=np.float64) + 1.0
    ref_s = np.arange(n, dtype=np.float64) * 0.5 + 3.0
    ok = True
    for ref, own in ((ref_x, darcy), (ref_s, cells)):
        a = np.full_like(ref, -7.0)           # stale values everywhere ...
        for lo, hi in own:
            a[lo:hi] = ref[lo:hi]             # ... except what this rank computed
        m = np.zeros(a.shape[0], dtype=bool)
        for lo, hi in own:
            m[lo:hi] = True
        a[~m] = 0.0
        t = torch.from_numpy(a)
        dist.all_reduce(t)
        ok = ok and np.array_equal(t.numpy(), ref)
        cnt = torch.from_numpy(m.astype(np.float64))
        dist.all_reduce(cnt)
        ok = ok and bool((cnt.numpy() == 1.0).all())   # every dof owned by exactly one rank
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,dim,nx,ny,nz,order,mg", [(2, 3, 5, 4, 7, 0, 0), (3, 2, 6, 8, 1, 0, 0), (2, 2, 4, 3, 1, 0, 0),
                                                         (2, 3, 3, 4, 5, 1, 0), (3, 2, 5, 7, 1, 1, 0), (2, 3, 4, 3, 9, 1, 1),
                                                         (3, 2, 6, 21, 1, 0, 1)])
def test_slab_gather(world, dim, nx, ny, nz, order, mg):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, dim, nx, ny, nz, order, mg, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res) and len(res) == world


def test_slab_ranges_balanced(pkg):
    """85 layers on 8 GPUs -> 11,11,11,11,11,10,10,10; ranges are contiguous and cover [0, n)."""
    r = [pkg.partition.slab_range(85, k, 8) for k in range(8)]
    assert [hi - lo for lo, hi in r] == [11, 11, 11, 11, 11, 10, 10, 10]
    assert r[0][0] == 0 and r[-1][1] == 85 and all(r[i][1] == r[i + 1][0] for i in range(7))


@pytest.mark.parametrize("dim,n,order,order_s", [(2, (6, 9, 1), 0, 0), (3, (4, 3, 8), 0, 1), (2, (5, 8, 1), 0, 1),
                                                 (3, (3, 4, 6), 1, 0), (3, (3, 4, 6), 1, 1), (2, (5, 8, 1), 1, 1)])
def test_owned_ranges_tile_the_vectors(pkg, dim, n, order, order_s):
    """The slabs' owned index ranges are disjoint and cover the Darcy vector and the saturation vector of every space
    triple (DG-Q1 saturation: 2^dim dofs per cell whatever the velocity order)."""
    nx, ny, nz = n
    world = 3
    ncell = nx * ny * (nz if dim == 3 else 1)
    ns = ncell * (2 ** dim if order_s else 1)
    if order == 0:
        nv = (nx + 1) * ny * (nz if dim == 3 else 1) + nx * (ny + 1) * (nz if dim == 3 else 1) + (nx * ny * (nz + 1) if dim == 3 else 0)
        npr = ncell
    else:
        if dim == 2:
            nv = (2 * nx + 1) * 2 * ny + 2 * nx * (2 * ny + 1)
        else:
            nv = (2 * nx + 1) * 2 * ny * 2 * nz + 2 * nx * (2 * ny + 1) * 2 * nz + 2 * nx * 2 * ny * (2 * nz + 1)
        npr = ncell * 2 ** dim
    cov_x, cov_s = np.zeros(nv + npr, int), np.zeros(ns, int)
    for r in range(world):
        dx, ds = pkg.partition.owned_ranges(dim, nx, ny, nz, r, world, order=order, order_s=order_s)
        for a, b in dx:
            cov_x[a:b] += 1
        for a, b in ds:
            cov_s[a:b] += 1
    assert np.all(cov_x == 1) and np.all(cov_s == 1)


def test_slab_ranges_multigrid_aligned(pkg):
    """Multi-GPU multigrid mode: slabs in units of 2^g layers, g = the first replicated level (aggregates of the
    distributed levels never straddle GPUs); small levels run replicated instead of exchanging halos."""
    P = pkg.partition
    assert P.mg_first_replicated_level(256, 8, (256, 256)) == 2          # 64^3 = 2^18 cells <= 2^19
    r = [P.slab_range_mg(256, k, 8, (256, 256)) for k in range(8)]
    assert [hi - lo for lo, hi in r] == [32] * 8
    assert P.mg_first_replicated_level(85, 8, (60, 220)) == 1            # SPE10 shape: level 1 = 30x110x43
    r = [P.slab_range_mg(85, k, 8, (60, 220)) for k in range(8)]
    assert [hi - lo for lo, hi in r] == [12, 12, 12, 10, 10, 10, 10, 9]
    assert r[0][0] == 0 and r[-1][1] == 85 and all(r[i][1] == r[i + 1][0] for i in range(7))
    assert all(lo % 2 == 0 for lo, _ in r)
    # no plane dims: the alignment cap itself (as many distributed levels as the layer count allows)
    r = [P.slab_range_mg(85, k, 8) for k in range(8)]
    assert [hi - lo for lo, hi in r] == [16, 16, 16, 8, 8, 8, 8, 5]
    assert P.slab_range_mg(21, 0, 2, (33,)) == (0, 12) and P.slab_range_mg(21, 1, 2, (33,)) == (12, 21)
