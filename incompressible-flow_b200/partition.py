"""Slab decomposition of the Cartesian mesh across GPUs (one rank per GPU): pure-Python mirror of the rule
the native library uses (csrc/api.cu create_impl), for drivers that gather host arrays across ranks.

Slabs are cut along the slowest axis (z in 3-D, y in 2-D); the first (n % P) ranks get one extra cell
layer (85 layers on 8 GPUs -> 11,11,11,11,11,10,10,10).  A rank owns its cells and, by the rule
"a cell owns its lower faces", the face layers [lo, hi) of the slab axis (the last rank also owns the top
boundary layer).
"""


def slab_range(n_layers, rank, world):
    base, rem = divmod(n_layers, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


MG_REPL_CELLS = 1 << 19


def mg_first_replicated_level(n_layers, world, plane_dims=None, repl=MG_REPL_CELLS):
    """First multigrid level that runs replicated on every GPU (csrc/api.cu create_impl): the first level >= 1 with
    at most `repl` cells, capped by the largest g whose 2^g-layer units still give every GPU a unit.
    plane_dims: (nx,) in 2-D, (nx, ny) in 3-D; None: the cap itself (the round-1 rule)."""
    gmax = 0
    while gmax < 6 and (n_layers >> (gmax + 1)) >= world:
        gmax += 1
    if plane_dims is None or gmax < 1:
        return gmax

    def cells(l):
        dims = list(plane_dims) + [n_layers]
        for _ in range(l):
            dims = [(v + 1) // 2 for v in dims]
        out = 1
        for v in dims:
            out *= v
        return out

    g = 1
    while g < gmax and cells(g) > repl:
        g += 1
    return g


def slab_range_mg(n_layers, rank, world, plane_dims=None, repl=MG_REPL_CELLS):
    """Slabs of the multi-GPU multigrid mode (IMPES_SCHUR_MG, world > 1): cut in units of 2^g layers, g the first
    replicated level (mg_first_replicated_level), so that the aggregates of the distributed levels never straddle
    two GPUs.  256^3 on 8 GPUs -> g = 2 (level 2 = 64^3 runs replicated), 32 layers each; the SPE10 shape
    60x220x85 on 8 GPUs -> g = 1, units of 2 layers -> 12,12,12,10,10,10,10,9."""
    if world == 1:
        return 0, n_layers
    unit = 1 << mg_first_replicated_level(n_layers, world, plane_dims, repl)
    nu = (n_layers + unit - 1) // unit
    base, rem = divmod(nu, world)
    ulo = rank * base + min(rank, rem)
    uhi = ulo + base + (1 if rank < rem else 0)
    return ulo * unit, min(uhi * unit, n_layers)


def owned_ranges(dim, nx, ny, nz, rank, world, order=0, slab=None, order_s=0):
    """Index ranges this rank owns in the native host layout (include/impes_b200.h):
    -> (ranges in a Darcy vector [V | P], ranges in a cell / saturation vector).
    order = 0: RT0/P0; order = 1: RT1/L2-Q1 (node grids of 2n+1 closed / 2n open nodes, 2^dim L2 dofs per cell).
    slab: explicit (lo, hi) cell-layer range (impes_get_dims slab_lo/slab_hi) overriding the balanced rule.
    order_s = 1: DG-Q1 saturation, 2^dim dofs per cell (cell-major)."""
    nsl = 2 ** dim if order_s == 1 else 1
    if order == 1:
        if dim == 2:
            n_layers, px, py, pz, npl = ny, 2 * nx + 1, 0, 2 * nx, 4
            nvx, nvy = (2 * nx + 1) * 2 * ny, 0
        else:
            n_layers, px, py, pz, npl = nz, (2 * nx + 1) * 2 * ny, 2 * nx * (2 * ny + 1), 4 * nx * ny, 8
            nvx, nvy = (2 * nx + 1) * 2 * ny * 2 * nz, 2 * nx * (2 * ny + 1) * 2 * nz
        plane = nx * (ny if dim == 3 else 1)
        nv = nvx + nvy + pz * (2 * n_layers + 1)
        lo, hi = slab if slab is not None else slab_range(n_layers, rank, world)
        top = 1 if hi == n_layers else 0
        darcy = [(px * 2 * lo, px * 2 * hi)]
        if py:
            darcy.append((nvx + py * 2 * lo, nvx + py * 2 * hi))
        darcy.append((nvx + nvy + pz * 2 * lo, nvx + nvy + pz * (2 * hi + top)))
        darcy.append((nv + plane * npl * lo, nv + plane * npl * hi))
        return darcy, [(plane * nsl * lo, plane * nsl * hi)]
    if dim == 2:
        n_layers, plane, px, py = ny, nx, nx + 1, 0
        nvx, nvy = (nx + 1) * ny, 0
    else:
        n_layers, plane, px, py = nz, nx * ny, (nx + 1) * ny, nx * (ny + 1)
        nvx, nvy = (nx + 1) * ny * nz, nx * (ny + 1) * nz
    nv = nvx + nvy + plane * (n_layers + 1)
    lo, hi = slab if slab is not None else slab_range(n_layers, rank, world)
    top = 1 if hi == n_layers else 0
    darcy = [(px * lo, px * hi)]
    if py:
        darcy.append((nvx + py * lo, nvx + py * hi))
    darcy.append((nvx + nvy + plane * lo, nvx + nvy + plane * (hi + top)))
    darcy.append((nv + plane * lo, nv + plane * hi))
    return darcy, [(plane * nsl * lo, plane * nsl * hi)]
