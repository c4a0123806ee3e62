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


def owned_ranges(dim, nx, ny, nz, rank, world):
    """Index ranges this rank owns in the native host layout (include/impes_b200.h):
    -> (ranges in a Darcy vector [V | P], ranges in a cell / saturation vector)."""
    if dim == 2:
        n_layers, plane, px, py = ny, nx, nx + 1, 0
        nvx, nvy = (nx + 1) * ny, 0
    else:
        n_layers, plane, px, py = nz, nx * ny, (nx + 1) * ny, nx * (ny + 1)
        nvx, nvy = (nx + 1) * ny * nz, nx * (ny + 1) * nz
    nv = nvx + nvy + plane * (n_layers + 1)
    lo, hi = slab_range(n_layers, rank, world)
    top = 1 if hi == n_layers else 0
    darcy = [(px * lo, px * hi)]
    if py:
        darcy.append((nvx + py * lo, nvx + py * hi))
    darcy.append((nvx + nvy + plane * lo, nvx + nvy + plane * (hi + top)))
    darcy.append((nv + plane * lo, nv + plane * hi))
    return darcy, [(plane * lo, plane * hi)]
