"""Reader of the parity dump a REAL reference build writes when patched with tools/reference_dump.patch
(IMPES_DUMP_DIR=<dir> ./incompressible-flow ...): fp64 x = [V | P] and S per global step in MFEM's dof numbering,
MINRES iteration counts, and per space the table (signed global dof, physical node position, RT basis value at the
node) for every element-local dof.  The reader identifies every MFEM dof geometrically -- component and orientation
from the basis value, grid position from the node -- and returns all vectors in the library's native layout
(include/impes_b200.h), so no rule about MFEM's numbering has to be trusted.  The same identification runs inside the
real-MFEM adapter (incompressible-flow_b200/host/dofmap.cpp).

    d = load_dump(path)      # d.meta, d.K (Kx, Ky, Kz, phi), d.x[ti], d.S[ti], d.iters, d.desc() -> impes descriptor
Also: write_synthetic_dump() builds such a dump from native arrays under an arbitrary (scrambled) numbering, for the
reader's own test (tests/test_reference_dump.py)."""
import os
import struct

import numpy as np

MAGIC = 0x494D5044


def _read_space(path):
    with open(path, "rb") as f:
        magic, ver, dim, vdim, ne, nd = struct.unpack("<6i", f.read(24))
        assert magic == MAGIC and ver == 1, "not an impes dump space file"
        raw, X, sh = [], [], []
        for _ in range(ne):
            (n,) = struct.unpack("<i", f.read(4))
            rec = np.frombuffer(f.read(n * 52), dtype=np.dtype([("raw", "<i4"), ("v", "<f8", 6)]))
            raw.append(rec["raw"].copy())
            X.append(rec["v"][:, :3].copy())
            sh.append(rec["v"][:, 3:].copy())
    return dict(dim=dim, vdim=vdim, ne=ne, nd=nd, raw=raw, X=X, shape=sh)


def dof_map(space, n, size, order):
    """-> (nat, sgn): native index and orientation of every MFEM dof (native = sgn * mfem); mirrors host/dofmap.cpp"""
    dim, rt = space["dim"], space["vdim"] > 0
    k1 = order + 1
    h = [size[d] / n[d] for d in range(dim)]
    ext = [[(n[e] * k1 + 1) if e == d else n[e] * k1 for e in range(dim)] for d in range(dim)]
    off = [0] * dim
    for d in range(1, dim):
        off[d] = off[d - 1] + int(np.prod(ext[d - 1]))
    nat = np.full(space["nd"], -1, dtype=np.int64)
    sgn = np.ones(space["nd"])
    for e in range(space["ne"]):
        raw, X, sh = space["raw"][e], space["X"][e], space["shape"][e]
        g = np.where(raw >= 0, raw, -1 - raw)
        s = np.where(raw >= 0, 1.0, -1.0)
        u = X[:, :dim] / np.array(h)
        if rt:
            comp = np.argmax(np.abs(sh[:, :dim]), axis=1)
            s = s * np.sign(sh[np.arange(len(raw)), comp])
        else:
            comp = np.full(len(raw), -1)
        cell = np.clip(np.floor(u).astype(np.int64), 0, np.array(n[:dim]) - 1)
        loc = ((u - cell) > 0.5).astype(np.int64) if order == 1 else np.zeros_like(cell)
        gi = cell * k1 + loc
        idx = np.zeros(len(raw), dtype=np.int64)
        for k in range(len(raw)):
            if rt:
                d = comp[k]
                gk = gi[k].copy()
                gk[d] = int(round(u[k, d] * k1))
                ex = ext[d]
                idx[k] = off[d] + gk[0] + ex[0] * (gk[1] + (ex[1] * gk[2] if dim == 3 else 0))
            else:
                c = cell[k, 0] + n[0] * (cell[k, 1] + (n[1] * cell[k, 2] if dim == 3 else 0))
                lo = loc[k, 0] + k1 * (loc[k, 1] + (k1 * loc[k, 2] if dim == 3 else 0))
                idx[k] = c * k1 ** dim + lo
        seen = nat[g] >= 0
        assert np.all(nat[g][seen] == idx[seen]) and np.all(sgn[g][seen] == s[seen]), "elements disagree about a shared dof"
        nat[g], sgn[g] = idx, s
    assert np.all(nat >= 0), "a dof belongs to no element"
    assert len(np.unique(nat)) == len(nat), "dof map is not a permutation"
    return nat, sgn


def _read_vec(path):
    with open(path, "rb") as f:
        (n,) = struct.unpack("<q", f.read(8))
        return np.frombuffer(f.read(8 * n), dtype="<f8").copy()


class Dump:
    pass


def load_dump(path):
    d = Dump()
    meta = {}
    for ln in open(os.path.join(path, "meta.txt")):
        t = ln.split()
        if t:
            meta[t[0]] = [float(v) for v in t[1:]] if len(t) > 2 else (float(t[1]) if "." in t[1] or "e" in t[1] else int(t[1]))
    d.meta = meta
    dim = int(meta["dim"])
    n = [int(meta["nx"]), int(meta["ny"]), int(meta["nz"])][:dim]
    size = [float(meta["sx"]), float(meta["sy"]), float(meta["sz"])][:dim]
    maps = {}
    for name, key in (("V", "order_v"), ("P", "order_p"), ("S", "order_s")):
        maps[name] = dof_map(_read_space(os.path.join(path, f"space_{name}.bin")), n, size, int(meta[key]))
    d.maps = maps
    nv = int(meta["n_v"])

    def to_native(vec, name):
        nat, sgn = maps[name]
        out = np.empty_like(vec)
        out[nat] = sgn * vec
        return out

    d.x, d.S = {}, {}
    ti = 1
    while os.path.exists(os.path.join(path, f"x_{ti:06d}.bin")):
        x = _read_vec(os.path.join(path, f"x_{ti:06d}.bin"))
        d.x[ti] = np.concatenate([to_native(x[:nv], "V"), to_native(x[nv:], "P")])
        d.S[ti] = to_native(_read_vec(os.path.join(path, f"S_{ti:06d}.bin")), "S")
        ti += 1
    d.K = {k: (_read_vec(os.path.join(path, f"coef_{k}_000000.bin")) if os.path.exists(os.path.join(path, f"coef_{k}_000000.bin")) else None)
           for k in ("Kx", "Ky", "Kz", "phi")}
    d.iters = []
    ip = os.path.join(path, "iters.txt")
    if os.path.exists(ip):
        for ln in open(ip):
            t = ln.split()
            d.iters.append((int(t[0]), int(t[1]), float(t[2])))

    def desc():
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from __graft_entry__ import load_package
        abi = load_package().abi
        ds = abi.default_desc()
        ds.dim = dim
        ds.nx, ds.ny, ds.nz = int(meta["nx"]), int(meta["ny"]), int(meta["nz"]) if dim == 3 else 1
        ds.sx, ds.sy, ds.sz = float(meta["sx"]), float(meta["sy"]), float(meta["sz"])
        ds.order_v, ds.order_p, ds.order_s = int(meta["order_v"]), int(meta["order_p"]), int(meta["order_s"])
        ds.two_phase = int(meta["two_phase"])
        ds.Sor, ds.Swc, ds.mu_w, ds.mu_o = float(meta["Sor"]), float(meta["Swc"]), float(meta["mu_w"]), float(meta["mu_o"])
        ds.inflow, ds.outflow, ds.saturation_source = float(meta["inflow"]), float(meta["outflow"]), float(meta["sat_source"])
        for w, key in ((ds.injection, "inj"), (ds.production, "pro")):
            w.center[0], w.center[1], w.center[2], w.radius, w.height = meta[key]
        ds.darcy_rtol, ds.darcy_atol, ds.darcy_maxiter = float(meta["darcy_rtol"]), float(meta["darcy_atol"]), int(meta["darcy_maxiter"])
        return ds
    d.desc = desc
    return d


# ----------------------------------------------------------------------------------------------------------------
def write_synthetic_dump(path, desc, K, x_steps, S_steps, iters, seed=1):
    """A dump as the patched reference would write it, from NATIVE arrays, under a scrambled numbering with MFEM-like
    conventions (outward-pointing bases on the lower faces, global orientation = a random sign per RT dof)."""
    os.makedirs(path, exist_ok=True)
    dim = desc.dim
    n = [desc.nx, desc.ny, desc.nz][:dim]
    size = [desc.sx, desc.sy, desc.sz][:dim]
    h = [size[d] / n[d] for d in range(dim)]
    rng = np.random.default_rng(seed)
    g1 = 0.5 / np.sqrt(3.0)

    def nodes(order):
        return ([0.0, 1.0], [0.5]) if order == 0 else ([0.0, 0.5, 1.0], [0.5 - g1, 0.5 + g1])

    perms, signs = {}, {}

    def write_space(name, order, rt):
        k1 = order + 1
        cl, op = nodes(order)
        ext = [[(n[e] * k1 + 1) if e == d else n[e] * k1 for e in range(dim)] for d in range(dim)]
        off = [0] * dim
        for d in range(1, dim):
            off[d] = off[d - 1] + int(np.prod(ext[d - 1]))
        nd = (off[-1] + int(np.prod(ext[-1]))) if rt else int(np.prod(n)) * k1 ** dim
        perm = rng.permutation(nd)
        gs = rng.choice([-1.0, 1.0], nd) if rt else np.ones(nd)
        perms[name], signs[name] = perm, gs
        ne = int(np.prod(n))
        with open(os.path.join(path, f"space_{name}.bin"), "wb") as f:
            f.write(struct.pack("<6i", MAGIC, 1, dim, dim if rt else 0, ne, nd))
            for e in range(ne):
                c = [e % n[0], (e // n[0]) % n[1]] + ([e // (n[0] * n[1])] if dim == 3 else [])
                recs = []
                if rt:
                    for d in range(dim):
                        others = [q for q in range(dim) if q != d]
                        for nn in range(k1 + 1):
                            for tt in np.ndindex(*([k1] * (dim - 1))):
                                ref = [0.0] * dim
                                gi = [0] * dim
                                ref[d], gi[d] = cl[nn], c[d] * k1 + nn
                                for q, t in zip(others, tt):
                                    ref[q], gi[q] = op[t], c[q] * k1 + t
                                nat = off[d] + gi[0] + ext[d][0] * (gi[1] + (ext[d][1] * gi[2] if dim == 3 else 0))
                                lsign = -1.0 if nn == 0 else 1.0
                                raw = perm[nat] if gs[nat] == lsign else -1 - perm[nat]
                                X = [(c[q] + ref[q]) * h[q] for q in range(dim)] + [0.0] * (3 - dim)
                                sh = [0.0, 0.0, 0.0]
                                sh[d] = lsign * h[d] / float(np.prod(h))
                                recs.append((int(raw), X + sh))
                else:
                    for tt in np.ndindex(*([k1] * dim)):
                        t = tt[::-1]   # x fastest
                        nat = e * k1 ** dim + t[0] + k1 * (t[1] + (k1 * t[2] if dim == 3 else 0))
                        X = [(c[q] + op[t[q]]) * h[q] for q in range(dim)] + [0.0] * (3 - dim)
                        recs.append((int(perm[nat]), X + [0.0, 0.0, 0.0]))
                rng.shuffle(recs)   # local dof order is arbitrary too
                f.write(struct.pack("<i", len(recs)))
                for raw, v in recs:
                    f.write(struct.pack("<i6d", raw, *v))
        return nd

    nv = write_space("V", desc.order_v, True)
    np_ = write_space("P", desc.order_p, False)
    ns = write_space("S", desc.order_s, False)

    def from_native(vec, name):
        out = np.empty_like(vec)
        out[perms[name]] = signs[name] * vec
        return out

    def wvec(fname, v):
        with open(os.path.join(path, fname), "wb") as f:
            f.write(struct.pack("<q", len(v)))
            f.write(np.ascontiguousarray(v, dtype="<f8").tobytes())

    for ti, (x, S) in enumerate(zip(x_steps, S_steps), start=1):
        wvec(f"x_{ti:06d}.bin", np.concatenate([from_native(x[:nv], "V"), from_native(x[nv:], "P")]))
        wvec(f"S_{ti:06d}.bin", from_native(S, "S"))
    for k, a in K.items():
        if a is not None:
            wvec(f"coef_{k}_000000.bin", a)
    with open(os.path.join(path, "iters.txt"), "w") as f:
        for it, conv, fn in iters:
            f.write(f"{it} {conv} {fn:.17g}\n")
    inj, pro = desc.injection, desc.production
    with open(os.path.join(path, "meta.txt"), "w") as f:
        f.write(f"dim {dim}\nnx {desc.nx}\nny {desc.ny}\nnz {desc.nz}\nsx {desc.sx!r}\nsy {desc.sy!r}\nsz {desc.sz!r}\n"
                f"order_v {desc.order_v}\norder_p {desc.order_p}\norder_s {desc.order_s}\ntwo_phase {desc.two_phase}\n"
                f"Sor {desc.Sor!r}\nSwc {desc.Swc!r}\nmu_w {desc.mu_w!r}\nmu_o {desc.mu_o!r}\ninflow {desc.inflow!r}\n"
                f"outflow {desc.outflow!r}\nsat_source {desc.saturation_source!r}\n"
                f"inj {inj.center[0]!r} {inj.center[1]!r} {inj.center[2]!r} {inj.radius!r} {inj.height!r}\n"
                f"pro {pro.center[0]!r} {pro.center[1]!r} {pro.center[2]!r} {pro.radius!r} {pro.height!r}\n"
                f"darcy_rtol {desc.darcy_rtol!r}\ndarcy_atol {desc.darcy_atol!r}\ndarcy_maxiter {desc.darcy_maxiter}\n"
                f"n_v {nv}\nn_p {np_}\nn_s {ns}\n")
