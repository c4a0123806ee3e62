"""Per-operator bit comparison of the DG-Q1 GPU path with the oracle's canonical mirror (development tool)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package, load_oracle
pkg = load_package(); orc = load_oracle()
from incompressible_flow_b200 import workloads, abi


def cmp(name, a, b):
    eq = np.array_equal(a, b)
    nb = np.linalg.norm(b)
    print(f"   {name:14s} equal={eq} rel={np.linalg.norm(a - b) / (nb if nb else 1):.2e} ndiff={int((a != b).sum())}/{a.size}")
    return eq


for cfg in (dict(dim=2, nx=9, ny=7), dict(dim=3, nx=6, ny=5, nz=4)):
    for tp in (True, False):
        d, Kx, Ky, Kz, phi = workloads.build(K="lognormal", order=1, order_s=1, two_phase=tp, **cfg)
        n = d.nx * d.ny * (d.nz if d.dim == 3 else 1)
        phi = np.random.default_rng(7).uniform(0.1, 0.4, n)
        o = orc.Oracle(d, Kx, Ky, Kz, phi); o.set_arith(1)
        s = pkg.Solver(d, Kx, Ky, Kz, phi)
        rng = np.random.default_rng(3)
        S = rng.uniform(0.05, 0.95, o.ns); x = rng.standard_normal(o.nv + o.np_)
        s.update_coefficients(S); o.update_coefficients(S)
        print(cfg, "two_phase", tp)
        yg, yo = s.darcy_apply(x), o.darcy_apply(x)
        cmp("K1 u-part", yg[:o.nv], yo[:o.nv]); cmp("K1 p-part", yg[o.nv:], yo[o.nv:])
        zg, zo = s.prec_apply(x), o.prec_apply(x)
        cmp("prec u (1/Md)", zg[:o.nv], zo[:o.nv]); cmp("prec p (SGS)", zg[o.nv:], zo[o.nv:])
        xo, sto, _ = o.pressure_solve(S); xg, stg = s.pressure_solve(S)
        print(f"   solve iters gpu={stg.iterations} oracle={sto.iterations} final {stg.final_norm == sto.final_norm}")
        cmp("solve x", xg, xo)
        So, _ = o.saturation_advance(xo[:o.nv], S, 0.4, 0.05)
        Sg, _ = s.saturation_advance(xo[:o.nv], S, 1, 0.4, 0.05)
        cmp("saturation", Sg, So)
        s.close()
