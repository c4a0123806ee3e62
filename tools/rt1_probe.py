"""GPU timing probe for the RT1/Q1 path (development tool): step time, iterations, class breakdown, K1 roofline.
K1 algorithmic bytes (SURVEY 8d, S in P0): reads x (24 u + 8 p), m_d (3), writes y (32) per 3-D cell = 67 doubles."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
from incompressible_flow_b200 import workloads, abi

NAMES = ["darcy", "minres", "prec", "sat", "coeff", "dot", "misc", "comm"]


def probe(n3, mg=True, steps=2, two_d=False):
    cfg = dict(dim=2, nx=n3, ny=n3) if two_d else dict(dim=3, nx=n3, ny=n3, nz=n3)
    d, Kx, Ky, Kz, phi = workloads.build(K="lognormal", order=1, order_s=0,
                                         schur_prec=abi.SCHUR_MG if mg else abi.SCHUR_SGS_LEX, **cfg)
    n = d.nx * d.ny * (d.nz if d.dim == 3 else 1)
    tag = f"RT1 {'2-D' if two_d else '3-D'} n={n3} {'MG' if mg else 'SGS'}"
    with pkg.Solver(d, Kx, Ky, Kz, phi) as s:
        s.set_field(abi.FIELD_SATURATION, np.zeros(n))
        for ti in range(steps):
            t0 = time.time()
            _, st = s.pressure_solve(None, want_x=False)
            t1 = time.time()
            s.saturation_advance(None, None, ti + 1, 100.0, 1.0)
            bad, mx = s.guards()
            t2 = time.time()
            print(f"{tag}: step {ti+1} iters={st.iterations} solve={t1-t0:.4f}s ({1e3*(t1-t0)/max(st.iterations,1):.3f} ms/it) "
                  f"sat={t2-t1:.4f}s step={t2-t0:.4f}s cellsteps/s={n/(t2-t0):.3e} Smax={mx:.4f}", flush=True)
        s.profile_enable(True)
        _, st = s.pressure_solve(None, want_x=False)
        ms, cnt = s.profile_read()
        s.profile_enable(False)
        print("   breakdown (sync'd per launch, ms total / launches): " +
              ", ".join(f"{NAMES[i]}={ms[i]:.1f}/{cnt[i]}" for i in range(8) if cnt[i]), flush=True)
        bpc = (67 if d.dim == 3 else 26) * 8
        for kid, kname, b in ((abi.KERNEL_DARCY_APPLY, "K1 rt1 darcy_apply", bpc * n),
                              (abi.KERNEL_MINRES_UPDATE, "K2 minres b+c", 13 * 8 * (s.nv + s.np_)),
                              (abi.KERNEL_PREC_APPLY, "K3 prec", 0)):
            for fl in (0, 1):
                a, m = s.time_kernel(kid, 5, fl)
                print(f"   {kname} flush={fl}: avg {a*1e3:.1f} us min {m*1e3:.1f} us" +
                      (f" -> {b/(a*1e-3)/1e9:.0f} GB/s algorithmic" if b else ""), flush=True)


if __name__ == "__main__":
    for a in sys.argv[1:]:
        kind, n3 = a.split(":")
        probe(int(n3), mg=kind.startswith("mg"), two_d=kind.endswith("2"))
