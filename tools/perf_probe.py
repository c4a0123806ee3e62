"""Quick GPU timing probe (development tool): per-config step time, iteration counts, class breakdown."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
from incompressible_flow_b200 import workloads, abi

NAMES = ["darcy", "minres", "prec", "sat", "coeff", "dot", "misc", "comm"]

def probe(name, cfg, steps=2, breakdown=True, **kw):
    if os.environ.get('PROBE_MG'):
        kw['schur_prec'] = abi.SCHUR_MG
    d, Kx, Ky, Kz, phi = workloads.build(**cfg, **kw)
    n = d.nx * d.ny * (d.nz if d.dim == 3 else 1)
    with pkg.Solver(d, Kx, Ky, Kz, phi) as s:
        s.set_field(abi.FIELD_SATURATION, np.zeros(n))
        for ti in range(steps):
            t0 = time.time()
            _, st = s.pressure_solve(None, want_x=False)
            t1 = time.time()
            s.saturation_advance(None, None, ti + 1, 100.0, 1.0)
            t2 = time.time()
            print(f"{name}: step {ti+1} iters={st.iterations} solve={t1-t0:.4f}s ({1e3*(t1-t0)/max(st.iterations,1):.3f} ms/it) "
                  f"sat={t2-t1:.4f}s cellsteps/s={n/(t2-t0):.3e}", flush=True)
        if breakdown:
            s.profile_enable(True)
            _, st = s.pressure_solve(None, want_x=False)
            s.saturation_advance(None, None, 9, 100.0, 1.0)
            ms, cnt = s.profile_read()
            s.profile_enable(False)
            print("   breakdown (sync'd per launch, ms total / launches): " +
                  ", ".join(f"{NAMES[i]}={ms[i]:.1f}/{cnt[i]}" for i in range(8) if cnt[i]), flush=True)
        for kid, kname, bpc in ((abi.KERNEL_DARCY_APPLY, "K1 darcy_apply", 88 if d.dim == 3 else 64),
                                (abi.KERNEL_SAT_SUBSTEP, "K6 sat_substep", 48 if d.dim == 3 else 40)):
            for fl in (0, 1):
                a, m = s.time_kernel(kid, 10, fl)
                print(f"   {kname} flush={fl}: avg {a*1e3:.1f} us min {m*1e3:.1f} us -> {bpc*n/(a*1e-3)/1e9:.0f} GB/s algorithmic", flush=True)

if __name__ == "__main__":
    which = sys.argv[1:] or ["C1", "L256", "C3"]
    cfgs = {
        "C1": dict(dim=2, nx=64, ny=64, K="homogeneous"),
        "L256": dict(dim=2, nx=256, ny=256, K="lognormal"),
        "C2": dict(dim=2, nx=512, ny=512, K="lognormal"),
        "C3": dict(dim=3, nx=128, ny=128, nz=32, K="lognormal-extruded"),
        "C4": dict(dim=3, nx=60, ny=220, nz=85, K="lognormal", sigma=2.0, kz_ratio=0.1),
        "S256": dict(dim=3, nx=256, ny=256, nz=256, K="lognormal"),
        "S384": dict(dim=3, nx=384, ny=384, nz=384, K="lognormal"),
    }
    for w in which:
        probe(w, cfgs[w], steps=1 if w.startswith("S") else 2, breakdown=not w.startswith("S"))
