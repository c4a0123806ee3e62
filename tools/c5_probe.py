"""GPU timing probe for the C5 path (RT1 / L2-Q1 / DG-Q1, multigrid mode; development tool): step time, iterations,
class breakdown and isolated kernel timings with L2 flushed.  usage: c5_probe.py <n> [order_s] [steps] [sat_source]"""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
from incompressible_flow_b200 import workloads, abi

NAMES = ["darcy", "minres", "prec", "sat", "coeff", "dot", "misc", "comm"]
n3 = int(sys.argv[1])
order_s = int(sys.argv[2]) if len(sys.argv) > 2 else 1
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
src = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
d, Kx, Ky, Kz, phi = workloads.build(dim=3, nx=n3, ny=n3, nz=n3, K="lognormal", order=1, order_s=order_s,
                                     schur_prec=abi.SCHUR_MG, saturation_source=src)
n = n3 ** 3
tag = f"RT1/Q1/DG{order_s} n={n3} MG src={src}"
with pkg.Solver(d, Kx, Ky, Kz, phi) as s:
    s.set_field(abi.FIELD_SATURATION, np.zeros(s.ns))
    for ti in range(steps):
        t0 = time.time()
        _, st = s.pressure_solve(None, want_x=False)
        t1 = time.time()
        s.saturation_advance(None, None, ti + 1, 25.0, 0.25)
        bad, mx = s.guards()
        t2 = time.time()
        print(f"{tag}: step {ti+1} iters={st.iterations} solve={t1-t0:.4f}s ({1e3*(t1-t0)/max(st.iterations,1):.3f} ms/it) "
              f"sat={t2-t1:.4f}s step={t2-t0:.4f}s cellsteps/s={n/(t2-t0):.3e} Smax={mx:.4f} bad={bad}", flush=True)
    if os.environ.get("PROBE_NO_DETAIL"):
        sys.exit(0)
    s.profile_enable(True)
    _, st = s.pressure_solve(None, want_x=False)
    s.saturation_advance(None, None, 99, 25.0, 0.25)
    ms, cnt = s.profile_read()
    s.profile_enable(False)
    print("   breakdown (sync'd per launch, ms total / launches): " +
          ", ".join(f"{NAMES[i]}={ms[i]:.1f}/{cnt[i]}" for i in range(8) if cnt[i]), flush=True)
    ndof = s.nv + s.np_
    k1b = (600 if order_s else 536) * n
    k6b = (328 if order_s else 48) * n
    for kid, kname, b in ((abi.KERNEL_DARCY_APPLY, "K1 darcy_apply", k1b),
                          (abi.KERNEL_MINRES_UPDATE, "K2 minres b+c", 13 * 8 * ndof),
                          (abi.KERNEL_SAT_SUBSTEP, "K6 sat stage", k6b),
                          (abi.KERNEL_PREC_APPLY, "K3 prec (jacobi_u + V-cycle)", 0),
                          (abi.KERNEL_COEFF, "K4 coeff", 0)):
        for fl in (0, 1):
            a, m = s.time_kernel(kid, 5, fl)
            print(f"   {kname} flush={fl}: avg {a*1e3:.1f} us min {m*1e3:.1f} us" +
                  (f" -> {b/(a*1e-3)/1e9:.0f} GB/s algorithmic ({b/(a*1e-3)/1e9/6548.2:.3f})" if b else ""), flush=True)
