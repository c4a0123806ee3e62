"""Generates tests/golden/*.npz: small regression fixtures produced by the CPU oracle in canonical
arithmetic.  They are NOT reference output (the reference cannot be built here and ships no golden vectors,
SURVEY 8c); they pin the oracle and the GPU path against silent drift between rounds.
Usage: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from __graft_entry__ import load_oracle, load_package  # noqa: E402

CASES = {
    "q5spot_2d_16x12": dict(dim=2, nx=16, ny=12, K="lognormal", sigma=1.0),
    "box_3d_7x6x5": dict(dim=3, nx=7, ny=6, nz=5, K="lognormal", sigma=1.0),
    # the reference's default spaces RT1 / L2-Q1 / DG_0 (bit-exact on the GPU)
    "rt1_2d_10x8": dict(dim=2, nx=10, ny=8, K="lognormal", sigma=1.0, order=1, order_s=0),
    "rt1_3d_5x4x4": dict(dim=3, nx=5, ny=4, nz=4, K="lognormal", sigma=1.0, order=1, order_s=0),
    # RT1 / L2-Q1 / DG-Q1 (config C5's spaces): canonical arithmetic = oracle/orc_q1.inc, bit-exact on the GPU too
    "rt1q_2d_8x6": dict(dim=2, nx=8, ny=6, K="lognormal", sigma=1.0, order=1, order_s=1),
}
# (global dt, local dt) of the two IMPES steps; DG-Q1 with forward Euler needs a small step
DT = {"rt1q_2d_8x6": (1.0, 0.05), "rt1_3d_5x4x4": (4.0, 0.5)}   # small meshes: stay before breakthrough
EXACT_ON_GPU = {name: True for name in CASES}

if __name__ == "__main__":
    load_package()
    orc = load_oracle()
    from incompressible_flow_b200 import workloads
    for name, cfg in CASES.items():
        built = workloads.build(**cfg)
        o = orc.Oracle(*built)
        o.set_arith(1)
        dtg, dtl = DT.get(name, (10.0, 0.5))
        S, its, xs, Ss = np.zeros(o.ns), [], [], []
        for ti in range(2):
            x, st, rc = o.pressure_solve(S)
            S, nsub = o.saturation_advance(x[:o.nv], S, dtg, dtl)
            its.append(st.iterations)
            xs.append(x)
            Ss.append(S)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), iterations=np.array(its), x=np.array(xs), S=np.array(Ss),
                            rhs=o.rhs())
        print(name, its, "S max", S.max())
