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
}

if __name__ == "__main__":
    load_package()
    orc = load_oracle()
    from incompressible_flow_b200 import workloads
    for name, cfg in CASES.items():
        built = workloads.build(**cfg)
        o = orc.Oracle(*built)
        o.set_arith(1)
        n = o.n_cells
        S, its, xs, Ss = np.zeros(n), [], [], []
        for ti in range(2):
            x, st, rc = o.pressure_solve(S)
            S, nsub = o.saturation_advance(x[:o.nv], S, 10.0, 0.5)
            its.append(st.iterations)
            xs.append(x)
            Ss.append(S)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), iterations=np.array(its), x=np.array(xs), S=np.array(Ss),
                            rhs=o.rhs())
        print(name, its, "S max", S.max())
