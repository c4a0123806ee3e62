"""GPU parity tests for `-ov 0 -op 0 -os 1` (RT0 / P0 velocity-pressure, saturation in DG-Q1) through the C ABI against
the CPU oracle's assembled-CSR formulation (arith 0: element matrices by MFEM's quadrature rules, entry by entry).

The GPU stores the three entries of every 2x2 RT0 element block (k0q_coeff: the 3^3 / 2^2-point rule of order
OrderW + 2), applies K1 with the RT0 kernel, and feeds the DG-Q1 stage kernel with records formed from the RT0 face
fluxes.  No canonical mirror exists for this triple, so everything is compared to round-off (fp64, written here on
purpose): operator applications <= 1e-12 relative L2; MINRES solves: iteration counts within +-2, solution within the
Krylov drift bound 1e-6 at rtol 1e-6 and <= 1e-10 at a tight tolerance."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    dict(dim=2, nx=9, ny=7),
    dict(dim=2, nx=24, ny=20),
    dict(dim=3, nx=6, ny=5, nz=4),
    dict(dim=3, nx=3, ny=17, nz=5),
]


def rel(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


def make(pkg, orc, cfg, K="lognormal", **kw):
    from incompressible_flow_b200 import workloads
    d, Kx, Ky, Kz, phi = workloads.build(K=K, sigma=1.0, order=0, order_s=1, **cfg, **kw)
    assert (d.order_v, d.order_p, d.order_s) == (0, 0, 1)
    n = d.nx * d.ny * (d.nz if d.dim == 3 else 1)
    phi = np.random.default_rng(7).uniform(0.1, 0.4, n)
    o = orc.Oracle(d, Kx, Ky, Kz, phi)
    o.set_arith(0)
    s = pkg.Solver(d, Kx, Ky, Kz, phi)
    return d, o, s, n


@pytest.mark.parametrize("cfg", CASES)
def test_rt0q_dims_and_rhs(pkg, orc, native_lib, cfg):
    d, o, s, n = make(pkg, orc, cfg)
    assert (s.nv, s.np_, s.ns) == (o.nv, o.np_, o.ns) and s.ns == n * 2 ** d.dim and s.np_ == n
    assert np.array_equal(s.rhs(), o.rhs())
    assert "RT0/L2_0/DG_1" in s.describe()
    s.close()


@pytest.mark.parametrize("cfg", CASES)
@pytest.mark.parametrize("two_phase", [True, False])
def test_rt0q_darcy_apply_and_prec(pkg, orc, native_lib, cfg, two_phase):
    d, o, s, n = make(pkg, orc, cfg, two_phase=two_phase)
    rng = np.random.default_rng(3)
    S = rng.uniform(0.05, 0.95, o.ns)          # rough inside the cells on purpose: exercises the quadrature
    x = rng.standard_normal(o.nv + o.np_)
    s.update_coefficients(S)
    o.update_coefficients(S)
    assert rel(s.darcy_apply(x), o.darcy_apply(x)) < 1e-12
    assert rel(s.prec_apply(x), o.prec_apply(x)) < 1e-12
    s.close()


@pytest.mark.parametrize("cfg", CASES)
@pytest.mark.parametrize("mg", [False, True])
def test_rt0q_pressure_solve(pkg, orc, native_lib, cfg, mg):
    from incompressible_flow_b200 import abi
    d, o, s, n = make(pkg, orc, cfg, schur_prec=abi.SCHUR_MG if mg else abi.SCHUR_SGS_LEX)
    S = np.random.default_rng(5).uniform(0.0, 0.9, o.ns)
    xo, sto, _ = o.pressure_solve(S)
    xg, stg = s.pressure_solve(S)
    assert stg.converged == 1 and sto.converged == 1
    assert abs(stg.iterations - sto.iterations) <= 2
    assert rel(xg[:o.nv], xo[:o.nv]) < 1e-6
    s.close()


def test_rt0q_tight_solve_meets_1e10(pkg, orc, native_lib):
    """north_star tolerance: at a tight MINRES tolerance GPU and oracle agree to 1e-10 in velocity and pressure."""
    d, o, s, n = make(pkg, orc, dict(dim=3, nx=6, ny=5, nz=4), rtol=1e-13, maxiter=4000)
    S = np.random.default_rng(9).uniform(0.0, 0.9, o.ns)
    xo, _, _ = o.pressure_solve(S)
    xg, _ = s.pressure_solve(S)
    pm = lambda a: a - a.mean()
    assert rel(xg[:o.nv], xo[:o.nv]) < 1e-10
    assert rel(pm(xg[o.nv:]), pm(xo[o.nv:])) < 1e-10
    s.close()


@pytest.mark.parametrize("cfg", CASES)
@pytest.mark.parametrize("two_phase", [True, False])
def test_rt0q_saturation_substeps(pkg, orc, native_lib, cfg, two_phase):
    """DG-Q1 upwind stages with the RT0 velocity frozen: volume term + face integrals with per-point upwinding."""
    d, o, s, n = make(pkg, orc, cfg, two_phase=two_phase)
    S0 = np.random.default_rng(11).uniform(0.0, 0.6, o.ns)
    xo, _, _ = o.pressure_solve(S0)
    So, nso = o.saturation_advance(xo[:o.nv], S0, 0.4, 0.05)
    Sg, nsg = s.saturation_advance(xo[:o.nv], S0, 1, 0.4, 0.05)
    assert nso == nsg == 8
    assert rel(Sg, So) < 1e-12
    bad, mx = s.guards()
    assert bad == 0 and abs(mx - So.max()) <= 1e-12 * abs(So.max())
    s.close()


def test_rt0q_impes_steps(pkg, orc, native_lib):
    """two full IMPES steps (state resident on the GPU between the calls) against the oracle"""
    d, o, s, n = make(pkg, orc, dict(dim=3, nx=8, ny=6, nz=5), rtol=1e-12, maxiter=4000)
    from incompressible_flow_b200 import abi
    So = np.zeros(o.ns)
    s.set_field(abi.FIELD_SATURATION, So)
    for ti in (1, 2):
        xo, _, _ = o.pressure_solve(So)
        So, _ = o.saturation_advance(xo[:o.nv], So, 2.0, 0.25)
        s.pressure_solve(None, want_x=False)
        s.saturation_advance(None, None, ti, 2.0, 0.25)
    assert rel(s.get_field(abi.FIELD_SATURATION), So) < 1e-9
    s.close()
