"""GPU parity tests for the RT1 / L2-Q1 pressure path (the reference's default orders `-ov 1 -op 1`
[REF src/param.cpp:43]) and the DG_0 saturation step driven by an RT1 velocity, through the C ABI,
against the CPU oracle.  Same tolerance scheme as test_gpu_parity.py:
  * oracle in canonical arithmetic: bit-exact (np.array_equal);
  * oracle in MFEM-like arithmetic (quadrature-assembled CSR -- the independent formulation):
    operators 1e-13 relative, solves: iteration counts within +-2 and the Krylov drift bound.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    dict(dim=2, nx=9, ny=7),
    dict(dim=2, nx=32, ny=32),
    dict(dim=3, nx=6, ny=5, nz=4),
    dict(dim=3, nx=3, ny=70, nz=5),    # more than one 256-cell block, ragged
    dict(dim=2, nx=1, ny=4),           # degenerate: single column
    dict(dim=3, nx=40, ny=40, nz=7),   # > 1024 pencils/8: cluster / cooperative sweep shapes
]


def rel(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


def make(pkg, orc, cfg, K="lognormal", **kw):
    from incompressible_flow_b200 import workloads
    d, Kx, Ky, Kz, phi = workloads.build(K=K, sigma=1.0, order=1, order_s=0, **cfg, **kw)
    rng = np.random.default_rng(7)
    n = d.nx * d.ny * (d.nz if d.dim == 3 else 1)
    phi = rng.uniform(0.1, 0.4, n)
    o = orc.Oracle(d, Kx, Ky, Kz, phi)
    s = pkg.Solver(d, Kx, Ky, Kz, phi)
    return d, o, s, n


@pytest.mark.parametrize("cfg", CASES)
def test_rt1_dims_and_rhs(pkg, orc, native_lib, cfg):
    d, o, s, n = make(pkg, orc, cfg)
    assert (s.nv, s.np_, s.ns) == (o.nv, o.np_, o.ns)
    assert list(s.dims.n_v_comp)[:d.dim] == o.nvc[:d.dim]
    o.set_arith(1)
    assert np.array_equal(s.rhs(), o.rhs())
    o.set_arith(0)
    assert rel(s.rhs(), o.rhs()) < 1e-14
    s.close()


@pytest.mark.parametrize("cfg", CASES)
@pytest.mark.parametrize("two_phase", [True, False])
def test_rt1_darcy_apply_and_prec(pkg, orc, native_lib, cfg, two_phase):
    d, o, s, n = make(pkg, orc, cfg, two_phase=two_phase)
    rng = np.random.default_rng(3)
    S = rng.uniform(0.05, 0.95, n)
    x = rng.standard_normal(o.nv + o.np_)
    s.update_coefficients(S)
    yg = s.darcy_apply(x)
    zg = s.prec_apply(x)
    o.set_arith(0)
    o.update_coefficients(S)
    assert rel(yg, o.darcy_apply(x)) < 1e-13
    assert rel(zg, o.prec_apply(x)) < 1e-12
    o.set_arith(1)
    o.update_coefficients(S)
    assert np.array_equal(yg, o.darcy_apply(x))
    assert np.array_equal(zg, o.prec_apply(x))
    s.close()


@pytest.mark.parametrize("cfg", CASES[:4])
def test_rt1_pressure_solve_bit_exact(pkg, orc, native_lib, cfg):
    d, o, s, n = make(pkg, orc, cfg)
    S = np.random.default_rng(5).uniform(0.0, 0.9, n)
    o.set_arith(1)
    xo, sto, rco = o.pressure_solve(S)
    xg, stg = s.pressure_solve(S)
    assert stg.converged == 1 and sto.converged == 1
    assert stg.iterations == sto.iterations
    assert stg.final_norm == sto.final_norm
    assert np.array_equal(xg, xo)
    o.set_arith(0)
    xm, stm, _ = o.pressure_solve(S)
    assert abs(stg.iterations - stm.iterations) <= 2
    assert rel(xg[:o.nv], xm[:o.nv]) < 1e-6
    s.close()


@pytest.mark.parametrize("cfg", CASES[:4])
@pytest.mark.parametrize("two_phase", [True, False])
def test_rt1_saturation_substeps_bit_exact(pkg, orc, native_lib, cfg, two_phase):
    """DG_0 saturation with an RT1 velocity: face flux = Gauss-weighted sum of the face's RT dofs."""
    d, o, s, n = make(pkg, orc, cfg, two_phase=two_phase)
    o.set_arith(1)
    S0 = np.zeros(n)
    xo, _, _ = o.pressure_solve(S0)
    So, nso = o.saturation_advance(xo[:o.nv], S0, 7.3, 0.5)
    Sg, nsg = s.saturation_advance(xo[:o.nv], S0, 1, 7.3, 0.5)
    assert nso == nsg == 14
    assert np.array_equal(Sg, So)
    s.close()


@pytest.mark.parametrize("cfg", [CASES[1], CASES[2]])
def test_rt1_impes_loop_bit_exact(pkg, orc, native_lib, cfg):
    """3 IMPES steps with resident state on the GPU (NULL host pointers) vs the oracle's loop."""
    from incompressible_flow_b200 import abi
    d, o, s, n = make(pkg, orc, cfg)
    o.set_arith(1)
    So = np.zeros(n)
    s.set_field(abi.FIELD_SATURATION, So)
    its_o, its_g = [], []
    for step in range(3):
        xo, sto, _ = o.pressure_solve(So)
        So, _ = o.saturation_advance(xo[:o.nv], So, 20.0, 1.0)
        _, stg = s.pressure_solve(None, want_x=False)
        s.saturation_advance(None, None, step + 1, 20.0, 1.0)
        its_o.append(sto.iterations)
        its_g.append(stg.iterations)
    assert its_o == its_g
    assert np.array_equal(s.get_field(abi.FIELD_SATURATION), So)
    assert np.array_equal(s.get_field(abi.FIELD_DARCY_X), xo)
    s.close()


def test_rt1_survey_iteration_counts(pkg, orc, native_lib):
    """[SURVEY-CHECK]: 2-D RT1/Q1 homogeneous quarter-five-spot, serial preconditioner,
    rtol 1e-6: 77 / 153 / 287 MINRES iterations at n = 8 / 16 / 32."""
    from incompressible_flow_b200 import workloads
    for nn, want in [(8, 77), (16, 153), (32, 287)]:
        d, Kx, Ky, Kz, phi = workloads.build(dim=2, nx=nn, ny=nn, K="homogeneous", order=1, order_s=0)
        with pkg.Solver(d, Kx, Ky, Kz, phi) as s:
            _, st = s.pressure_solve(np.zeros(nn * nn))
            assert st.converged == 1 and st.iterations == want


def test_rt1_unsupported_orders_fail_loudly(pkg, native_lib):
    from incompressible_flow_b200 import workloads
    d, Kx, Ky, Kz, phi = workloads.build(dim=2, nx=4, ny=4, K="homogeneous", order=1, order_p=0)
    with pytest.raises(pkg.binding.ImpesError):
        pkg.Solver(d, Kx, Ky, Kz, phi)


@pytest.mark.parametrize("cfg", CASES[:4])
def test_rt1_multigrid_mode_bit_exact(pkg, orc, native_lib, cfg):
    """Throughput mode (IMPES_SCHUR_MG) for RT1/Q1: Q1 level (operator S, damped Jacobi) aggregated cell-wise
    onto the two-point hierarchy.  V-cycle and full solve bit-identical to the oracle's restatement; the
    converged solution equals the reference algorithm's (SGS-preconditioned) one to solver tolerance."""
    from incompressible_flow_b200 import abi
    d, o, s, n = make(pkg, orc, cfg, schur_prec=abi.SCHUR_MG)
    o.set_arith(1)
    rng = np.random.default_rng(21)
    S = rng.uniform(0.0, 0.9, n)
    x = rng.standard_normal(o.nv + o.np_)
    s.update_coefficients(S)
    o.update_coefficients(S)
    assert np.array_equal(s.prec_apply(x), o.prec_apply(x))
    xo, sto, _ = o.pressure_solve(S)
    xg, stg = s.pressure_solve(S)
    assert stg.converged == 1 and stg.iterations == sto.iterations
    assert np.array_equal(xg, xo)
    s.close()
    d2, o2, s2, _ = make(pkg, orc, cfg)
    s2.close()
    xr, str_, _ = o2.pressure_solve(S)
    assert rel(xg[:o.nv], xr[:o.nv]) < 1e-4
    if n > 1000:
        assert stg.iterations < str_.iterations


@pytest.mark.parametrize("sweeps", [4, 5])
def test_rt1_multigrid_other_sweep_counts_bit_exact(pkg, orc, native_lib, sweeps):
    """-mg-smooth: other (rounded-up-to-even) sweep counts run the damping sequence 1.4, 0.55, 1.4, 0.55, ... forwards
    before and backwards after the coarse correction; same bits as the oracle's restatement, still a symmetric
    preconditioner (MINRES converges)."""
    from incompressible_flow_b200 import abi
    d, o, s, n = make(pkg, orc, dict(dim=3, nx=6, ny=5, nz=4), schur_prec=abi.SCHUR_MG, mg_smooth=sweeps)
    o.set_arith(1)
    rng = np.random.default_rng(22)
    S = rng.uniform(0.0, 0.9, n)
    x = rng.standard_normal(o.nv + o.np_)
    y = rng.standard_normal(o.nv + o.np_)
    s.update_coefficients(S)
    o.update_coefficients(S)
    px, py = s.prec_apply(x), s.prec_apply(y)
    assert np.array_equal(px, o.prec_apply(x))
    assert abs(y @ px - x @ py) < 1e-10 * abs(y @ px)
    xo, sto, _ = o.pressure_solve(S)
    xg, stg = s.pressure_solve(S)
    assert stg.converged == 1 and stg.iterations == sto.iterations and np.array_equal(xg, xo)
    s.close()
