"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Tolerances (fp64):
  * oracle in canonical arithmetic (same round-off choices as the GPU):
      bit-exact (np.array_equal) -- stronger than the 1e-10 of BASELINE.json's north_star;
  * oracle in MFEM-like arithmetic (quadrature-assembled CSR, sequential sums -- the independent
      formulation): operator level 1e-13 relative; solve level = identical iteration counts and the
      Krylov drift bound demonstrated in test_oracle_invariants.py::test_krylov_drift.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    dict(dim=2, nx=33, ny=21),
    dict(dim=2, nx=64, ny=64),
    dict(dim=3, nx=12, ny=9, nz=7),
    dict(dim=3, nx=5, ny=300, nz=3),   # ragged: more than one 256-block per plane row
    dict(dim=2, nx=1, ny=5),           # degenerate: single column
]


def rel(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


def make(pkg, orc, cfg, K="lognormal", **kw):
    from incompressible_flow_b200 import workloads
    d, Kx, Ky, Kz, phi = workloads.build(K=K, sigma=1.0, **cfg, **kw)
    rng = np.random.default_rng(7)
    n = d.nx * d.ny * (d.nz if d.dim == 3 else 1)
    phi = rng.uniform(0.1, 0.4, n)
    o = orc.Oracle(d, Kx, Ky, Kz, phi)
    s = pkg.Solver(d, Kx, Ky, Kz, phi)
    return d, o, s, n


@pytest.mark.parametrize("cfg", CASES)
def test_dims_and_rhs(pkg, orc, native_lib, cfg):
    d, o, s, n = make(pkg, orc, cfg)
    assert (s.nv, s.np_, s.ns) == (o.nv, o.np_, o.ns)
    assert np.array_equal(s.rhs(), o.rhs())
    s.close()


@pytest.mark.parametrize("cfg", CASES)
@pytest.mark.parametrize("two_phase", [True, False])
def test_darcy_apply_and_prec(pkg, orc, native_lib, cfg, two_phase):
    d, o, s, n = make(pkg, orc, cfg, two_phase=two_phase)
    rng = np.random.default_rng(3)
    S = rng.uniform(0.05, 0.95, n)
    x = rng.standard_normal(o.nv + o.np_)
    s.update_coefficients(S)
    yg = s.darcy_apply(x)
    zg = s.prec_apply(x)
    # independent formulation: quadrature-assembled CSR, sequential Gauss-Seidel
    o.set_arith(0)
    o.update_coefficients(S)
    assert rel(yg, o.darcy_apply(x)) < 1e-13
    assert rel(zg, o.prec_apply(x)) < 1e-12
    # canonical arithmetic: bit for bit
    o.set_arith(1)
    o.update_coefficients(S)
    assert np.array_equal(yg, o.darcy_apply(x))
    assert np.array_equal(zg, o.prec_apply(x))
    s.close()


@pytest.mark.parametrize("cfg", CASES[:4])
def test_pressure_solve_bit_exact(pkg, orc, native_lib, cfg):
    d, o, s, n = make(pkg, orc, cfg)
    S = np.random.default_rng(5).uniform(0.0, 0.9, n)
    o.set_arith(1)
    xo, sto, rco = o.pressure_solve(S)
    xg, stg = s.pressure_solve(S)
    assert stg.converged == 1 and sto.converged == 1
    assert stg.iterations == sto.iterations
    assert stg.final_norm == sto.final_norm
    assert np.array_equal(xg, xo)
    # against the MFEM-like arithmetic: same iteration count (+-2), solution within the Krylov drift bound
    o.set_arith(0)
    xm, stm, _ = o.pressure_solve(S)
    assert abs(stg.iterations - stm.iterations) <= 2
    assert rel(xg[:o.nv], xm[:o.nv]) < 1e-6
    s.close()


@pytest.mark.parametrize("cfg", CASES[:4])
@pytest.mark.parametrize("two_phase", [True, False])
def test_saturation_substeps_bit_exact(pkg, orc, native_lib, cfg, two_phase):
    d, o, s, n = make(pkg, orc, cfg, two_phase=two_phase)
    o.set_arith(1)
    S0 = np.zeros(n)
    xo, _, _ = o.pressure_solve(S0)
    So, nso = o.saturation_advance(xo[:o.nv], S0, 7.3, 0.5)
    Sg, nsg = s.saturation_advance(xo[:o.nv], S0, 1, 7.3, 0.5)
    assert nso == nsg == 14          # int(7.3/0.5): C truncation [REF src/saturation_solver.cpp:74-75]
    assert np.array_equal(Sg, So)
    # the reference's Jacobi-PCG mass solve agrees with the exact diagonal inverse to a few ulp
    o.set_mass_cg(1)
    Scg, _ = o.saturation_advance(xo[:o.nv], S0, 7.3, 0.5)
    assert rel(Scg, So) < 1e-14
    s.close()


def test_nontrivial_relperm_constants_bit_exact(pkg, orc, native_lib):
    """Sor, Swc, mu_w, mu_o != trivial values: the GPU evaluates the constant-denominator divisions with
    div_exact (reciprocal + FMA corrections), which must equal IEEE division bit for bit."""
    cfg = dict(dim=3, nx=17, ny=9, nz=11)
    d, o, s, n = make(pkg, orc, cfg, Sor=0.13, Swc=0.21, mu_w=0.7, mu_o=2.3)
    o.set_arith(1)
    S = np.random.default_rng(9).uniform(0.25, 0.8, n)
    xo, sto, _ = o.pressure_solve(S)
    xg, stg = s.pressure_solve(S)
    assert stg.iterations == sto.iterations and np.array_equal(xg, xo)
    So, _ = o.saturation_advance(xo[:o.nv], S, 6.0, 0.25)
    Sg, _ = s.saturation_advance(xo[:o.nv], S, 1, 6.0, 0.25)
    assert np.array_equal(Sg, So)
    s.close()


@pytest.mark.parametrize("cfg", [dict(dim=2, nx=300, ny=7), dict(dim=3, nx=61, ny=13, nz=5), dict(dim=3, nx=30, ny=6, nz=40)])
def test_saturation_tile_edges(pkg, orc, native_lib, cfg):
    """Mesh sizes around the K6 tile geometry (30x6 interior of 32x8 tiles, 254 of 256 in 2-D, z-chunks)."""
    d, o, s, n = make(pkg, orc, cfg)
    o.set_arith(1)
    rng = np.random.default_rng(4)
    u = rng.standard_normal(o.nv)
    S0 = rng.uniform(0, 1, n)
    So, _ = o.saturation_advance(u, S0, 0.02, 0.005)
    Sg, _ = s.saturation_advance(u, S0, 1, 0.02, 0.005)
    assert np.array_equal(Sg, So)
    s.close()


@pytest.mark.parametrize("ode", [1, 2, 3, 4])
def test_ode_variants_parallel_rule(pkg, orc, native_lib, ode):
    from incompressible_flow_b200 import abi
    cfg = dict(dim=3, nx=9, ny=8, nz=6)
    d, o, s, n = make(pkg, orc, cfg, substep_rule=abi.SUBSTEPS_PARALLEL, ode_solver_type=ode)
    o.set_arith(1)
    xo, _, _ = o.pressure_solve(np.zeros(n))
    So, nso = o.saturation_advance(xo[:o.nv], np.zeros(n), 5.2, 0.5)
    Sg, nsg = s.saturation_advance(xo[:o.nv], np.zeros(n), 1, 5.2, 0.5)
    assert nso == nsg == 10          # while t < global_dt - dt/2 [REF src/par_saturation_solver.cpp:149-155]
    assert np.array_equal(Sg, So)
    s.close()


MG_CASES = [dict(dim=2, nx=33, ny=21), dict(dim=2, nx=128, ny=96), dict(dim=3, nx=12, ny=9, nz=7),
            dict(dim=3, nx=30, ny=44, nz=17), dict(dim=2, nx=3, ny=2)]


@pytest.mark.parametrize("cfg", MG_CASES)
def test_multigrid_mode_bit_exact(pkg, orc, native_lib, cfg):
    """Throughput mode (IMPES_SCHUR_MG): the V-cycle preconditioner and the whole MINRES solve equal the
    oracle's restatement of the same algorithm bit for bit; the converged solution agrees with the reference
    algorithm's (SGS-preconditioned) one to solver tolerance."""
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
    # same discrete problem as the reference algorithm: solutions agree to solver tolerance
    d2, o2, s2, _ = make(pkg, orc, cfg)
    s2.close()
    xr, str_, _ = o2.pressure_solve(S)
    assert rel(xg[:o.nv], xr[:o.nv]) < 1e-4
    if n > 1000:
        assert stg.iterations < str_.iterations


def test_unknown_ode_is_an_error(pkg, orc, native_lib):
    from incompressible_flow_b200 import abi
    d, o, s, n = make(pkg, orc, dict(dim=2, nx=8, ny=8), substep_rule=abi.SUBSTEPS_PARALLEL, ode_solver_type=5)
    with pytest.raises(pkg.binding.ImpesError) as e:
        s.saturation_advance(np.zeros(o.nv), np.zeros(n), 1, 1.0, 0.5)
    assert "Uknown ODE solver" in str(e.value)   # message of the reference [REF src/par_saturation_solver.cpp:121-123]
    s.close()


@pytest.mark.parametrize("cfg", [dict(dim=2, nx=32, ny=32), dict(dim=3, nx=10, ny=10, nz=8)])
def test_impes_loop_resident_state(pkg, orc, native_lib, cfg):
    """Three global IMPES steps with device-resident state == the oracle's loop, bit for bit
    (the loop of run_serial [REF src/run_serial.cpp:109-143])."""
    d, o, s, n = make(pkg, orc, cfg)
    o.set_arith(1)
    So = np.zeros(n)
    s.set_field(0, So)
    its_o, its_g = [], []
    for ti in range(1, 4):
        xo, sto, _ = o.pressure_solve(So)
        So, _ = o.saturation_advance(xo[:o.nv], So, 20.0, 1.0)
        _, stg = s.pressure_solve(None, want_x=False)
        s.saturation_advance(None, None, ti, 20.0, 1.0)
        its_o.append(sto.iterations)
        its_g.append(stg.iterations)
    from incompressible_flow_b200 import abi
    assert its_o == its_g
    assert np.array_equal(s.get_field(abi.FIELD_SATURATION), So)
    assert np.array_equal(s.get_field(abi.FIELD_DARCY_X), xo)
    bad, mx = s.guards()
    bo, mo = o.guards(So)
    assert (bad, mx) == (bo, mo)
    s.close()


def test_not_converged_status(pkg, orc, native_lib):
    from incompressible_flow_b200 import abi
    d, o, s, n = make(pkg, orc, dict(dim=2, nx=32, ny=32), maxiter=10)
    o.set_arith(1)
    xo, sto, rco = o.pressure_solve(np.zeros(n))
    xg, stg = s.pressure_solve(np.zeros(n), allow_not_converged=True)
    assert rco == abi.IMPES_ERR_NOT_CONVERGED
    assert stg.converged == 0 and stg.iterations == sto.iterations == 10
    assert np.array_equal(xg, xo)
    with pytest.raises(pkg.binding.ImpesError):
        s.pressure_solve(np.zeros(n))
    s.close()


def test_guards_nonfinite(pkg, orc, native_lib):
    d, o, s, n = make(pkg, orc, dict(dim=2, nx=16, ny=16))
    S = np.linspace(-0.5, 1.25, n)
    S[3] = np.nan
    S[17] = np.inf
    s.set_field(0, S)
    bad, mx = s.guards()
    assert bad == 2 and mx == 1.25
    s.close()


def test_graph_and_direct_paths_agree(pkg, orc, native_lib, monkeypatch):
    cfg = dict(dim=2, nx=48, ny=40)
    d, o, s, n = make(pkg, orc, cfg)
    S = np.random.default_rng(11).uniform(0, 0.8, n)
    xg, stg = s.pressure_solve(S)
    s.close()
    monkeypatch.setenv("IMPES_NO_GRAPH", "1")
    d, o, s2, n = make(pkg, orc, cfg)
    xd, std = s2.pressure_solve(S)
    s2.close()
    assert stg.iterations == std.iterations and np.array_equal(xg, xd)


@pytest.mark.gpu
def test_describe_names_the_selected_variants(pkg):
    """impes_describe: which sweep kernel / multigrid set-up a handle chose for its mesh."""
    from incompressible_flow_b200 import abi, workloads
    d, Kx, Ky, Kz, phi = workloads.build(dim=2, nx=24, ny=20, K="lognormal")
    with pkg.Solver(d, Kx, Ky, Kz, phi) as s:
        t = s.describe()
        assert "RT0/L2_0/DG_0" in t and "Gauss-Seidel" in t and "k_sgs" in t
    d, Kx, Ky, Kz, phi = workloads.build(dim=3, nx=8, ny=8, nz=8, K="lognormal", order=1, order_s=1, schur_prec=abi.SCHUR_MG)
    with pkg.Solver(d, Kx, Ky, Kz, phi) as s:
        t = s.describe()
        assert "RT1/L2_1/DG_1" in t and "multigrid V(2,2)" in t and "L2-Q1 level" in t and "k1q_sat" in t
