"""GPU parity at the BASELINE sizes the oracle can reach: C1 exactly as specified, one mesh of
>= 256^2 / >= 48^3 cells per space triple, the slab-local host API, and the multi-GPU path (skipped below 2 GPUs).

Tolerances as in test_gpu_parity.py / test_gpu_q1.py: bit-exact against the oracle's canonical arithmetic;
iteration counts +-2 and the Krylov drift bound 1e-6 against the independent assembled formulation (1e-10 only at a
tight MINRES tolerance, see test_q1_tight_solve_meets_1e10)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


def test_c1_exactly_as_specified(pkg, orc, native_lib):
    """BASELINE config C1: 2-D quarter-five-spot, 64x64, RT0/P0, homogeneous K, the reference's solver defaults
    rtol 1e-6 / atol 1e-10 [REF src/param.cpp:8-13], serial preconditioner: 273 / 314 / 312 MINRES iterations for
    the first three two-phase IMPES steps (SURVEY-CHECK), fields bit-identical to the oracle."""
    from incompressible_flow_b200 import workloads
    c, dtg, dtl = workloads.bench_setup("C1")
    d, Kx, Ky, Kz, phi = workloads.build(**c)
    assert (d.nx, d.ny, d.order_v, d.order_p, d.order_s) == (64, 64, 0, 0, 0) and np.all(Kx == 1.0)
    o = orc.Oracle(d, Kx, Ky, Kz, phi)
    o.set_arith(1)
    o0 = orc.Oracle(d, Kx, Ky, Kz, phi)   # the independent MFEM-like arithmetic
    o0.set_arith(0)
    with pkg.Solver(d, Kx, Ky, Kz, phi) as s:
        S = np.zeros(64 * 64)
        S0 = S.copy()
        its = []
        for ti in range(3):
            xo, sto, _ = o.pressure_solve(S)
            xg, stg = s.pressure_solve(S)
            xm, stm, _ = o0.pressure_solve(S0)
            assert np.array_equal(xg, xo) and stg.iterations == sto.iterations
            assert abs(stg.iterations - stm.iterations) <= 2 and rel(xg[:o.nv], xm[:o.nv]) < 1e-6
            So, _ = o.saturation_advance(xo[:o.nv], S, dtg, dtl)
            Sg, nsub = s.saturation_advance(xg[:o.nv], S, ti + 1, dtg, dtl)
            S0, _ = o0.saturation_advance(xm[:o.nv], S0, dtg, dtl)
            assert nsub == 100 and np.array_equal(Sg, So) and rel(Sg, S0) < 1e-6
            S = Sg
            its.append(stg.iterations)
        assert its == [273, 314, 312]


BIG = [
    ("rt0 2-D 256^2, serial preconditioner", dict(dim=2, nx=256, ny=256), "sgs", (100.0, 1.0)),
    ("rt0 3-D 64^3, multigrid", dict(dim=3, nx=64, ny=64, nz=64), "mg", (50.0, 0.5)),
    ("rt1/dg0 3-D 48^3, multigrid", dict(dim=3, nx=48, ny=48, nz=48, order=1, order_s=0), "mg", (50.0, 0.5)),
    ("rt1/dgq1 2-D 256^2, multigrid", dict(dim=2, nx=256, ny=256, order=1, saturation_source=0.5), "mg", (25.0, 0.25)),
    ("rt1/dgq1 3-D 48^3, multigrid", dict(dim=3, nx=48, ny=48, nz=48, order=1, saturation_source=0.5), "mg", (25.0, 0.25)),
]


@pytest.mark.parametrize("tag,cfg,prec,dts", BIG, ids=[b[0] for b in BIG])
def test_full_step_bit_exact_at_size(pkg, orc, native_lib, tag, cfg, prec, dts):
    """One full IMPES step (pressure solve + 100 sub-steps) on a log-normal field at a size comparable to the
    BASELINE configs: iterations, final norm, x and S bit-identical to the oracle's canonical arithmetic."""
    from incompressible_flow_b200 import abi, workloads
    d, Kx, Ky, Kz, phi = workloads.build(K="lognormal", schur_prec=abi.SCHUR_MG if prec == "mg" else abi.SCHUR_SGS_LEX, **cfg)
    o = orc.Oracle(d, Kx, Ky, Kz, phi)
    o.set_arith(1)
    o.set_threads_det(len(os.sched_getaffinity(0)))   # deterministic threading: same bits as one thread
    S = np.zeros(o.ns)
    xo, sto, rc = o.pressure_solve(S)
    assert rc == 0
    So, nso = o.saturation_advance(xo[:o.nv], S, *dts)
    with pkg.Solver(d, Kx, Ky, Kz, phi) as s:
        xg, stg = s.pressure_solve(S)
        Sg, nsg = s.saturation_advance(None, S, 1, *dts)
    assert nso == nsg == 100
    assert stg.iterations == sto.iterations and stg.final_norm == sto.final_norm
    assert np.array_equal(xg, xo)
    assert np.array_equal(Sg, So)


def test_local_api_equals_global_api_on_one_gpu(pkg, orc, native_lib):
    """world_size == 1: the slab-local entry points see the whole mesh, so they must reproduce the global ones."""
    from incompressible_flow_b200 import abi, workloads
    for cfg in (dict(dim=3, nx=10, ny=8, nz=6), dict(dim=2, nx=12, ny=10, order=1, order_s=0),
                dict(dim=3, nx=6, ny=5, nz=4, order=1)):
        d, Kx, Ky, Kz, phi = workloads.build(K="lognormal", **cfg)
        with pkg.Solver(d, Kx, Ky, Kz, phi) as s:
            ld = s.local_dims()
            assert (ld.n_v, ld.n_p, ld.n_s) == (s.nv, s.np_, s.ns) and ld.p_start == 0 and ld.s_start == 0
            S = np.random.default_rng(1).uniform(0.0, 0.6, s.ns)
            xg, st = s.pressure_solve(S)
            Sg, _ = s.saturation_advance(xg[:s.nv], S, 1, 0.4, 0.05)
            xl = np.zeros(ld.n_v + ld.n_p)
            s.pressure_solve_local(S, out=xl)
            Sl = S.copy()
            s.saturation_advance_local(xl[:ld.n_v], Sl, 1, 0.4, 0.05)
            assert np.array_equal(xl, xg) and np.array_equal(Sl, Sg)
            s.set_field_local(abi.FIELD_DARCY_X, xg)
            assert np.array_equal(s.get_field_local(abi.FIELD_DARCY_X), xg)
            assert np.array_equal(s.get_field(abi.FIELD_DARCY_X), xg)


def test_bad_time_steps_are_rejected(pkg, native_lib):
    from incompressible_flow_b200 import abi, workloads
    d, Kx, Ky, Kz, phi = workloads.build(dim=2, nx=8, ny=8)
    with pkg.Solver(d, Kx, Ky, Kz, phi) as s:
        for gdt, dtl in ((float("nan"), 1.0), (float("inf"), 1.0), (1.0, 0.0), (1.0, float("nan")), (1e30, 1e-9)):
            with pytest.raises(pkg.binding.ImpesError) as e:
                s.saturation_advance(None, np.zeros(64), 1, gdt, dtl)
            assert e.value.code == abi.IMPES_ERR_BAD_ARG


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_parity(world):
    """N slabs (one process per GPU, NCCL) against the single-domain CPU oracle: tools/multi_gpu_check.py
    (12 cases: all three space triples, both preconditioners, global and slab-local host arrays)."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(29610 + world),
                          os.path.join(ROOT, "tools", "multi_gpu_check.py")], capture_output=True, text=True, timeout=1500)
    sys.stdout.write(out.stdout[-6000:])
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "FAIL" not in out.stdout and out.stdout.count("-> OK") >= 9
