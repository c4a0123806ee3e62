"""The C++ host layer (incompressible-flow_b200/host): the reference's Param / option surface /
PressureSolver / SaturationSolver / run_serial interfaces over the C ABI."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "incompressible-flow_b200", "host")


@pytest.fixture(scope="module")
def host_bins(native_lib):
    if not os.path.exists(os.path.join(HOST, "impes_run")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "incompressible-flow_b200")], stdout=subprocess.DEVNULL)
    return os.path.join(HOST, "impes_run"), os.path.join(HOST, "host_selftest")


def test_host_selftest(host_bins, tmp_path):
    out = subprocess.run([host_bins[1], str(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    for name in ("defaults", "options", "binary_io", "vts", "init_arrays", "dofmap"):
        assert f"ok {name}" in out.stdout


def test_driver_usage_and_exit_codes(host_bins):
    """argc == 1 -> hint + exit 1; bad option / -h -> usage + exit 1; -info -> info + exit 1
    [REF src/run_serial.cpp:16-38] [REF src/main.cpp:46-65]."""
    run = host_bins[0]
    o = subprocess.run([run], capture_output=True, text=True)
    assert o.returncode == 1 and "To see the available options, run" in o.stdout
    o = subprocess.run([run, "-nx", "4", "-bogus"], capture_output=True, text=True)
    assert o.returncode == 1 and "Unrecognized option: -bogus" in o.stdout and "--darcy-maxiter" in o.stdout
    o = subprocess.run([run, "-h"], capture_output=True, text=True)
    assert o.returncode == 1
    # every flag of the reference's option table is present
    for flag in ("-d", "-nx", "-ny", "-nz", "-sx", "-sy", "-sz", "-ov", "-op", "-os", "-tf", "-dtg", "-dtl", "-vsg", "-vsl",
                 "-ss", "-Kx", "-Ky", "-Kz", "-phi", "-K-file-x", "-K-file-y", "-K-file-z", "-phi-file", "-outdir", "-extra",
                 "-two-phase", "-single-phase", "-ode", "-inj-xcen", "-inj-ycen", "-inj-zcen", "-inj-r", "-inj-h",
                 "-pro-xcen", "-pro-ycen", "-pro-zcen", "-pro-r", "-pro-h", "-inflow", "-outflow", "-sat-source", "-Sor",
                 "-Swc", "-darcy-maxiter", "-darcy-rtol", "-darcy-atol", "-darcy-print-level", "-info", "-no-info"):
        assert re.search(r"\s%s[ ,<]" % re.escape(flag), o.stdout), flag
    o = subprocess.run([run, "-info"], capture_output=True, text=True)
    assert o.returncode == 1 and "C ABI version" in o.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("order", [0, 1])   # 1 = the reference's default spaces RT1 / L2-Q1 (-ov 1 -op 1 -os 0)
@pytest.mark.parametrize("dim", [2, 3])
def test_driver_run_matches_oracle(pkg, orc, host_bins, tmp_path, dim, order):
    """A whole run of the run_serial-compatible driver (K from a float32 file, 3 global steps) reproduces the
    oracle's loop: same MINRES iteration counts in the log, saturation dumps equal to float32 precision."""
    from incompressible_flow_b200 import workloads
    nx, ny, nz = (20, 16, 1) if dim == 2 else (9, 8, 9)   # hx == hz: the well box r = h covers exactly one cell column
    n = nx * ny * nz
    K32 = np.exp(np.random.default_rng(3).standard_normal(n)).astype(np.float32)
    kfile = tmp_path / "K.bin"
    K32.tofile(kfile)
    out = tmp_path / "out"
    dtg, dtl = (20.0, 1.0) if dim == 2 else (2.0, 0.125)
    h = 100.0 / nx
    common = ["-d", str(dim), "-nx", str(nx), "-ny", str(ny), "-nz", str(nz), "-sx", "100", "-sy", "100", "-sz", "100",
              "-ov", str(order), "-op", str(order), "-os", "0", "-two-phase", "-tf", str(3 * dtg), "-dtg", str(dtg), "-dtl", str(dtl), "-vsg", "0",
              "-K-file-x", str(kfile), "-K-file-y", str(kfile), "-K-file-z", str(kfile), "-outdir", str(out),
              "-extra", "t", "-darcy-maxiter", "5000"]
    if dim == 2:
        wells = ["-inj-xcen", "0", "-inj-ycen", "0", "-inj-r", str(h), "-pro-xcen", "100", "-pro-ycen", "100", "-pro-r", str(h)]
    else:
        wells = ["-inj-xcen", "0", "-inj-zcen", "0", "-inj-ycen", "0", "-inj-h", "100", "-inj-r", str(h),
                 "-pro-xcen", "100", "-pro-zcen", "100", "-pro-ycen", "0", "-pro-h", "100", "-pro-r", str(h)]
    o = subprocess.run([host_bins[0]] + common + wells, capture_output=True, text=True, timeout=300)
    assert o.returncode == 0, o.stdout + o.stderr
    its = [int(m) for m in re.findall(r"Darcy solver converged\. Iter = (\d+)", o.stdout)]
    assert "Number of global time steps: 3" in o.stdout and o.stdout.count("is done") == 3
    # oracle loop on the same inputs
    d = workloads.make_desc(dim, nx, ny, nz, sx=100, sy=100, sz=100, maxiter=5000, order=order, order_s=0)
    for w, c, r in ((d.injection, (0, 0, 0), h), (d.production, (100, 100 if dim == 2 else 0, 100), h)):
        w.center[0], w.center[1], w.center[2] = c
        w.radius, w.height = r, 100.0
    K = K32.astype(np.float64)
    oc = orc.Oracle(d, K, K, K if dim == 3 else None, np.ones(n))
    oc.set_arith(1)
    S, its_o = np.zeros(n), []
    for ti in range(1, 4):
        x, st, rc = oc.pressure_solve(S)
        S, _ = oc.saturation_advance(x[:oc.nv], S, dtg, dtl)
        its_o.append(st.iterations)
        dump = np.fromfile(out / f"saturation_t_{dim}D_2phase_{ti:06d}.bin", dtype=np.float32)
        assert np.array_equal(dump, S.astype(np.float32))
    assert its == its_o


@pytest.mark.gpu
def test_driver_aborts_when_not_converged(host_bins, tmp_path):
    """Non-convergence is fatal like in the reference [REF src/pressure_solver.cpp:110-117]."""
    o = subprocess.run([host_bins[0], "-nx", "16", "-ny", "16", "-ov", "0", "-op", "0", "-os", "0", "-darcy-maxiter", "3",
                        "-inj-r", "7", "-pro-xcen", "100", "-pro-ycen", "100", "-pro-r", "7", "-outdir", str(tmp_path / "o")],
                       capture_output=True, text=True, timeout=120)
    assert o.returncode != 0
    assert "Darcy solver DIDN'T converge. Iter = 3" in o.stderr and "The solver didn't converge => terminating" in o.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [2, 3])
def test_driver_run_dgq1(pkg, orc, host_bins, tmp_path, dim):
    """The driver with `-os 1` (RT1 / L2-Q1 / DG-Q1): iteration counts as the oracle's loop (+-2) and the reference's
    dump semantics -- S.ProjectCoefficient(ValuesInCells) keeps the value of each element's LAST local dof
    [REF src/run_serial.cpp:123-127] [REF src/param.hpp:424-435] -- to float32 precision."""
    from incompressible_flow_b200 import workloads
    nx, ny, nz = (14, 12, 1) if dim == 2 else (6, 5, 6)
    n = nx * ny * nz
    K32 = np.exp(np.random.default_rng(3).standard_normal(n)).astype(np.float32)
    kfile = tmp_path / "K.bin"
    K32.tofile(kfile)
    out = tmp_path / "out"
    dtg, dtl = 0.5, 0.05
    h = 100.0 / nx
    common = ["-d", str(dim), "-nx", str(nx), "-ny", str(ny), "-nz", str(nz), "-sx", "100", "-sy", "100", "-sz", "100",
              "-ov", "1", "-op", "1", "-os", "1", "-two-phase", "-tf", str(2 * dtg), "-dtg", str(dtg), "-dtl", str(dtl), "-vsg", "0",
              "-K-file-x", str(kfile), "-K-file-y", str(kfile), "-K-file-z", str(kfile), "-outdir", str(out),
              "-extra", "q", "-darcy-maxiter", "5000"]
    if dim == 2:
        wells = ["-inj-xcen", "0", "-inj-ycen", "0", "-inj-r", str(h), "-pro-xcen", "100", "-pro-ycen", "100", "-pro-r", str(h)]
    else:
        wells = ["-inj-xcen", "0", "-inj-zcen", "0", "-inj-ycen", "0", "-inj-h", "100", "-inj-r", str(h),
                 "-pro-xcen", "100", "-pro-zcen", "100", "-pro-ycen", "0", "-pro-h", "100", "-pro-r", str(h)]
    o = subprocess.run([host_bins[0]] + common + wells, capture_output=True, text=True, timeout=300)
    assert o.returncode == 0, o.stdout + o.stderr
    its = [int(m) for m in re.findall(r"Darcy solver converged\. Iter = (\d+)", o.stdout)]
    assert o.stdout.count("is done") == 2 and f"dim(S)  = {n * 2 ** dim}" in o.stdout
    d = workloads.make_desc(dim, nx, ny, nz, sx=100, sy=100, sz=100, maxiter=5000, order=1, order_s=1)
    for w, c, r in ((d.injection, (0, 0, 0), h), (d.production, (100, 100 if dim == 2 else 0, 100), h)):
        w.center[0], w.center[1], w.center[2] = c
        w.radius, w.height = r, 100.0
    K = K32.astype(np.float64)
    oc = orc.Oracle(d, K, K, K if dim == 3 else None, np.ones(n))
    oc.set_arith(1)
    S, nsl = np.zeros(oc.ns), 2 ** dim
    for ti in range(1, 3):
        x, st, rc = oc.pressure_solve(S)
        S, _ = oc.saturation_advance(x[:oc.nv], S, dtg, dtl)
        assert abs(its[ti - 1] - st.iterations) <= 2
        dump = np.fromfile(out / f"saturation_q_{dim}D_2phase_{ti:06d}.bin", dtype=np.float32)
        want = S.reshape(n, nsl)[:, nsl - 1]
        assert dump.shape == (n,) and np.allclose(dump, want, rtol=2e-6, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("spaces", ["0 0 0", "0 0 1", "1 1 0", "1 1 1"])
def test_driver_with_foreign_dof_numbering(host_bins, tmp_path, spaces):
    """The reference hands the solvers vectors in MFEM's numbering (RT dofs by mesh entity, face dofs oriented by the
    face's first element).  IMPES_COMPAT_SCRAMBLE makes the MFEM shim number its spaces with a random permutation and
    random RT orientations; the adapter's dof maps (host/dofmap.cpp: identified from GetElementDofs / GetNodes /
    CalcVShape, i.e. MFEM's public element interface) must make the run indistinguishable: same iteration counts,
    bit-identical dumps."""
    ov, op, os_ = spaces.split()
    runs = []
    for seed in ("0", "5"):
        out = tmp_path / f"o{seed}"
        args = [host_bins[0], "-d", "3", "-nx", "6", "-ny", "5", "-nz", "6", "-sx", "6", "-sy", "5", "-sz", "6", "-ov", ov, "-op", op,
                "-os", os_, "-two-phase", "-tf", "1.0", "-dtg", "0.5", "-dtl", "0.05", "-vsg", "0", "-outdir", str(out), "-extra", "s",
                "-inj-xcen", "0", "-inj-zcen", "0", "-inj-ycen", "0", "-inj-h", "5", "-inj-r", "1", "-pro-xcen", "6", "-pro-zcen", "6",
                "-pro-ycen", "0", "-pro-h", "5", "-pro-r", "1", "-sat-source", "0.5", "-darcy-maxiter", "5000"]
        o = subprocess.run(args, capture_output=True, text=True, timeout=300, env=dict(os.environ, IMPES_COMPAT_SCRAMBLE=seed))
        assert o.returncode == 0, o.stdout + o.stderr
        its = re.findall(r"Darcy solver converged\. Iter = (\d+)\. Final norm = (\S+)\.", o.stdout)
        dumps = [np.fromfile(out / f"saturation_s_3D_2phase_{ti:06d}.bin", dtype=np.float32) for ti in (1, 2)]
        runs.append((its, dumps))
    assert len(runs[0][0]) == 2 and runs[0][0] == runs[1][0]
    for a, b in zip(runs[0][1], runs[1][1]):
        assert a.size == 180 and np.array_equal(a, b)


def _run_parallel_driver(host_bins, out, world, extra_args, env_extra=None):
    """Launch impes_run_parallel as `world` processes (one per GPU) the way a job script would: rank / world size from
    the environment, NCCL id through a file."""
    exe = os.path.join(os.path.dirname(host_bins[0]), "impes_run_parallel")
    idf = str(out) + ".ncclid"
    procs = []
    for r in range(world):
        env = dict(os.environ, IMPES_RANK=str(r), IMPES_WORLD_SIZE=str(world), IMPES_NCCL_ID_FILE=idf)
        env.update(env_extra or {})
        procs.append(subprocess.Popen([exe] + extra_args + ["-outdir", str(out)], stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True, env=env))
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, so[-2000:] + se[-2000:]
    return outs[0][0]


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2])
@pytest.mark.parametrize("spaces", ["0 0 0", "1 1 1"])
def test_parallel_driver_matches_oracle(pkg, orc, host_bins, tmp_path, world, spaces):
    """The run_parallel-compatible driver [REF src/run_parallel.cpp:118-230]: ParPressureSolver -> V.Distribute ->
    ParSaturationSolver with the MPI path's sub-step rule and -ode, rank-0 log lines, pressure and saturation per-cell
    dumps; on 2 GPUs every process holds only its slab.  Iteration counts and dumps as the oracle's loop."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from incompressible_flow_b200 import abi, workloads
    ov, op, os_ = (int(v) for v in spaces.split())
    nx, ny, nz, dtg, dtl = 8, 6, 8, 0.6, 0.05
    args = ["-d", "3", "-nx", str(nx), "-ny", str(ny), "-nz", str(nz), "-sx", "8", "-sy", "6", "-sz", "8", "-ov", str(ov), "-op", str(op),
            "-os", str(os_), "-two-phase", "-tf", str(2 * dtg), "-dtg", str(dtg), "-dtl", str(dtl), "-vsg", "1", "-ode", "3",
            "-extra", "p", "-schur-prec", "1", "-sat-source", "0.5", "-darcy-maxiter", "5000",
            "-inj-xcen", "0", "-inj-zcen", "0", "-inj-ycen", "0", "-inj-h", "6", "-inj-r", "1",
            "-pro-xcen", "8", "-pro-zcen", "8", "-pro-ycen", "0", "-pro-h", "6", "-pro-r", "1"]
    so = _run_parallel_driver(host_bins, tmp_path / "o", world, args)
    its = [int(m) for m in re.findall(r"Darcy solver converged\. Iter = (\d+)", so)]
    assert "Number of global time steps: 2" in so and so.count("is done") == 2 and "Kx: min 1 max 1" in so
    d = workloads.make_desc(3, nx, ny, nz, sx=8, sy=6, sz=8, maxiter=5000, order=ov, order_s=os_, schur_prec=abi.SCHUR_MG,
                            substep_rule=abi.SUBSTEPS_PARALLEL, ode_solver_type=3, saturation_source=0.5)
    for w, c in ((d.injection, (0, 0, 0)), (d.production, (8, 0, 8))):
        w.center[0], w.center[1], w.center[2] = c
        w.radius, w.height = 1.0, 6.0
    n = nx * ny * nz
    oc = orc.Oracle(d, np.ones(n), np.ones(n), np.ones(n), np.ones(n))
    oc.set_arith(1)
    S, nsl, npl = np.zeros(oc.ns), oc.ns // n, oc.np_ // n
    for ti in (1, 2):
        x, st, rc = oc.pressure_solve(S)
        S, nsub = oc.saturation_advance(x[:oc.nv], S, dtg, dtl)
        assert nsub == 12 and abs(its[ti - 1] - st.iterations) <= (0 if world == 1 else 2)
        sat = np.fromfile(tmp_path / "o" / f"saturation_p_3D_2phase_{ti:06d}.bin", dtype=np.float32)
        prs = np.fromfile(tmp_path / "o" / f"pressure_p_3D_2phase_{ti:06d}.bin", dtype=np.float32)
        want_s = S.reshape(n, nsl)[:, nsl - 1]
        want_p = x[oc.nv:].reshape(n, npl)[:, npl - 1]
        assert sat.shape == (n,) and np.allclose(sat, want_s, rtol=2e-5, atol=1e-6)
        # the pressure is determined up to a constant (all-Neumann problem)
        assert np.allclose(prs - prs.mean(), want_p - want_p.mean(), rtol=1e-4, atol=1e-4 * np.abs(want_p).max())
