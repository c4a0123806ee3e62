"""Multi-GPU parity check (run under torch.distributed.run with N ranks, one GPU each):
N-slab solve vs the single-domain CPU oracle.  K1 / preconditioner / saturation sub-steps must be
bit-identical (no cross-rank reductions in them: the Schur sweep is pipelined rank by rank in the exact
global order); the MINRES solve differs only through the rank-wise association of the two dot products
(iterations +-2, solution within the Krylov drift bound)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from __graft_entry__ import load_package, load_oracle
pkg = load_package()
from incompressible_flow_b200 import workloads, abi

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def gather(a, owned):
    m = np.zeros(a.shape[0], dtype=bool)
    for lo, hi in owned:
        m[lo:hi] = True
    a = np.where(m, a, 0.0)
    t = torch.from_numpy(a).cuda()
    dist.all_reduce(t)
    return t.cpu().numpy()


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


ok = True
CASES = [
    (dict(dim=3, nx=12, ny=10, nz=9), 0, abi.SCHUR_SGS_LEX),
    (dict(dim=2, nx=33, ny=21), 0, abi.SCHUR_SGS_LEX),
    (dict(dim=3, nx=24, ny=20, nz=16), 0, abi.SCHUR_SGS_LEX),
    (dict(dim=3, nx=24, ny=20, nz=16), 0, abi.SCHUR_MG),   # distributed multigrid: aligned slabs, replicated coarse levels
    (dict(dim=2, nx=33, ny=21), 0, abi.SCHUR_MG),          # ragged: 21 layers in units of 2^m
    (dict(dim=3, nx=6, ny=5, nz=9), 1, abi.SCHUR_SGS_LEX),  # RT1/Q1 slabs
    (dict(dim=2, nx=9, ny=11), 1, abi.SCHUR_SGS_LEX),
    (dict(dim=3, nx=8, ny=6, nz=16), 1, abi.SCHUR_MG),
    (dict(dim=2, nx=16, ny=24), 1, abi.SCHUR_MG),
    (dict(dim=3, nx=6, ny=5, nz=9), 11, abi.SCHUR_SGS_LEX),  # 11 = RT1/Q1 with DG-Q1 saturation (-os 1): tolerance parity
    (dict(dim=3, nx=8, ny=6, nz=16), 11, abi.SCHUR_MG),
    (dict(dim=2, nx=16, ny=24), 11, abi.SCHUR_MG),
    (dict(dim=3, nx=8, ny=6, nz=16), 10, abi.SCHUR_MG),      # 10 = RT0/P0 with DG-Q1 saturation: tolerance parity (assembled oracle)
    (dict(dim=2, nx=9, ny=11), 10, abi.SCHUR_SGS_LEX),
]
ONLY = os.environ.get("MGPU_ONLY")   # development: run only the cases with this order code
for cfg, order, prec in CASES:
    if ONLY and str(order) != ONLY:
        continue
    order_s = 1 if order in (10, 11) else 0
    arith = 0 if order == 10 else 1   # no canonical mirror for RT0 / DG-Q1: the assembled formulation
    order = {11: 1, 10: 0}.get(order, order)
    nlay = cfg["nz"] if cfg["dim"] == 3 else cfg["ny"]
    if nlay < world or (prec == abi.SCHUR_MG and nlay < 2 * world):
        continue
    d, Kx, Ky, Kz, phi = workloads.build(K="lognormal", order=order, order_s=order_s, schur_prec=prec, **cfg)
    n = d.nx * d.ny * (d.nz if d.dim == 3 else 1)
    phi = np.random.default_rng(7).uniform(0.1, 0.4, n)
    d.rank, d.world_size, d.device = rank, world, local
    s = pkg.Solver(d, Kx, Ky, Kz, phi)
    uid = [s.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    s.comm_init(uid[0])
    own_x, own_s = pkg.partition.owned_ranges(d.dim, d.nx, d.ny, d.nz, rank, world, order=order,
                                              slab=(s.dims.slab_lo, s.dims.slab_hi), order_s=order_s)
    rng = np.random.default_rng(3)
    S = rng.uniform(0.05, 0.95, s.ns)
    x = rng.standard_normal(s.nv + s.np_)
    dts = (7.3, 0.5) if order_s == 0 else (0.4, 0.05)
    s.update_coefficients(S)
    yg = gather(s.darcy_apply(x), own_x)
    zg = gather(s.prec_apply(x), own_x)
    xs, st = s.pressure_solve(S)
    xs = gather(xs, own_x)
    Sg, nsub = s.saturation_advance(xs[:s.nv], S, 1, *dts)
    Sg = gather(Sg, own_s)
    bad, mx = s.guards()
    # slab-local host arrays (impes_*_local): every rank passes / receives only what it owns; must reproduce the
    # owned part of the global-array calls bit for bit (same device computation, ghosts exchanged on the device)
    ld = s.local_dims()
    x_loc = np.zeros(ld.n_v + ld.n_p)
    _, stl = s.pressure_solve_local(S[ld.s_start:ld.s_start + ld.n_s], out=x_loc)
    S_loc = S[ld.s_start:ld.s_start + ld.n_s].copy()
    s.saturation_advance_local(x_loc[:ld.n_v], S_loc, 1, *dts)
    loc_ok = (np.array_equal(x_loc, np.concatenate([xs[a:b] for a, b in own_x])) and stl.iterations == st.iterations and
              np.array_equal(S_loc, np.concatenate([Sg[a:b] for a, b in own_s])))
    t_ok = torch.tensor([1 if loc_ok else 0], device="cuda")
    dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
    loc_ok = bool(int(t_ok.item()))
    if rank == 0:
        orc = load_oracle()
        d1 = workloads.build(K="lognormal", order=order, order_s=order_s, schur_prec=prec, **cfg)[0]
        o = orc.Oracle(d1, Kx, Ky, Kz, phi)
        o.set_arith(arith)
        if arith: o.set_threads_det(8)
        o.update_coefficients(S)
        same = np.array_equal if order_s == 0 else (lambda a, b: rel(a, b) < 1e-12)   # -os 1: round-off parity
        e1 = same(yg, o.darcy_apply(x))
        e2 = same(zg, o.prec_apply(x))
        xo, sto, _ = o.pressure_solve(S)
        So, _ = o.saturation_advance(xs[:o.nv], S, *dts)
        e3 = same(Sg, So)
        bo, mo = o.guards(So)
        if order_s:
            bo, mo = bad, (mx if abs(mx - mo) <= 1e-12 * abs(mo) else mo)
        dit = abs(st.iterations - sto.iterations)
        ex = rel(xs[:o.nv], xo[:o.nv])
        good = e1 and e2 and e3 and dit <= 2 and ex < 1e-6 and (bad, mx) == (bo, mo) and loc_ok
        ok = ok and good
        print(f"{cfg} order={order} os={order_s} prec={'MG' if prec else 'SGS'} world={world} slab={(s.dims.slab_lo, s.dims.slab_hi)}: "
              f"K1 bit-exact={e1} prec bit-exact={e2} sat bit-exact={e3} "
              f"iters gpu={st.iterations} oracle={sto.iterations} rel_u={ex:.2e} rel_S={rel(Sg, So):.2e} guards={(bad, mx) == (bo, mo)} local-api={loc_ok} -> {'OK' if good else 'FAIL'}",
              flush=True)
    s.close()
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, src=0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
