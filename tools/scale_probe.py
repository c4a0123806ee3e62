"""Throughput-mode scaling probe (run under torch.distributed.run, one rank per GPU; also works with 1 process):
IMPES steps of an n^3 mesh cut into slabs, multigrid Schur preconditioner.  Prints per-step device-side times
(max over ranks) and cell-steps/s.  usage: scale_probe.py <order> <n> [steps] [order_s] [dt_local] [ode]
(100 sub-steps per step; ode > 1 selects the MPI path's RK integrators and its sub-step rule)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from __graft_entry__ import load_package
pkg = load_package()
from incompressible_flow_b200 import workloads, abi

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
order, n3 = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
order_s = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dtl = float(sys.argv[5]) if len(sys.argv) > 5 else 0.5
ode = int(sys.argv[6]) if len(sys.argv) > 6 else 1
kw = dict(substep_rule=abi.SUBSTEPS_PARALLEL, ode_solver_type=ode) if ode > 1 else {}
d, Kx, Ky, Kz, phi = workloads.build(dim=3, nx=n3, ny=n3, nz=n3, K="lognormal", order=order, order_s=order_s,
                                     schur_prec=abi.SCHUR_MG, **kw)
d.rank, d.world_size, d.device = rank, world, local
n = n3 ** 3
s = pkg.Solver(d, Kx, Ky, Kz, phi)
if world > 1:
    uid = [s.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    s.comm_init(uid[0])
s.set_field(abi.FIELD_SATURATION, np.zeros(s.ns))


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


for ti in range(steps + 1):
    sync()
    t0 = time.time()
    _, st = s.pressure_solve(None, want_x=False)
    sync()
    t1 = time.time()
    s.saturation_advance(None, None, ti + 1, 100 * dtl, dtl)
    bad, mx = s.guards()
    sync()
    t2 = time.time()
    if rank == 0:
        print(f"order={order} os={order_s} dtl={dtl} ode={ode} n={n3} world={world} step {ti}{' (warm-up)' if ti == 0 else ''}: iters={st.iterations} solve={t1-t0:.4f}s "
              f"({1e3*(t1-t0)/max(st.iterations,1):.3f} ms/it) sat={t2-t1:.4f}s step={t2-t0:.4f}s -> {1/(t2-t0):.2f} steps/s, "
              f"{n/(t2-t0):.3e} cell-steps/s Smax={mx:.4f} bad={bad}", flush=True)
s.close()
if world > 1:
    dist.destroy_process_group()
