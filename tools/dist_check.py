"""Distributed Jacobi-PCG check, run under torchrun on >= 2 GPUs of one node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/dist_check.py

Every rank builds its node-block partition of the same global cantilever, assembles its rows
(no exchange), and the ranks solve together with the halo exchange and reductions inside the
persistent kernel (NVLink peer memory).  Rank 0 also solves the whole problem on one GPU and the
distributed solution must agree to 1e-9 relative; iteration counts must match within 2."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fes_b200 as F  # noqa: E402
from fes_b200.dist import (Communicator, DistributedSystem, NodePartition, all_to_all_objects,  # noqa: E402
                           halo_schedule)
from fes_b200.regions import on_plane  # noqa: E402


def problem(kind, n):
    if kind == 'tri':
        mesh = F.create_rect_mesh([[0, 0], [4, 1]], (4 * n + 1, n + 1))
        load, length = [0, -1], 4.0
    else:
        mesh = F.create_box_mesh([[0, 0, 0], [1, 1, 1]], (n, n, n))
        load, length = [0, -5, 0], 1.0
    d = mesh.spatial_dim
    bc = F.BoundaryConditions()
    bc.add(F.BCType.DIRICHLET, on_plane(0, 0.0), [0] * d)
    bc.add(F.BCType.NEUMANN, on_plane(0, length), load)
    return mesh, bc


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    ok = True
    for kind, n, rtol in [('tri', 24, 1e-12), ('tet', 13, 1e-12), ('tet', int(os.environ.get('FES_DIST_N', 61)), 1e-10)]:
        mesh, bc = problem(kind, n)
        d = mesh.spatial_dim
        # global load and constraints (host logic; every rank computes the same arrays)
        prob = F.linear_elastic(mesh, F.LinearElasticMaterial(200.0, 0.3), bc)
        free, fixed, vals = prob.constraints
        b_global = prob.load
        part = NodePartition(mesh.elements, len(mesh.vertices), rank, world)
        Vl = F.FunctionSpace(part.local_mesh(mesh.vertices), n_components=d)
        Al = Vl.assemble_device(F.LinearElasticForm(F.LinearElasticMaterial(200.0, 0.3)))
        ldofs = (d * part.l2g[:, None] + np.arange(d)).ravel()
        mask_g = np.zeros(d * len(mesh.vertices), dtype=np.uint8)
        mask_g[free] = 1
        mask_l = torch.as_tensor(mask_g[ldofs]).to(dev)
        b_l = torch.as_tensor(b_global[ldofs]).to(dev)
        max_dofs = torch.tensor([len(ldofs)], device=dev)
        dist.all_reduce(max_dofs, op=dist.ReduceOp.MAX)
        comm = Communicator(int(max_dofs.item()))
        sched = halo_schedule(part, all_to_all_objects)
        system = DistributedSystem(part, Al, d, mask_l, comm, sched, F.JacobiPCGBackend(rtol=rtol))
        x_l = torch.zeros(len(ldofs), dtype=torch.float64, device=dev)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        system.solve(b_l, x_l)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        its = system.solver.last_iterations
        # gather the owned pieces on rank 0
        owned = x_l[d * part.owned_begin:d * part.owned_end].cpu().numpy()
        pieces = [None] * world
        dist.all_gather_object(pieces, owned)
        if rank == 0:
            u_dist = np.concatenate(pieces)
            single = F.DiscreteSystem(prob.tangent(None), prob.constraints, F.JacobiPCGBackend(rtol=rtol))
            torch.cuda.synchronize(); t1 = time.perf_counter()
            u_one = single.solve(b_global)
            torch.cuda.synchronize(); dt1 = time.perf_counter() - t1
            its1 = single.solver.last_iterations
            err = np.abs(u_dist - u_one).max() / np.abs(u_one).max()
            good = err < 1e-9 and abs(its - its1) <= 2
            ok &= good
            print(f'{kind} n={n}: dofs={len(u_one)} world={world} iters {its} vs {its1} (1 GPU), '
                  f'max rel diff {err:.2e}, {1e6 * dt / max(its, 1):.1f} us/iter vs {1e6 * dt1 / max(its1, 1):.1f} '
                  f'-> {"OK" if good else "MISMATCH"}', flush=True)
        # warm start from the converged solution at a looser tolerance: nothing left to do
        system.solver.rtol = 1e-8
        system.solve(b_l, x_l, warm_start=True)
        if rank == 0:
            w = system.solver.last_iterations
            print(f'   warm restart: {w} iterations', flush=True)
            ok &= w <= 3
        del system, comm
        dist.barrier()
    if rank == 0:
        print('DIST_CHECK', 'PASS' if ok else 'FAIL', flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
