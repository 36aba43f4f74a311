"""Distributed Jacobi-PCG check, run under torchrun on >= 2 GPUs of one node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/dist_check.py

Every rank builds its node-block partition of the same global cantilever, assembles its rows
(no exchange), and the ranks solve together with the halo exchange and reductions inside the
persistent kernel (NVLink peer memory).  Rank 0 also assembles and solves the whole problem on one GPU.
Checked, per case (ref call path fem/backends.py:201-213 behind fem/system.py:42-52):

* the ranks' CSR row blocks concatenate to the one-GPU pattern (indptr / indices array_equal) and
  values (bit-identical: same contributions, same order);
* the distributed solution agrees with the one-GPU solution to 1e-9 relative, for the overlapped
  single-reduction loop (default) and the textbook loop (FES_PCG_DIST_CLASSIC=1); iteration counts agree
  within 2 % (the dot products are summed in a different order, and the single-reduction recurrence
  predicts r.r one step ahead, so counts are equal only up to rounding);
* a warm restart from the converged solution at a looser tolerance needs (almost) no iteration;
* an iteration cap that cannot be met raises RuntimeError('CG failed ...') on EVERY rank and leaves the
  communicator usable (the next solve converges).

FES_DIST_N = nodes per side of the large tet case (default 61); FES_DIST_OUT = JSON file for rank 0."""
import json
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
                           halo_schedule, owned_row_block)
from fes_b200.regions import on_plane  # noqa: E402

RTOL = 1e-10   # the reference backend's default (fem/backends.py:95-103)


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


def timed_solve(system, b, x, classic=False):
    if classic:
        os.environ['FES_PCG_DIST_CLASSIC'] = '1'
    else:
        os.environ.pop('FES_PCG_DIST_CLASSIC', None)
    x.zero_()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    system.solve(b, x)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    os.environ.pop('FES_PCG_DIST_CLASSIC', None)
    return dt, system.solver.last_iterations


def run_case(kind, n, rank, world, dev, timers=False):
    """One parity case on the ranks of the default process group; returns (record, ok) on rank 0, (None, True)
    elsewhere.  Collective: every rank must call it."""
    mesh, bc = problem(kind, n)
    d = mesh.spatial_dim
    material = F.LinearElasticMaterial(200.0, 0.3)
    # global load and constraints (host logic; every rank computes the same arrays)
    prob = F.linear_elastic(mesh, material, bc)
    free, fixed, vals = prob.constraints
    b_global = prob.load
    part = NodePartition(mesh.elements, len(mesh.vertices), rank, world)
    Vl = F.FunctionSpace(part.local_mesh(mesh.vertices), n_components=d)
    Al = Vl.assemble_device(F.LinearElasticForm(material))
    ldofs = (d * part.l2g[:, None] + np.arange(d)).ravel()
    mask_g = np.zeros(d * len(mesh.vertices), dtype=np.uint8)
    mask_g[free] = 1
    mask_l = torch.as_tensor(mask_g[ldofs]).to(dev)
    b_l = torch.as_tensor(b_global[ldofs]).to(dev)
    max_dofs = torch.tensor([len(ldofs)], device=dev)
    dist.all_reduce(max_dofs, op=dist.ReduceOp.MAX)
    comm = Communicator(int(max_dofs.item()))
    sched = halo_schedule(part, all_to_all_objects)
    system = DistributedSystem(part, Al, d, mask_l, comm, sched, F.JacobiPCGBackend(rtol=RTOL))
    assert system.n_free_global == len(free)
    x_l = torch.zeros(len(ldofs), dtype=torch.float64, device=dev)
    timed_solve(system, b_l, x_l)                                   # first call builds the SELL copy
    dt, its = timed_solve(system, b_l, x_l)
    owned = x_l[d * part.owned_begin:d * part.owned_end].cpu().numpy()
    dtc, itsc = timed_solve(system, b_l, x_l, classic=True)
    owned_c = x_l[d * part.owned_begin:d * part.owned_end].cpu().numpy()
    pieces, pieces_c, blocks = [None] * world, [None] * world, [None] * world
    dist.all_gather_object(pieces, owned)
    dist.all_gather_object(pieces_c, owned_c)
    dist.all_gather_object(blocks, owned_row_block(part, Al, d))
    rec, ok = None, True
    if rank == 0:
        u_dist, u_classic = np.concatenate(pieces), np.concatenate(pieces_c)
        A1 = prob.tangent(None)
        S1 = A1.to_scipy()
        indptr = np.concatenate([[0]] + [np.diff(bk[0]) for bk in blocks]).cumsum()
        pattern_ok = (np.array_equal(indptr, S1.indptr)
                      and np.array_equal(np.concatenate([bk[1] for bk in blocks]), S1.indices))
        values_identical = bool(np.array_equal(np.concatenate([bk[2] for bk in blocks]), S1.data))
        single = F.DiscreteSystem(A1, prob.constraints, F.JacobiPCGBackend(rtol=RTOL))
        single.solve(b_global)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        u_one = single.solve(b_global)
        torch.cuda.synchronize(); dt1 = time.perf_counter() - t1
        its1 = single.solver.last_iterations
        scale = np.abs(u_one).max()
        err, err_c = np.abs(u_dist - u_one).max() / scale, np.abs(u_classic - u_one).max() / scale
        tol_it = max(2, int(0.02 * its1))
        ok = bool(pattern_ok and values_identical and err < 1e-9 and err_c < 1e-9
                  and abs(its - its1) <= tol_it and abs(itsc - its1) <= tol_it)
        rec = {'case': f'{kind} n={n}', 'dofs': int(len(u_one)), 'world': world, 'rtol': RTOL, 'iters_N': int(its),
               'iters_N_classic': int(itsc), 'iters_1': int(its1), 'max_rel_diff': float(err),
               'max_rel_diff_classic': float(err_c), 'pattern_array_equal': bool(pattern_ok),
               'values_bit_identical': values_identical, 'us_per_iter_N': 1e6 * dt / max(its, 1),
               'us_per_iter_N_classic': 1e6 * dtc / max(itsc, 1), 'us_per_iter_1': 1e6 * dt1 / max(its1, 1),
               'ok': ok}
    # warm start from the converged solution at a looser tolerance: nothing left to do
    x_l.zero_()
    system.solve(b_l, x_l)
    system.solver.rtol = 1e-8
    system.solve(b_l, x_l, warm_start=True)
    w = system.solver.last_iterations
    # an iteration cap that cannot be met: every rank must raise, nobody may hang or trap
    system.solver.rtol, system.solver.maxiter = 1e-14, 5
    raised = False
    try:
        system.solve(b_l, torch.zeros_like(x_l))
    except RuntimeError as exc:
        raised = 'CG failed' in str(exc)
    flags = torch.tensor([int(raised)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    system.solver.rtol, system.solver.maxiter = RTOL, None
    x2 = torch.zeros_like(x_l)
    system.solve(b_l, x2)                       # the communicator's epochs are still in step
    again = system.solver.last_iterations
    if rank == 0:
        good = w <= 3 and bool(flags.item()) and abs(again - its) <= 1
        ok = ok and good
        rec.update({'warm_restart_iterations': int(w), 'cg_failed_on_every_rank': bool(flags.item()),
                    'iterations_after_failure': int(again), 'ok': ok})
    del system, comm
    dist.barrier()
    return rec, ok


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    ok, report = True, []
    big = int(os.environ.get('FES_DIST_N', 61))
    for kind, n in [('tri', 24), ('tet', 13), ('tet', big)]:
        rec, good = run_case(kind, n, rank, world, dev)
        ok &= good
        if rank == 0:
            report.append(rec)
            print(json.dumps(rec), flush=True)
    if rank == 0:
        print('DIST_CHECK', 'PASS' if ok else 'FAIL', flush=True)
        out = os.environ.get('FES_DIST_OUT')
        if out:
            with open(out, 'w') as f:
                json.dump({'world': world, 'pass': bool(ok), 'cases': report}, f, indent=1)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
