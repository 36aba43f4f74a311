"""Scratch: where does the distributed warm start get its residual from?  (2 GPUs)"""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tools'))
import fes_b200 as F
from fes_b200.dist import Communicator, DistributedSystem, NodePartition, all_to_all_objects, halo_schedule
from dist_check import problem

def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local); dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    for kind, n in [('tri', 24), ('tet', 13)]:
        mesh, bc = problem(kind, n); d = mesh.spatial_dim
        prob = F.linear_elastic(mesh, F.LinearElasticMaterial(200.0, 0.3), bc)
        free, fixed, vals = prob.constraints; b_global = prob.load
        part = NodePartition(mesh.elements, len(mesh.vertices), rank, world)
        Vl = F.FunctionSpace(part.local_mesh(mesh.vertices), n_components=d)
        Al = Vl.assemble_device(F.LinearElasticForm(F.LinearElasticMaterial(200.0, 0.3)))
        ldofs = (d * part.l2g[:, None] + np.arange(d)).ravel()
        mask_g = np.zeros(d * len(mesh.vertices), dtype=np.uint8); mask_g[free] = 1
        mask_l = torch.as_tensor(mask_g[ldofs]).to(dev); b_l = torch.as_tensor(b_global[ldofs]).to(dev)
        md = torch.tensor([len(ldofs)], device=dev); dist.all_reduce(md, op=dist.ReduceOp.MAX)
        comm = Communicator(int(md.item())); sched = halo_schedule(part, all_to_all_objects)
        system = DistributedSystem(part, Al, d, mask_l, comm, sched, F.JacobiPCGBackend(rtol=1e-12))
        x_l = torch.zeros(len(ldofs), dtype=torch.float64, device=dev)
        system.solve(b_l, x_l)
        its = system.solver.last_iterations
        owned = x_l[d * part.owned_begin:d * part.owned_end].cpu().numpy()
        pieces = [None] * world; dist.all_gather_object(pieces, owned)
        u_dist = np.concatenate(pieces)
        A = prob.tangent(None).to_scipy() if rank == 0 else None
        # the kernel's own view of the starting residual: rtol = 1 stops before the first iteration
        system.solver.rtol = 1.0
        system.solve(b_l, x_l, warm_start=True)
        rel_kernel = system.solver.last_relres
        for rt in (1e-8, 1e-10):
            xs = x_l.clone(); system.solver.rtol = rt
            system.solve(b_l, xs, warm_start=True)
            if rank == 0: print(f'   dist warm rtol {rt:g}: {system.solver.last_iterations} its relres {system.solver.last_relres:.3e}', flush=True)
        if rank == 0:
            r = (b_global - A @ u_dist)[free]
            print(f'{kind} n={n} world={world}: iters {its}; host true relres {np.linalg.norm(r)/np.linalg.norm(b_global[free]):.3e}; '
                  f'kernel warm relres {rel_kernel:.3e}', flush=True)
            single = F.DiscreteSystem(prob.tangent(None), prob.constraints, F.JacobiPCGBackend(rtol=1e-12))
            u1 = single.solve(prob.load_d)
            r1 = (b_global - A @ u1.cpu().numpy())[free]
            single.solver.rtol = 1.0
            single.solve(prob.load_d, x0=u1.clone())
            print(f'   one GPU: iters? true relres {np.linalg.norm(r1)/np.linalg.norm(b_global[free]):.3e}; kernel warm relres {single.solver.last_relres:.3e}', flush=True)
            for rt in (1e-8, 1e-10):
                single.solver.rtol = rt
                single.solve(prob.load_d, x0=u1.clone())
                print(f'   one GPU warm rtol {rt:g}: {single.solver.last_iterations} its', flush=True)
        del system, comm
        dist.barrier()
    dist.destroy_process_group()

if __name__ == '__main__':
    main()
