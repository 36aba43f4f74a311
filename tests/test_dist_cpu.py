"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): node-block partition with ghost
elements, order-preserving local numbering, halo schedule, and owner-computes assembly -- each
rank's owned row block, assembled by the oracle on the rank's local mesh, concatenates to the
single-process CSR bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, kind, out_dir):
    for p in (ROOT, os.path.join(ROOT, 'oracle')):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import fes_oracle as fo
        from fes_b200.dist import NodePartition, all_to_all_objects, halo_schedule, owned_row_block
        if kind == 'tri':
            v, e = fo.rect_mesh([[0, 0], [3, 1]], (13, 7))
            okind, c = fo.P1_TRI, 2
        else:
            v, e = fo.box_mesh([[0, 0, 0], [1, 1, 1]], (5, 4, 4))
            okind, c = fo.P1_TET, 3
        rng = np.random.default_rng(3)
        e = e[rng.permutation(len(e))]                      # element order is arbitrary but global
        E = np.linspace(10.0, 90.0, len(e))
        part = NodePartition(e, len(v), rank, world)
        # structure of the partition
        assert np.all(np.diff(part.local_elements) > 0) and np.all(np.diff(part.l2g) > 0)
        assert np.array_equal(part.l2g[part.owned_begin:part.owned_end], np.arange(part.lo, part.hi))
        assert np.array_equal(part.l2g[part.elements_local], e[part.local_elements])
        # owner-computes assembly with the oracle on the local mesh
        vl = v[part.l2g]
        Ke = fo.ke_elastic(fo.geometry(okind, vl[part.elements_local]), E[part.local_elements], 0.3)
        A_local = fo.assemble(part.elements_local, c, part.n_local, Ke)
        ip, ix, data = owned_row_block(part, A_local, c)
        # halo schedule + an emulated exchange of a vector
        ptr, src, dst = halo_schedule(part, all_to_all_objects)
        f = lambda g: np.sin(0.37 * g) + 2.0
        x = np.full(part.n_local, np.nan)
        x[part.owned_begin:part.owned_end] = f(part.l2g[part.owned_begin:part.owned_end])
        sends = [(dst[ptr[r]:ptr[r + 1]], x[src[ptr[r]:ptr[r + 1]]]) if ptr[r + 1] > ptr[r] else None
                 for r in range(world)]
        for item in all_to_all_objects(sends):
            if item is not None:
                x[item[0]] = item[1]
        assert np.array_equal(x, f(part.l2g)), 'ghost values after the halo exchange'
        np.savez(os.path.join(out_dir, f'rank{rank}.npz'), ip=ip, ix=ix, data=data, lo=part.lo, hi=part.hi)
        dist.barrier()
        if rank == 0:
            A = fo.assemble(e, c, len(v), fo.ke_elastic(fo.geometry(okind, v[e]), E, 0.3))
            blocks = [np.load(os.path.join(out_dir, f'rank{r}.npz')) for r in range(world)]
            indptr = np.concatenate([[0]] + [b['ip'][1:] + off for b, off in zip(
                blocks, np.cumsum([0] + [b['ip'][-1] for b in blocks[:-1]]))])
            np.testing.assert_array_equal(indptr, A.indptr)
            np.testing.assert_array_equal(np.concatenate([b['ix'] for b in blocks]), A.indices)
            np.testing.assert_array_equal(np.concatenate([b['data'] for b in blocks]), A.data)
            open(os.path.join(out_dir, 'ok'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,kind', [(2, 'tri'), (3, 'tri'), (2, 'tet')])
def test_partitioned_assembly_and_halo_schedule(tmp_path, world, kind):
    port = 29600 + world * 7 + (0 if kind == 'tri' else 3) + os.getpid() % 50
    mp.spawn(_worker, args=(world, port, kind, str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / 'ok').exists()


def test_partition_edge_cases():
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import fes_oracle as fo
    from fes_b200.dist import NodePartition, even_bounds
    v, e = fo.rect_mesh([[0, 0], [1, 1]], (4, 4))
    one = NodePartition(e, len(v), 0, 1)
    assert one.n_local == len(v) and one.n_owned == len(v) and len(one.ghost_local) == 0
    assert np.array_equal(one.elements_local, e)
    np.testing.assert_array_equal(even_bounds(10, 4), [0, 2, 5, 7, 10])
    # more ranks than node rows: some ranks own very few nodes but still see their ghost elements
    parts = [NodePartition(e, len(v), r, 8) for r in range(8)]
    assert sum(p.n_owned for p in parts) == len(v)
    assert all(len(p.local_elements) > 0 for p in parts)
