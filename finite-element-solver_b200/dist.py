"""Sharding the path across the GPUs of one node.

Partition: contiguous NODE blocks (row blocks of the operator).  A rank keeps every element that
touches a node it owns -- its own elements plus one layer of ghost elements -- in ascending
global element id, and renumbers their nodes with an ORDER-PRESERVING map (local id = rank of the
global id among the local nodes).  Consequences:

* assembly is owner-computes with no communication: the rows of owned nodes receive exactly the
  contributions, in exactly the order, they receive on one GPU, so the row blocks of all ranks
  concatenate to the single-GPU CSR (indptr/indices bit-exact after mapping columns back);
* owned rows are one contiguous local range, so the distributed PCG restricts its vector updates
  and its SELL copy to that range;
* the only exchange step is the halo of the CG direction vector (ghost nodes' values, written by
  their owners straight into this rank's vector over NVLink peer memory) plus two scalar
  all-reductions per iteration, all inside the persistent kernel (csrc/pcg.cu).

`torch.distributed` is plumbing only: it carries the halo index lists and the CUDA IPC handles.
Everything in `NodePartition` / `halo_schedule` is numpy and runs on CPU (tests use gloo).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .mesh import Mesh


def even_bounds(n_nodes, world):
    return np.array([(n_nodes * r) // world for r in range(world + 1)], dtype=np.int64)


class NodePartition:
    """What one rank holds of a global mesh."""

    def __init__(self, elements, n_nodes, rank, world, bounds=None):
        elements = np.asarray(elements, dtype=np.int64)
        self.rank, self.world, self.n_nodes_global = rank, world, int(n_nodes)
        self.bounds = even_bounds(n_nodes, world) if bounds is None else np.asarray(bounds, dtype=np.int64)
        self.lo, self.hi = int(self.bounds[rank]), int(self.bounds[rank + 1])
        touch = ((elements >= self.lo) & (elements < self.hi)).any(axis=1)
        self.local_elements = np.flatnonzero(touch)                # ascending global element ids
        le = elements[touch]
        # local nodes = the owned block plus every node of a touching element, ascending; a mark array and a
        # global -> local lookup instead of a sort of all (element, corner) entries (20 M tets: 4.1 -> 1.x s per rank)
        mark = np.zeros(self.n_nodes_global, dtype=bool)
        mark[le.ravel()] = True
        mark[self.lo:self.hi] = True
        self.l2g = np.flatnonzero(mark)
        g2l = np.empty(self.n_nodes_global, dtype=np.int64)
        g2l[self.l2g] = np.arange(len(self.l2g), dtype=np.int64)
        self.elements_local = g2l[le]
        self.owned_begin = int(np.searchsorted(self.l2g, self.lo))
        self.owned_end = int(np.searchsorted(self.l2g, self.hi))
        ghost = np.ones(len(self.l2g), dtype=bool)
        ghost[self.owned_begin:self.owned_end] = False
        self.ghost_local = np.flatnonzero(ghost)
        self.ghost_owner = np.searchsorted(self.bounds, self.l2g[self.ghost_local], side='right') - 1

    @property
    def n_local(self):
        return len(self.l2g)

    @property
    def n_owned(self):
        return self.owned_end - self.owned_begin

    def local_mesh(self, vertices, boundary=None):
        """The rank's mesh in local numbering.  `boundary` (global facets) is restricted to facets
        whose nodes are all local; facets of the cut are not boundary facets and never appear."""
        v = np.asarray(vertices)[self.l2g]
        if boundary is None:
            b = np.zeros((0, self.elements_local.shape[1] - 1), dtype=np.int64)
        else:
            boundary = np.asarray(boundary, dtype=np.int64)
            pos = np.searchsorted(self.l2g, boundary)
            pos = np.minimum(pos, len(self.l2g) - 1)
            keep = (self.l2g[pos] == boundary).all(axis=1)
            b = pos[keep]
        return Mesh(v, self.elements_local, b)

    def needs(self):
        """{owner rank: (global ids, local ids)} of the ghost nodes this rank must receive."""
        out = {}
        for r in np.unique(self.ghost_owner):
            sel = self.ghost_owner == r
            out[int(r)] = (self.l2g[self.ghost_local[sel]], self.ghost_local[sel])
        return out


def halo_schedule(part, exchange):
    """Send lists of `part`: for every peer, which owned local nodes it wants and where they land
    in ITS numbering.  `exchange(list_per_rank) -> list_per_rank` is an all-to-all of python
    objects (torch.distributed.all_to_all_object-like; the tests pass a gloo implementation).
    Returns (send_ptr[world+1], send_src_node[], send_dst_node[])."""
    needs = part.needs()
    outgoing = [None] * part.world
    for r in range(part.world):
        if r in needs:
            gids, lids = needs[r]
            outgoing[r] = (gids, lids)
    incoming = exchange(outgoing)      # incoming[r] = (global ids r needs from me, r's local ids for them)
    ptr, src, dst = [0], [], []
    for r in range(part.world):
        if incoming[r] is not None:
            gids, their_lids = incoming[r]
            mine = np.searchsorted(part.l2g, gids)
            assert np.array_equal(part.l2g[mine], gids), 'peer asked for a node this rank does not hold'
            assert ((mine >= part.owned_begin) & (mine < part.owned_end)).all(), 'peer asked for a node this rank does not own'
            src.append(mine)
            dst.append(np.asarray(their_lids))
        ptr.append(ptr[-1] + (len(src[-1]) if incoming[r] is not None else 0))
    cat = (lambda parts: np.concatenate(parts).astype(np.int32) if parts else np.zeros(0, dtype=np.int32))
    return np.array(ptr, dtype=np.int32), cat(src), cat(dst)


def all_to_all_objects(outgoing, group=None):
    """all-to-all of picklable objects over torch.distributed (any backend)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    gathered = [None] * world
    dist.all_gather_object(gathered, outgoing, group=group)
    me = dist.get_rank(group)
    return [gathered[r][me] for r in range(world)]


def owned_row_block(part, A, n_comp):
    """(indptr, global indices, data) of the rows this rank owns, from its local DeviceCSR / csr."""
    S = A.to_scipy() if hasattr(A, 'to_scipy') else A
    r0, r1 = n_comp * part.owned_begin, n_comp * part.owned_end
    indptr = S.indptr[r0:r1 + 1] - S.indptr[r0]
    sl = slice(S.indptr[r0], S.indptr[r1])
    cols = S.indices[sl]
    gcols = n_comp * part.l2g[cols // n_comp] + cols % n_comp
    return indptr, gcols, S.data[sl]


class Communicator:
    """Peer-memory communicator of the distributed PCG: one CUDA IPC region per rank holding the
    halo/reduction mailboxes and the shared direction vector; peers map each other's regions."""

    def __init__(self, max_local_dofs, group=None):
        import torch.distributed as dist
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        L = _lib.lib()
        h = C.c_void_p()
        _lib.check(L.fes_comm_create(self.rank, self.world, int(max_local_dofs), C.byref(h)))
        self._h = h
        mine = (C.c_ubyte * 64)()
        _lib.check(L.fes_comm_handle(h, mine))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(mine), group=group)
        blob = (C.c_ubyte * (64 * self.world)).from_buffer_copy(b''.join(handles))
        _lib.check(L.fes_comm_connect(h, blob))
        dist.barrier(group)

    @property
    def handle(self):
        return self._h

    def __del__(self):
        try:
            if self._h:
                _lib.lib().fes_comm_destroy(self._h)
                self._h = None
        except Exception:
            pass


class DistributedSystem:
    """DiscreteSystem over a node partition: owned rows assembled locally, Jacobi-PCG with the
    halo exchange and the dot-product reductions inside the persistent kernel."""

    def __init__(self, part, A_local, n_comp, free_mask_local, comm, schedule, backend, group=None):
        """A_local: DeviceCSR on the local numbering (from FunctionSpace.assemble_device on
        part.local_mesh); free_mask_local: uint8 (n_local_dofs) 1 = free; schedule: halo_schedule()."""
        self.part, self.n_comp, self.comm = part, n_comp, comm
        self._solver = backend.prepare(A_local, free_mask_local)
        ptr, src, dst = schedule
        c = n_comp
        # node-level lists -> dof-level lists
        rep = lambda a: (c * a.astype(np.int64)[:, None] + np.arange(c)).ravel().astype(np.int32)
        dptr = (ptr.astype(np.int64) * c).astype(np.int32)
        self._lists = (np.ascontiguousarray(dptr), np.ascontiguousarray(rep(src)), np.ascontiguousarray(rep(dst)))
        # free dofs of the whole system = sum over ranks of the free dofs they OWN (every exit condition of
        # the collective kernel must be the same on all ranks)
        import torch.distributed as dist
        owned_free = free_mask_local[c * part.owned_begin:c * part.owned_end].sum(dtype=torch.int64).reshape(1)
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            if dist.get_backend(group) == 'gloo':
                owned_free = owned_free.cpu()
            dist.all_reduce(owned_free, group=group)
        self.n_free_global = int(owned_free.item())
        _lib.check(_lib.lib().fes_solver_set_dist(
            self._solver._h, comm.handle, c * part.owned_begin, c * part.owned_end,
            _lib.ptr(self._lists[0]), _lib.ptr(self._lists[1]), _lib.ptr(self._lists[2]), self.n_free_global))

    @property
    def solver(self):
        return self._solver

    def solve(self, b_local, x_local=None, warm_start=False):
        """b_local / x_local: device vectors on the local numbering; only owned rows are read /
        written.  Collective: every rank of the communicator must call it."""
        return self._solver.solve_device(b_local, x_local, warm_start=warm_start)
