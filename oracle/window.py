"""Windowed oracle parity at sizes the oracle cannot assemble whole (TEST INFRASTRUCTURE: only tests/,
smoke() and bench.py's parity gate import this).

The CSR rows of a contiguous node window [v0, v1) depend only on the elements that touch a node of the
window.  `oracle_rows` takes those elements in ascending global element id (so every slot sums its
contributions in the order np.bincount adds them, ref fem/space.py:133-145), renumbers their nodes with
the order-preserving map the node-block partition uses, assembles that sub-mesh with the oracle
(`fes_oracle.assemble` = the reference's _ScatterPlan, ref fem/space.py:107-131) and returns the
window's rows with GLOBAL column ids.  `compare_window` checks a device CSR against them: row lengths and
column ids array_equal, values rtol 1e-12 with the absolute floor 1e-12 max|A| (SURVEY A.5,
ref tests/test_assembly.py:196-218)."""
import numpy as np

import fes_oracle as fo


def oracle_rows(okind, c, coords, elem_nodes, v0, v1, ke_of):
    """(indptr, global indices, data) of rows [c*v0, c*v1).  ke_of(geometry, element_ids) -> (n, k, k)."""
    elem_nodes = np.asarray(elem_nodes)
    touch = ((elem_nodes >= v0) & (elem_nodes < v1)).any(axis=1)
    ids = np.flatnonzero(touch)                                   # ascending global element ids
    le = elem_nodes[ids]
    l2g = np.union1d(le.ravel(), np.arange(v0, v1, dtype=le.dtype))
    el = np.searchsorted(l2g, le)
    X = np.asarray(coords)[l2g]
    A = fo.assemble(el, c, len(l2g), ke_of(fo.geometry(okind, X[el]), ids))
    b0, b1 = int(np.searchsorted(l2g, v0)), int(np.searchsorted(l2g, v1 - 1)) + 1
    r0, r1 = c * b0, c * b1
    sl = slice(A.indptr[r0], A.indptr[r1])
    cols = A.indices[sl]
    return A.indptr[r0:r1 + 1] - A.indptr[r0], c * l2g[cols // c] + cols % c, A.data[sl], len(ids)


def device_rows(A, c, v0, v1):
    """The same rows of a fes_b200.DeviceCSR, copied to the host (only the window travels)."""
    r0, r1 = c * v0, c * v1
    ip = A.indptr_d[r0:r1 + 1].cpu().numpy().astype(np.int64)
    lo, hi = int(ip[0]), int(ip[-1])
    return ip - lo, A.indices_d[lo:hi].cpu().numpy().astype(np.int64), A.data_d[lo:hi].cpu().numpy()


def compare_window(A, okind, c, coords, elem_nodes, v0, v1, ke_of, scale=None, rtol=1e-12):
    """Raises AssertionError on any mismatch; returns a small record for reports."""
    ip_o, ix_o, d_o, n_el = oracle_rows(okind, c, coords, elem_nodes, v0, v1, ke_of)
    ip_d, ix_d, d_d = device_rows(A, c, v0, v1)
    assert np.array_equal(ip_d, ip_o), f'row lengths differ in node window [{v0}, {v1})'
    assert np.array_equal(ix_d, ix_o), f'column indices differ in node window [{v0}, {v1})'
    floor = rtol * (float(np.abs(d_o).max()) if scale is None else scale)
    err = np.abs(d_d - d_o) - rtol * np.abs(d_o)
    assert float(err.max()) <= floor, f'values differ in node window [{v0}, {v1}): excess {float(err.max()):.3e}'
    return {'nodes': [int(v0), int(v1)], 'rows': int(c * (v1 - v0)), 'nnz': int(len(d_o)), 'oracle_elements': int(n_el),
            'max_abs_diff': float(np.abs(d_d - d_o).max()), 'max_abs_value': float(np.abs(d_o).max())}
