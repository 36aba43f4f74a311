"""The windowed oracle (oracle/window.py) used for parity at BASELINE.json's full sizes: the rows it
assembles from the elements touching a node window are, bit for bit, the rows of the whole operator."""
import numpy as np
import pytest

import fes_oracle as fo
import window as ow


@pytest.mark.parametrize('kind', ['tri', 'tet', 'p2'])
def test_window_rows_equal_the_whole_operator(kind):
    rng = np.random.default_rng(4)
    if kind == 'tet':
        v, e = fo.box_mesh([[0, 0, 0], [1, 2, 1]], (6, 7, 5))
        okind, c = fo.P1_TET, 3
    else:
        v, e = fo.rect_mesh([[0, 0], [3, 1]], (19, 11))
        okind, c = fo.P1_TRI, 2
    v = v + 0.02 * rng.standard_normal(v.shape)
    e = e[rng.permutation(len(e))]
    if kind == 'p2':
        e, v, _ = fo.p2_nodes(v, e, fo.boundary_facets(e))
        okind = fo.P2_TRI
    E = rng.uniform(10.0, 300.0, len(e))
    whole = fo.assemble(e, c, len(v), fo.ke_elastic(fo.geometry(okind, v[e]), E, 0.3))
    for v0, v1 in ((0, 7), (len(v) // 2, len(v) // 2 + 23), (len(v) - 5, len(v))):
        ip, ix, data, n_el = ow.oracle_rows(okind, c, v, e, v0, v1, lambda g, ids: fo.ke_elastic(g, E[ids], 0.3))
        r0, r1 = c * v0, c * v1
        sl = slice(whole.indptr[r0], whole.indptr[r1])
        np.testing.assert_array_equal(ip, whole.indptr[r0:r1 + 1] - whole.indptr[r0])
        np.testing.assert_array_equal(ix, whole.indices[sl])
        np.testing.assert_array_equal(data, whole.data[sl])
        assert 0 < n_el < len(e)
