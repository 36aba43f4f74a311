import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo/tests')
import fes_b200 as F, fes_oracle as fo
from test_gpu_assembly import jittered
from fes_b200.dist import NodePartition, owned_row_block
v, e = jittered('tri', seed=21)
E = np.random.default_rng(2).uniform(10.0, 300.0, len(e))
form = lambda EE: F.LinearElasticForm(F.LinearElasticMaterial(EE, 0.3))
whole = F.FunctionSpace(F.Mesh(v, e), n_components=2).assemble(form(E))
ref = fo.assemble(e, 2, len(v), fo.ke_elastic(fo.geometry(fo.P1_TRI, v[e]), E, 0.3))
print('whole vs oracle', np.abs(whole.data - ref.data).max() / np.abs(ref.data).max())
for r in range(2):
    part = NodePartition(e, len(v), r, 2)
    Vl = F.FunctionSpace(part.local_mesh(v), n_components=2)
    Al = Vl.assemble(form(E[part.local_elements]))
    refl = fo.assemble(part.elements_local, 2, part.n_local, fo.ke_elastic(fo.geometry(fo.P1_TRI, v[part.l2g][part.elements_local]), E[part.local_elements], 0.3))
    d = np.abs(Al.data - refl.data)
    print(r, 'local vs oracle', d.max() / np.abs(refl.data).max(), 'n bad', (d > 1e-9).sum(), 'of', len(d))
    ip, ix, data = owned_row_block(part, Al, 2)
    r0 = whole.indptr[2 * part.lo]; r1 = whole.indptr[2 * part.hi]
    dd = np.abs(data - whole.data[r0:r1]); print('   owned vs whole', dd.max(), (dd > 0).sum())
    if (dd > 0).any():
        k = np.flatnonzero(dd > 0)[:5]; print(k, data[k], whole.data[r0:r1][k])
print('---- more')
Vw = F.FunctionSpace(F.Mesh(v, e), n_components=2)
w1 = Vw.assemble(form(E)); w2 = Vw.assemble(form(E))
print('whole twice equal', np.array_equal(w1.data, w2.data), 'vs first whole', np.array_equal(w1.data, whole.data))
part = NodePartition(e, len(v), 0, 2)
Vl = F.FunctionSpace(part.local_mesh(v), n_components=2)
l1 = Vl.assemble(form(E[part.local_elements])); l2 = Vl.assemble(form(E[part.local_elements]))
print('local twice equal', np.array_equal(l1.data, l2.data))
# scalar-E variant: no per-element arrays
ws = Vw.assemble(form(200.0)); ls = Vl.assemble(form(200.0))
ip, ix, data = owned_row_block(part, ls, 2)
r0 = ws.indptr[2 * part.lo]; r1 = ws.indptr[2 * part.hi]
print('scalar E owned vs whole differing', (data != ws.data[r0:r1]).sum())
# volumes / element matrices
ke_w = form(E).element_matrices(Vw).cpu().numpy()[part.local_elements]
ke_l = form(E[part.local_elements]).element_matrices(Vl).cpu().numpy()
print('element matrices equal', np.array_equal(ke_w, ke_l))
print('plan tiles', Vw.plan().n_pairs, Vl.plan().n_pairs)
