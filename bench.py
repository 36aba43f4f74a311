"""Benchmark of the fes-b200 hot path: elements assembled/sec into CSR on BASELINE.json's
config 2 (2D P1 linear-elasticity cantilever, 3 998 792 triangles), with the HBM roofline of the
fused assembly kernel, the CPU oracle timed beside it, and (in `extra`) the Jacobi-PCG solve of
the same cantilever and SIMP design iterations/sec on config 4.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-extra]

A step is one fused assembly of the stiffness values into the cached CSR pattern (the reference's
warm `FunctionSpace.assemble`: geometry and scatter plan cached, element matrices + scatter per
call).  Under torchrun (N > 1) every rank assembles its own slab of an N-times longer cantilever
(owner-computes, ghost elements, no collective): weak scaling.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'oracle')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

RES = (2829, 708)                       # config 2: 3 998 792 triangles, 4 005 864 dofs
CORNERS = [[0.0, 0.0], [4.0, 1.0]]
E, NU = 200.0, 0.3
METRIC, UNIT = 'elements assembled/sec into CSR', 'elements/s'


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread (the timed
    region is milliseconds long, too short for `nvidia-smi -lms`); nvidia-smi is the fallback."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.sm, self.reasons, self.max_mhz = index, [], set(), None
        self.stop, self.thread, self.proc, self.rows = threading.Event(), None, None, []

    def _poll(self):
        import pynvml as nv
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        names = {getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8): 'hw_slowdown',
                 getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
                 getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
                 getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4): 'sw_power_cap'}
        get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop.is_set():
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            mask = int(get_reasons(h))
            self.reasons.update(n for bit, n in names.items() if mask & bit)
            time.sleep(0.0005)

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
            try:
                self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                              '--format=csv,noheader,nounits', '-lms', '100'],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.reader = threading.Thread(target=self._read, daemon=True)
                self.reader.start()
            except Exception:
                self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def sample_now(self):
        """One sample from the calling thread (the poller can be starved while Python enqueues launches)."""
        if self.thread:
            try:
                import pynvml as nv
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(nv.nvmlDeviceGetHandleByIndex(self.index), nv.NVML_CLOCK_SM)))
            except Exception:
                pass

    def __exit__(self, *a):
        self.stop.set()
        if self.thread:
            self.thread.join(timeout=2)
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.reader.join(timeout=2)

    def summary(self):
        if self.thread:
            return {'sm_mhz': float(np.median(self.sm)) if self.sm else None, 'sm_max_mhz': self.max_mhz,
                    'reasons': sorted(self.reasons), 'samples': len(self.sm), 'source': 'nvml polling thread'}
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, f in zip(names, r[2:6]) if f == 'Active'})
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace('.', '').isdigit()]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm), 'source': 'nvidia-smi -lms 100'}


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs NVML reports as local to GPU `index`, so the pinned host buffers of
    the end-to-end leg are first-touched on that GPU's NUMA node (matters when 8 ranks share the host)."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        words = nv.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port (numpy/scipy restatement of the reference's warm assemble)
# ---------------------------------------------------------------------------------------------
def reference_or_port_assembly(kind, res, repeats=2):
    """Warm assembly rate of the CPU arm on one of BASELINE.json's cases (bounded sizes, stated):
    the UNMODIFIED reference's FunctionSpace.assemble when its tree is reachable, else the oracle port.
    kind: 'p1_laplacian' | 'p2_elastic' | 'p2_mass' | 'tet_elastic'."""
    import fes_oracle as fo
    import ref_shim
    t_cold = time.perf_counter()
    if ref_shim.available():
        ref_shim.install()
        from fem.elements import QuadraticTriangleElement
        from fem.forms import LaplacianForm, LinearElasticForm, MassForm
        from fem.materials import LinearElasticMaterial
        from fem.mesh.mesh import Mesh
        from fem.space import FunctionSpace
        if kind == 'tet_elastic':
            v, e = fo.box_mesh([[0, 0, 0], [1, 1, 1]], res)
            V = FunctionSpace(Mesh(v, e, np.zeros((0, 3), dtype=np.int64)), n_components=3)
            form = LinearElasticForm(LinearElasticMaterial(E, NU))
        else:
            v, e = fo.rect_mesh([[0, 0], [1, 1]], res)
            mesh = Mesh(v, e, np.zeros((0, 2), dtype=np.int64))
            if kind == 'p1_laplacian':
                V, form = FunctionSpace(mesh), LaplacianForm()
            else:
                V = FunctionSpace(mesh, QuadraticTriangleElement, n_components=2)
                form = LinearElasticForm(LinearElasticMaterial(E, NU)) if kind == 'p2_elastic' else MassForm(2)
        V.assemble(form)
        cold = time.perf_counter() - t_cold
        best = min(_timed(lambda: V.assemble(form)) for _ in range(repeats))
        n_el, impl = len(V.mesh.elements), 'reference'
    else:
        if kind == 'tet_elastic':
            v, e = fo.box_mesh([[0, 0, 0], [1, 1, 1]], res)
            okind, c = fo.P1_TET, 3
        else:
            v, e = fo.rect_mesh([[0, 0], [1, 1]], res)
            okind, c = fo.P1_TRI, (1 if kind == 'p1_laplacian' else 2)
            if kind != 'p1_laplacian':
                e, v, _ = fo.p2_nodes(v, e, fo.boundary_facets(e))
                okind = fo.P2_TRI
        geo = fo.geometry(okind, v[e])
        plan = fo.scatter_plan(e, c, len(v))
        ke = {'p1_laplacian': lambda: fo.ke_laplacian(geo), 'p2_mass': lambda: fo.ke_mass(geo, 2)}.get(
            kind, lambda: fo.ke_elastic(geo, E, NU))
        cold = time.perf_counter() - t_cold
        best = min(_timed(lambda: plan.scatter(ke())) for _ in range(repeats))
        n_el, impl = len(e), 'port'
    return {'elements_per_s': n_el / best, 'elements': int(n_el), 'warm_s': best, 'cold_s': cold, 'kind': impl,
            'sample': f'{kind} on {"x".join(str(r) for r in res)} nodes, best of {repeats} warm assemblies'}


def _timed(fn):
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


def cpu_pcg_twin():
    """scipy cg(A_ff, b, rtol=1e-10, atol=0, M=diag^-1) -- the algorithmic twin of the GPU Jacobi-PCG and
    of the reference's _CGSolver call (ref fem/backends.py:201-213) -- on the 504x126-node 4:1 cantilever
    of BASELINE.md 2.3 (127 008 dofs), operator from the oracle."""
    import fes_oracle as fo
    from scipy.sparse.linalg import LinearOperator, cg
    v, e = fo.rect_mesh(CORNERS, (504, 126))
    A = fo.assemble(e, 2, len(v), fo.ke_elastic(fo.geometry(fo.P1_TRI, v[e]), E, NU))
    fixed_nodes = np.flatnonzero(v[:, 0] == 0.0)
    fixed = (2 * fixed_nodes[:, None] + np.arange(2)).ravel()
    free = np.setdiff1d(np.arange(2 * len(v)), fixed)
    b = np.zeros(2 * len(v))
    tip = np.flatnonzero(v[:, 0] == CORNERS[1][0])
    b[2 * tip + 1] = -1.0 / (len(tip) - 1)                      # a tip traction; the exact load does not matter here
    Aff, bf = A[np.ix_(free, free)].tocsr(), b[free]
    dinv = 1.0 / Aff.diagonal()
    its = [0]
    t0 = time.perf_counter()
    x, info = cg(Aff, bf, rtol=1e-10, atol=0.0, M=LinearOperator(Aff.shape, matvec=lambda r: dinv * r),
                 callback=lambda xk: its.__setitem__(0, its[0] + 1))
    dt = time.perf_counter() - t0
    return {'dofs': int(Aff.shape[0]), 'iterations': its[0], 'seconds': dt, 'us_per_iteration': 1e6 * dt / max(its[0], 1),
            'info': int(info), 'impl': 'scipy.sparse.linalg.cg with M = diag(A)^-1, one thread',
            'sample': '504x126-node cantilever (BASELINE.md 2.3), 127 008 dofs'}


def cpu_simp_iterations(n_iters, budget_s):
    """Design iterations of the UNMODIFIED reference's TopologyOptimizer (ref fem/topology.py:291-317) at
    config 4 (1201x401 nodes, 960 000 triangles) on the host cores: constructor once, then up to n_iters
    iterations within budget_s.  None when the reference tree is absent."""
    import ref_shim
    if not ref_shim.available():
        return None
    ref_shim.install()
    import fes_oracle as fo
    from fem.boundary import BCType, BoundaryConditions
    from fem.equations import LinearElastic
    from fem.mesh.mesh import Mesh
    from fem.regions import in_box, intersect, on_plane
    from fem.topology import TopologyOptimizer
    import fes_b200.mesh as fm
    w, h = 3.0, 1.0
    v, e = fo.rect_mesh([[0, 0], [w, h]], (1201, 401))
    mesh = Mesh(v, e, fm.rect_boundary((1201, 401)))
    bc = BoundaryConditions()
    bottom, top = on_plane(1, 0.0), on_plane(1, h)
    bc.add(BCType.DIRICHLET, intersect(bottom, in_box([None, None], [0.04 * w, None])), [0, 0])
    bc.add(BCType.DIRICHLET, intersect(bottom, in_box([0.96 * w, None], [None, None])), [None, 0])
    bc.add(BCType.NEUMANN, intersect(top, in_box([0.4 * w, None], [0.6 * w, None])), [0, -0.5])
    t0 = time.perf_counter()
    topo = TopologyOptimizer(mesh, LinearElastic(200.0, 0.4), bc, iters=1, volume_frac=0.5,
                             smoothing_radius=0.00625, penalty=3.0)
    ctor = time.perf_counter() - t0
    times, comps = [], []
    for _ in range(n_iters):
        t0 = time.perf_counter()
        hist = topo.solve()
        times.append(time.perf_counter() - t0)
        comps.append(float(np.sum(hist.compliance[-1])))
        if sum(times) + ctor + times[-1] > budget_s:
            break
    return {'iterations_per_s': len(times) / sum(times), 'iterations_timed': len(times), 'seconds_per_iteration': times,
            'constructor_s': ctor, 'total_compliance': comps, 'kind': 'reference',
            'impl': 'TopologyOptimizer.solve of the unmodified reference (DirectBackend = SuperLU), one design iteration per call'}


def cpu_assemble_rate(sample_res=(1001, 251), repeats=3):
    """Warm `assemble` of the oracle on a bounded sample of the same workload: the scatter plan
    and geometry are built once (as the reference caches them), then element matrices + bincount
    scatter are timed.  Returns (elements/s, description)."""
    import fes_oracle as fo
    v, e = fo.rect_mesh(CORNERS, sample_res)
    geo = fo.geometry(fo.P1_TRI, v[e])
    plan = fo.scatter_plan(e, 2, len(v))
    best = float('inf')
    for _ in range(repeats):
        t0 = time.perf_counter()
        plan.scatter(fo.ke_elastic(geo, E, NU))
        best = min(best, time.perf_counter() - t0)
    return len(e) / best, f'{len(e)} triangles ({sample_res[0]}x{sample_res[1]} nodes of the same cantilever), best of {repeats} warm assemblies'


def blas_threads():
    """Threads the numpy/scipy BLAS pools may use (the reference's hot calls -- einsum, argsort,
    bincount, SuperLU -- are effectively single-threaded whatever this says)."""
    try:
        from threadpoolctl import threadpool_info
        return max([int(p.get('num_threads', 1)) for p in threadpool_info()] or [1])
    except Exception:
        return int(os.environ.get('OMP_NUM_THREADS', os.cpu_count() or 1))


def reference_space(res=RES):
    """The UNMODIFIED reference (`baseline/_ref`, or /root/reference in the build container)
    set up for config 2: FunctionSpace + LinearElasticForm on the cantilever mesh.  The mesh
    arrays come from the vectorised generator (array_equal to the reference's create_rect_mesh,
    tests/test_host_logic.py) because the reference's Python-loop generator needs ~20 s at this
    size and is not on the timed path."""
    import ref_shim
    ref_shim.install()
    import fes_oracle as fo
    from fem.forms import LinearElasticForm
    from fem.materials import LinearElasticMaterial
    from fem.mesh.mesh import Mesh
    from fem.space import FunctionSpace
    v, e = fo.rect_mesh(CORNERS, res)
    mesh = Mesh(v, e, np.zeros((0, 2), dtype=np.int64))
    return FunctionSpace(mesh, n_components=2), LinearElasticForm(LinearElasticMaterial(E, NU)), len(e)


def cpu_baseline_sample(sample_res=(1001, 251)):
    """`cpu_baseline` of the GPU arm: a bounded sample of config 2 (500 000 triangles of the same cantilever),
    the unmodified reference's warm FunctionSpace.assemble when its tree is reachable (kind "reference"),
    else the oracle port (kind "port").  The full-size figure is `bench.py --impl reference`."""
    import ref_shim
    if ref_shim.available():
        V, form, n_el = reference_space(sample_res)
        c0, t0 = time.process_time(), time.perf_counter()
        V.assemble(form)
        best = min(_timed(lambda: V.assemble(form)) for _ in range(3))
        wall = time.perf_counter() - t0
        return {'value': n_el / best, 'unit': UNIT, 'cores': max(1, int(round((time.process_time() - c0) / wall))),
                'kind': 'reference', 'host_cpus': os.cpu_count(), 'blas_threads': blas_threads(),
                'sample': f'{n_el} triangles ({sample_res[0]}x{sample_res[1]} nodes of the same cantilever): warm '
                          'FunctionSpace.assemble(LinearElasticForm) of the unmodified reference, best of 3'}
    rate, sample = cpu_assemble_rate(sample_res)
    return {'value': rate, 'unit': UNIT, 'cores': 1, 'kind': 'port', 'sample': sample, 'host_cpus': os.cpu_count()}


def run_reference(args):
    """`--impl reference`: the reference's own warm `FunctionSpace.assemble` (ref fem/space.py:679-696:
    geometry and scatter plan cached by the first call, element matrices + bincount scatter per
    call) at the FULL config-2 size on the host cores; the oracle port on a 500 000-triangle
    sample only when the reference tree is absent."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import ref_shim
    n_el = 2 * (RES[0] - 1) * (RES[1] - 1)
    times, t_all, budget, cores_used = [], time.perf_counter(), 420.0, 1.0
    if ref_shim.available():
        V, form, n_el = reference_space()
        t0 = time.perf_counter()
        V.assemble(form)                                  # cold: geometry + _ScatterPlan.build + scatter
        cold_s = time.perf_counter() - t0
        for _ in range(max(0, args.warmup - 1)):
            V.assemble(form)
        c0 = time.process_time()
        for _ in range(args.steps):
            t0 = time.perf_counter()
            V.assemble(form)
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_all > budget:
                break
        cores_used = (time.process_time() - c0) / max(sum(times), 1e-9)   # CPU seconds per wall second
        ms_step = 1e3 * float(np.mean(times))
        kind = 'reference'
        sample = (f'full config 2 ({n_el} triangles): FunctionSpace.assemble(LinearElasticForm) of the unmodified '
                  f'reference, warm (plan + geometry cached; cold first call {cold_s:.1f} s), mean of {len(times)} calls')
        workload = ('config 2: 2D P1 linear-elasticity cantilever, create_rect_mesh([[0,0],[4,1]], (2829,708)) = '
                    '3 998 792 triangles, 4 005 864 dofs, E=200 nu=0.3')
    else:
        rates = []
        for _ in range(max(1, args.warmup)):
            cpu_assemble_rate(repeats=1)
        for _ in range(args.steps):
            r, sample = cpu_assemble_rate(repeats=1)
            rates.append(r)
            times.append(0.0)
            if time.perf_counter() - t_all > 240:
                break
        ms_step = 1e3 * n_el / float(np.mean(rates))
        kind = 'port'
        workload = ('config 2: 2D P1 linear-elasticity cantilever, 3 998 792 triangles (reference tree absent: oracle '
                    'port timed on a 500 000-triangle sample, rate per element)')
    value = n_el / (ms_step * 1e-3)
    out = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
           'steps': len(times), 'warmup': args.warmup, 'ms_per_step': ms_step,
           'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
           'config': {'workload': workload},
           'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': max(1, int(round(cores_used))), 'kind': kind,
                            'sample': sample, 'cpu_seconds_per_wall_second': round(cores_used, 2),
                            'host_cpus': os.cpu_count(), 'blas_threads': blas_threads(),
                            'omp_num_threads_env': os.environ.get('OMP_NUM_THREADS')},
           'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def slab_mesh(rank, world):
    """Rank `rank`'s slab of a cantilever `world` times longer: the nodes it owns plus one ghost
    column of cells on each interior side (owner-computes assembly needs no communication)."""
    import fes_b200 as F
    if world == 1:
        return F.create_rect_mesh(CORNERS, RES)
    corners = [[CORNERS[0][0] + rank * 4.0 - (4.0 / (RES[0] - 1)) * (rank > 0), 0.0],
               [CORNERS[0][0] + (rank + 1) * 4.0 + (4.0 / (RES[0] - 1)) * (rank < world - 1), 1.0]]
    nx = RES[0] + (rank > 0) + (rank < world - 1)
    return F.create_rect_mesh(corners, (nx, RES[1]))


def parity_gate(V, plan, values, mesh):
    """The parity gate of SURVEY 8(d), inside the bench job: the CSR the timed kernel just wrote is compared
    with the CPU oracle on three node windows of THIS mesh (first mesh rows with the boundary stencils, the
    middle, the last rows): row lengths and column indices array_equal, values rtol 1e-12 with the absolute
    floor 1e-12 max|A| (ref tests/test_assembly.py:196-218).  Any mismatch raises."""
    import types
    import fes_oracle as fo
    import window as ow
    A = types.SimpleNamespace(indptr_d=plan.indptr_d, indices_d=plan.indices_d, data_d=values)
    nx, n_nodes = RES[0] if len(mesh.vertices) == RES[0] * RES[1] else len(mesh.vertices) // RES[1], len(mesh.vertices)
    scale = float(values.abs().max())
    ke = lambda geo, ids: fo.ke_elastic(geo, E, NU)
    t0 = time.perf_counter()
    windows = [ow.compare_window(A, fo.P1_TRI, 2, mesh.vertices, mesh.elements, v0, min(v0 + 3 * nx, n_nodes), ke,
                                 scale=scale)
               for v0 in (0, (n_nodes // 2 // nx) * nx - 17, n_nodes - 3 * nx)]
    return {'status': 'ok', 'against': 'CPU oracle (oracle/window.py), rows of 3 node windows of the timed mesh',
            'pattern': 'indptr differences and indices array_equal', 'values': 'rtol 1e-12 + 1e-12 max|A|',
            'rows_compared': int(sum(w['rows'] for w in windows)), 'nnz_compared': int(sum(w['nnz'] for w in windows)),
            'max_abs_diff': max(w['max_abs_diff'] for w in windows), 'max_abs_value': scale,
            'seconds': time.perf_counter() - t0}


def pinned_copy_probe(n_bytes, dev, world, reps=3):
    """Plain pinned host<->device copies of the e2e leg's sizes, all ranks at once: what the host fabric gives
    each GPU when N of them copy concurrently (the e2e leg is bounded by exactly this)."""
    import torch
    import torch.distributed as dist
    h = torch.empty(n_bytes // 8, dtype=torch.float64).pin_memory()
    d = torch.empty(n_bytes // 8, dtype=torch.float64, device=dev)
    out = {}
    for name, (dst, src) in (('d2h', (h, d)), ('h2d', (d, h))):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[f'pinned_{name}_gbs_per_gpu_slowest_rank'] = n_bytes / float(t.item()) / 1e9
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import fes_b200 as F
    from fes_b200 import _lib

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    local_cpus = bind_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')  # stdout carries the one JSON line only
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)
    L = _lib.lib()

    mesh = slab_mesh(rank, world)
    n_el, n_nodes = len(mesh.elements), len(mesh.vertices)
    V = F.FunctionSpace(mesh, n_components=2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    plan = V.plan()
    torch.cuda.synchronize()
    plan_s = time.perf_counter() - t0
    form = F.LinearElasticForm(F.LinearElasticMaterial(E, NU))
    values = torch.empty(plan.nnz, dtype=torch.float64, device=dev)
    cf = form.lower(V).coeffs()

    def step():
        _lib.check(L.fes_assemble(plan.handle, _lib.P1_TRI, _lib.FORM_ELASTIC, 2, _lib.ptr(V._en_d),
                                  _lib.ptr(V._coords_d), C.byref(cf), _lib.ptr(values), _lib.stream()))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    parity = parity_gate(V, plan, values, mesh) if rank == 0 else None   # raises on any mismatch: no number without parity
    barrier()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        clocks.sample_now()      # the queue is still draining: a sample under load even for a short region
        barrier()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_step = ms / args.steps
    total_el = n_el * world if world == 1 else int(2 * (RES[0] - 1) * (RES[1] - 1) * world)
    value = total_el / (ms_step * 1e-3)

    # roofline of the fused assembly (algorithmic bytes per SURVEY 8(d) over the whole step).  A step is two launches
    # on a structured mesh: k_assemble_runs (the stencil-run tiles, 99.9 % of the nodes of config 2) on the caller's
    # stream and k_assemble_tiles (the remainder: the two end nodes of every mesh row) concurrently on a side stream;
    # `achieved` divides by the step time, so it charges the dominant kernel with everything.
    alg_bytes = 4 * 3 * n_el + 8 * 2 * n_nodes + 8 * plan.nnz
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (ms_step * 1e-3) / 1e9
    runs = plan.n_run_tiles > 0
    kernel = 'k_assemble_runs<P1_TRI,ELASTIC,2,2,128>' if runs else 'k_assemble_tiles<P1_TRI,ELASTIC,2,2,128>'
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):   # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full, this round
        traffic = json.load(open(tpath)).get(('k_assemble_runs' if runs else 'k_assemble_tiles') + '_c2_bytes_per_launch')
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic, 'kernel': kernel, 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': alg_bytes, 'bytes_per_element': alg_bytes / n_el,
                'plan_bytes_read_per_launch': plan.read_bytes,
                'run_tiles': plan.n_run_tiles, 'nodes_in_run_tiles': plan.run_nodes, 'nodes': n_nodes}

    # end to end through the C ABI with HOST buffers: H2D coordinates, assemble, D2H values
    h_coords = torch.as_tensor(mesh.vertices).pin_memory()
    h_values = torch.empty(plan.nnz, dtype=torch.float64).pin_memory()
    cf_host = _lib.Coeffs(E, NU, 1.0, None, None)

    def e2e_step():
        _lib.check(L.fes_assemble_host(plan.handle, _lib.P1_TRI, _lib.FORM_ELASTIC, 2, _lib.ptr(V._en_d),
                                       C.c_void_p(h_coords.data_ptr()), n_nodes, C.byref(cf_host), None, None,
                                       C.c_void_p(h_values.data_ptr()), _lib.stream()))
    e2e_step()
    barrier()
    k_e2e = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(k_e2e):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / k_e2e
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {'value': total_el / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': int(h_coords.numel() * 8),
           'd2h_bytes_per_step': int(plan.nnz * 8), 'ms_per_step': e2e_s * 1e3,
           'api': 'fes_assemble_host (pinned host coordinates in, pinned host CSR values out)',
           'host_cpus_bound_to_gpu_numa_node': local_cpus}
    del h_values
    probe = pinned_copy_probe(int(plan.nnz * 8), dev, world)
    e2e.update(probe)
    e2e['copy_bound_ms_per_step'] = 1e3 * (plan.nnz * 8 / (probe['pinned_d2h_gbs_per_gpu_slowest_rank'] * 1e9)
                                           + h_coords.numel() * 8 / (probe['pinned_h2d_gbs_per_gpu_slowest_rank'] * 1e9))
    e2e['limiter'] = ('host<->device copies: the step is the D2H of the CSR values plus the H2D of the coordinates at the '
                      'bandwidth the probe measures with all ranks copying at once; the kernel is %.1f %% of the step'
                      % (100.0 * ms_step / (e2e_s * 1e3)))

    out = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
           'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
           'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
           'config': {'workload': 'config 2: 2D P1 linear-elasticity cantilever, create_rect_mesh([[0,0],[4,1]], '
                                  '(2829,708)) = 3 998 792 triangles, 4 005 864 dofs, E=200 nu=0.3'
                                  + (f'; {world} slabs side by side, one per GPU, ghost-element owner-computes' if world > 1 else ''),
                      'n_elements_per_gpu': n_el, 'nnz_per_gpu': plan.nnz,
                      'l2': 'inputs+outputs per step (0.5 GB) exceed the 126 MB L2; no flush needed'},
           'roofline': roofline, 'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clocks.summary(),
           'parity': parity}

    extra = {'plan_build_s': plan_s, 'plan_device_bytes': plan.bytes}
    if rank == 0 and world == 1:
        out['cpu_baseline'] = cpu_baseline_sample()
        if not args.no_extra:
            extra.update(extra_metrics(F, V, mesh, dev, args.cpu_budget, args.cpu_simp_iters))
    if not args.no_extra:
        del values
        torch.cuda.empty_cache()
        try:
            extra['config5_tets'] = config5_metrics(F, world, rank, dev, n=args.tet_n)
        except Exception as exc:   # reported, never silently dropped
            extra['config5_tets'] = {'error': f'{type(exc).__name__}: {exc}'[:300]}
    out['extra'] = extra
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def config5_metrics(F, world, rank, dev, n=151, pcg_iters=300):
    """Config 5: 3D P1 elasticity on create_box_mesh(n^3 nodes) = 6(n-1)^3 tets, sharded by node
    blocks over `world` GPUs (strong scaling).  Times the owner-computes assembly (no collective)
    and the distributed Jacobi-PCG (halo + reductions inside the persistent kernel)."""
    import torch
    import torch.distributed as dist
    from fes_b200 import _lib
    from fes_b200.dist import (Communicator, DistributedSystem, NodePartition, all_to_all_objects,
                               halo_schedule)
    from fes_b200.mesh import boundary_facets, box_arrays
    t0 = time.perf_counter()
    verts, elems = box_arrays([[0, 0, 0], [1, 1, 1]], (n, n, n))
    part = NodePartition(elems, len(verts), rank, world)
    # only the two loaded/clamped faces matter for the boundary data: facets of the first and last
    # cell layers lying in the planes x = 0 and x = 1
    plane = n * n
    faces = []
    for layer, lo in ((slice(0, 6 * (n - 1) ** 2), 0), (slice(len(elems) - 6 * (n - 1) ** 2, len(elems)), len(verts) - plane)):
        f = boundary_facets(elems[layer])
        faces.append(f[((f >= lo) & (f < lo + plane)).all(axis=1)])
    mesh_l = part.local_mesh(verts, np.concatenate(faces))
    host_s = time.perf_counter() - t0
    n_el_global, n_el_local = len(elems), len(part.local_elements)
    elems_keep = elems if world == 1 else None       # the one-GPU parity gate needs the connectivity
    del elems
    Vl = F.FunctionSpace(mesh_l, n_components=3)
    form = F.LinearElasticForm(F.LinearElasticMaterial(200.0, 0.3))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    plan = Vl.plan()
    torch.cuda.synchronize()
    plan_s = time.perf_counter() - t0
    A = Vl.assemble_device(form)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    reps = 5
    ev0.record()
    for _ in range(reps):
        Vl.assemble_device(form, out=A.data_d)
    ev1.record()
    torch.cuda.synchronize()
    asm_ms = torch.tensor([ev0.elapsed_time(ev1) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(asm_ms, op=dist.ReduceOp.MAX)
    asm_ms = float(asm_ms.item())
    alg = 4 * 4 * n_el_local + 8 * 3 * len(mesh_l.vertices) + 8 * plan.nnz
    # cantilever: clamp x = 0, traction [0,-5,0] on x = 1 (ref tests/test_backends.py:109-113)
    from fes_b200.regions import on_plane
    bc = F.BoundaryConditions()
    bc.add(F.BCType.DIRICHLET, on_plane(0, 0.0), [0, 0, 0])
    bc.add(F.BCType.NEUMANN, on_plane(0, 1.0), [0, -5, 0])
    prob = F.LinearProblem(Vl, form, None, bc)
    free, fixed, vals = prob.constraints
    mask = torch.zeros(Vl.n_dofs, dtype=torch.uint8, device=dev)
    mask[torch.as_tensor(free, device=dev)] = 1
    backend = F.JacobiPCGBackend(rtol=1e-30, maxiter=pcg_iters)
    x = torch.zeros(Vl.n_dofs, dtype=torch.float64, device=dev)
    if world > 1:
        md = torch.tensor([Vl.n_dofs], device=dev)
        dist.all_reduce(md, op=dist.ReduceOp.MAX)
        comm = Communicator(int(md.item()))
        system = DistributedSystem(part, A, 3, mask, comm, halo_schedule(part, all_to_all_objects), backend)
        solve = lambda: system.solve(prob.load_d, x)
    else:
        solver = backend.prepare(A, mask)
        solve = lambda: solver.solve_device(prob.load_d, x)
    best = float('inf')
    for _ in range(2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        try:
            solve()
        except RuntimeError:
            pass                      # capped iteration count: "CG failed" is expected here
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    t = torch.tensor([best], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    us_iter = 1e6 * float(t.item()) / pcg_iters
    # parity, where the driver sees it.  N = 1: the rows the timed kernel wrote against the CPU oracle on node
    # windows of THIS mesh.  N > 1: a reduced box (41^3) solved to rtol 1e-10 on the N ranks and on rank 0
    # alone -- row blocks array_equal to the one-GPU pattern, solutions within 1e-9 -- see tools/dist_check.py.
    if world == 1:
        import types
        import fes_oracle as fo
        import window as ow
        smax = float(A.data_d.abs().max())
        ke = lambda geo, ids: fo.ke_elastic(geo, 200.0, 0.3)
        nn = len(verts)
        wins = [ow.compare_window(A, fo.P1_TET, 3, verts, elems_keep, v0, v1, ke, scale=smax)
                for v0, v1 in ((0, n * n // 2), (nn // 2, nn // 2 + n * n // 2), (nn - 10 * n, nn))]
        parity = {'status': 'ok', 'against': 'CPU oracle on 3 node windows of the timed mesh (pattern array_equal, values 1e-12 + floor)',
                  'rows_compared': int(sum(w['rows'] for w in wins)), 'nnz_compared': int(sum(w['nnz'] for w in wins)),
                  'max_abs_diff': max(w['max_abs_diff'] for w in wins), 'max_abs_value': smax}
    else:
        sys.path.insert(0, os.path.join(ROOT, 'tools'))
        import dist_check
        del A, x
        torch.cuda.empty_cache()
        rec, ok = dist_check.run_case('tet', 41, rank, world, dev)
        parity = rec
        flag = torch.tensor([int(ok)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if not bool(flag.item()):
            raise RuntimeError(f'distributed PCG parity failed: {rec}')
        if rank == 0:
            parity = dict(rec, status='ok')
    return {'n_nodes_per_side': n, 'elements': n_el_global, 'elements_per_gpu': n_el_local, 'dofs': 3 * n ** 3,
            'nnz_per_gpu': plan.nnz, 'assembly_ms': asm_ms, 'elements_per_s': n_el_global / (asm_ms * 1e-3),
            'assembly_hbm_frac_of_measured': alg / (asm_ms * 1e-3) / 1e9 / measured_peak()[0],
            'plan_build_s': plan_s, 'host_mesh_partition_s': host_s, 'pcg_us_per_iteration': us_iter,
            'pcg_iterations_timed': pcg_iters, 'scaling': 'strong', 'parity': parity,
            'pcg_loop': 'classic (3 barriers)' if world == 1 else
                        'classic recurrences, halo shipped inside the p update, interior rows first, mailbox reductions (distributed default)'}


def assembly_case(F, V, form, reps=20):
    """Warm fused assembly of `form` on space V: ms per assembly (CUDA events), elements/s and the
    algorithmic-byte fraction of the measured HBM peak (SURVEY 8d byte formula)."""
    import torch
    plan = V.plan()
    A = V.assemble_device(form)
    for _ in range(3):
        V.assemble_device(form, out=A.data_d)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(reps):
        V.assemble_device(form, out=A.data_d)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    n_el, N = len(V.element_nodes), V.element_nodes.shape[1]
    alg = 4 * N * n_el + 8 * V.spatial_dim * V.n_nodes + 8 * plan.nnz
    return {'ms': ms, 'elements': n_el, 'nnz': plan.nnz, 'elements_per_s': n_el / (ms * 1e-3),
            'algorithmic_bytes': alg, 'hbm_frac_of_measured': alg / (ms * 1e-3) / 1e9 / measured_peak()[0]}


def graph_replay_case(F, V, form, launches=200, replays=5):
    """Config 1 is 0.7 us of HBM traffic -- below a launch latency -- so its roofline fraction is quoted from a
    CUDA-graph replay of many reassemblies: `launches` fused assemblies captured once, replayed."""
    import torch
    from fes_b200 import _lib
    plan = V.plan()
    A = V.assemble_device(form)
    low = form.lower(V, False)
    cf = low.coeffs()
    L = _lib.lib()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            V.assemble_device(form, out=A.data_d)
        side.synchronize()
        with torch.cuda.graph(g, stream=side):
            for _ in range(launches):
                _lib.check(L.fes_assemble(plan.handle, V.element_type.KIND, low.form, V.spatial_dim, _lib.ptr(V._en_d),
                                          _lib.ptr(V._coords_d), C.byref(cf), _lib.ptr(A.data_d), _lib.stream()))
        g.replay()
        side.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(side)
        for _ in range(replays):
            g.replay()
        ev1.record(side)
        side.synchronize()
    torch.cuda.current_stream().wait_stream(side)
    ms = ev0.elapsed_time(ev1) / (replays * launches)
    n_el, N = len(V.element_nodes), V.element_nodes.shape[1]
    alg = 4 * N * n_el + 8 * V.spatial_dim * V.n_nodes + 8 * plan.nnz
    return {'ms_per_assembly_graph_replay': ms, 'launches_per_graph': launches, 'replays': replays,
            'elements_per_s_graph_replay': n_el / (ms * 1e-3),
            'hbm_frac_of_measured_graph_replay': alg / (ms * 1e-3) / 1e9 / measured_peak()[0],
            'note': 'the working set (4.8 MB) stays in L2 between replays: a launch-/latency-bound figure, not an HBM one'}


def extra_metrics(F, V, mesh, dev, cpu_budget_s=240.0, cpu_simp_iters=1):
    """Secondary numbers of BASELINE.json: configs 1 and 3 (assembly only), the PCG solve of
    config 2 and SIMP iterations/sec on config 4 (1201x401 nodes, 960 000 triangles, MBB beam, all 100
    design iterations), each with the CPU arm timed beside it in this run (SURVEY 8d i-v; bounded samples,
    stated) as long as the CPU budget lasts."""
    import torch
    from fes_b200.regions import in_box, intersect, on_plane
    out = {}
    t_cpu = [0.0]

    def cpu(fn, *a):
        if t_cpu[0] > cpu_budget_s:
            return {'skipped': f'CPU budget of {cpu_budget_s:.0f} s spent'}
        t0 = time.perf_counter()
        try:
            r = fn(*a)
        except Exception as exc:      # reported, never silently dropped
            r = {'error': f'{type(exc).__name__}: {exc}'[:300]}
        t_cpu[0] += time.perf_counter() - t0
        return r

    V1 = F.FunctionSpace(F.create_rect_mesh([[0, 0], [1, 1]], (225, 225)))
    c1 = assembly_case(F, V1, F.LaplacianForm())
    try:
        c1.update(graph_replay_case(F, V1, F.LaplacianForm()))
    except Exception as exc:
        c1['graph_replay_error'] = f'{type(exc).__name__}: {exc}'[:300]
    c1['cpu'] = cpu(reference_or_port_assembly, 'p1_laplacian', (225, 225), 3)
    out['config1_p1_laplacian_225x225'] = c1
    del V1
    V3 = F.FunctionSpace(F.create_rect_mesh([[0, 0], [1, 1]], (708, 708)), F.QuadraticTriangleElement, n_components=2)
    out['config3_p2_elastic_708x708'] = assembly_case(F, V3, F.LinearElasticForm(F.LinearElasticMaterial(E, NU)))
    out['config3_p2_mass_708x708'] = assembly_case(F, V3, F.MassForm(2))
    del V3
    torch.cuda.empty_cache()
    out['config3_p2_elastic_708x708']['cpu'] = cpu(reference_or_port_assembly, 'p2_elastic', (355, 355), 1)
    out['config3_p2_mass_708x708']['cpu'] = cpu(reference_or_port_assembly, 'p2_mass', (355, 355), 1)
    bc = F.BoundaryConditions()
    bc.add(F.BCType.DIRICHLET, on_plane(0, 0.0), [0, 0])
    bc.add(F.BCType.NEUMANN, on_plane(0, 4.0), [0, -1])
    prob = F.LinearProblem(V, F.LinearElasticForm(F.LinearElasticMaterial(E, NU)), None, bc)
    system = F.DiscreteSystem(prob.tangent(None), prob.constraints, F.JacobiPCGBackend(rtol=1e-10))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    system.solve(prob.load_d)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    its = system.solver.last_iterations
    nnz, n = prob.tangent(None).nnz, V.n_dofs
    s_it = dt / max(its, 1)
    csr_bytes = 12 * nnz + 4 * (n + 1) + 80 * n                       # scalar CSR + 10 vector passes
    sell_bytes = system.solver.operator_bytes + 14 * 8 * n             # what the kernel reads: block-SELL copy + 14 vector passes
    out['pcg_config2'] = {'dofs': n, 'iterations': its, 'seconds': dt, 'us_per_iteration': 1e6 * s_it,
                          'relres': system.solver.last_relres,
                          'hbm_frac_of_measured': (csr_bytes / s_it / 1e9) / measured_peak()[0],
                          'hbm_frac_of_measured_block_sell_bytes': (sell_bytes / s_it / 1e9) / measured_peak()[0],
                          'bytes_per_iteration_csr_formula': csr_bytes, 'bytes_per_iteration_block_sell': sell_bytes,
                          'note': 'first figure: SURVEY 8(d) scalar-CSR byte formula; second: the bytes the kernel streams '
                                  '(9 B per stored value in block-SELL, vectors partly L2-resident)'}
    del system, prob
    out['pcg_config2']['cpu'] = cpu(cpu_pcg_twin)
    w, h = 3.0, 1.0
    m4 = F.create_rect_mesh([[0, 0], [w, h]], (1201, 401))
    bc = F.BoundaryConditions()
    bottom, top = on_plane(1, 0.0), on_plane(1, h)
    bc.add(F.BCType.DIRICHLET, intersect(bottom, in_box([None, None], [0.04 * w, None])), [0, 0])
    bc.add(F.BCType.DIRICHLET, intersect(bottom, in_box([0.96 * w, None], [None, None])), [None, 0])
    bc.add(F.BCType.NEUMANN, intersect(top, in_box([0.4 * w, None], [0.6 * w, None])), [0, -0.5])
    n_it = 100                                                          # config 4 as BASELINE.json states it
    topo = F.TopologyOptimizer(m4, F.LinearElastic(200.0, 0.4), bc, iters=n_it, volume_frac=0.5,
                               smoothing_radius=0.00625, penalty=3.0, history=None)
    maxiter = 10 * int(topo._mask.sum().item())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    comps, lap = [], []
    for _ in range(n_it):
        comps.append(topo.step().total_compliance)                      # raises RuntimeError('CG failed ...') at maxiter
        lap.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    cg = topo.cg_iterations
    out['simp_config4'] = {'elements': len(m4.elements), 'iterations_per_s': n_it / dt, 'design_iterations_timed': n_it,
                           'seconds_total': dt, 'iterations_per_s_first_10': 10 / lap[9],
                           'iterations_per_s_last_10': 10 / (lap[-1] - lap[-11]),
                           'cg_iterations': cg, 'cg_iterations_total': int(sum(cg)), 'cg_iterations_max': int(max(cg)),
                           'cg_maxiter': maxiter, 'any_solve_hit_maxiter': bool(max(cg) >= maxiter),
                           'total_compliance_first_last': [comps[0], comps[-1]], 'total_compliance': comps,
                           'volume_fraction_final': topo._volume_fraction(topo.rho),
                           'warm_start': True}
    del topo
    torch.cuda.empty_cache()
    out['simp_config4']['cpu'] = cpu(cpu_simp_iterations, cpu_simp_iters, max(60.0, cpu_budget_s - t_cpu[0]))
    out['cpu_seconds_spent_on_baselines'] = t_cpu[0]
    return out


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--no-extra', action='store_true')
    ap.add_argument('--tet-n', type=int, default=151, help='nodes per side of the config-5 tet box')
    ap.add_argument('--cpu-simp-iters', type=int, default=1,
                    help='design iterations of the reference TopologyOptimizer timed in extra (about 50 s each on config 4)')
    ap.add_argument('--cpu-budget', type=float, default=240.0, help='seconds of host time for the CPU arms in extra')
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl != 'reference' else a.warmup
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_gpu(a)
