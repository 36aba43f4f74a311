"""ctypes binding of libfes_b200.so (the C ABI declared in include/fes_b200.h).

Torch tensors are used only as device buffers: every call passes `tensor.data_ptr()` and the
current torch CUDA stream.  There is no CPU fallback: `lib()` raises if the shared library is
missing, and `require_cuda()` raises if no CUDA device is visible.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('FES_B200_LIB') or os.path.join(_HERE, 'libfes_b200.so')   # env: A-B builds of the same library (tuning)

OK, ERR_VALUE, ERR_CUDA, ERR_UNSUPPORTED, ERR_CG, ERR_NCCL = 0, -1, -2, -3, -4, -5
P1_LINE, P1_TRI, P1_TET, P2_LINE, P2_TRI = 0, 1, 2, 3, 4
FORM_LAPLACIAN, FORM_MASS, FORM_ELASTIC = 0, 1, 2

_p, _i64, _int, _dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double


class Coeffs(C.Structure):
    _fields_ = [('E', _dbl), ('nu', _dbl), ('factor', _dbl), ('d_E', _p), ('d_scale', _p)]


# name -> (restype, argtypes); every symbol include/fes_b200.h declares
SIGNATURES = {
    'fes_version': (_int, []),
    'fes_last_error': (C.c_char_p, []),
    'fes_set_device': (_int, [_int]),
    'fes_device_info': (_int, [C.POINTER(_int), C.POINTER(_int), C.POINTER(_int), C.POINTER(_i64)]),
    'fes_launch_count': (_i64, []),
    'fes_plan_create': (_int, [_p, _i64, _int, _i64, _int, _p, C.POINTER(_p)]),
    'fes_plan_destroy': (None, [_p]),
    'fes_plan_n_dofs': (_i64, [_p]),
    'fes_plan_nnz': (_i64, [_p]),
    'fes_plan_n_pairs': (_i64, [_p]),
    'fes_plan_n_tiles': (_i64, [_p]),
    'fes_plan_n_run_tiles': (_i64, [_p]),
    'fes_plan_run_nodes': (_i64, [_p]),
    'fes_plan_bytes': (_i64, [_p]),
    'fes_plan_read_bytes': (_i64, [_p]),
    'fes_plan_indptr': (_p, [_p]),
    'fes_plan_indices': (_p, [_p]),
    'fes_plan_export_csr': (_int, [_p, _p, _p]),
    'fes_sort_pairs_u64': (_int, [_p, _p, _i64, _int, _p]),
    'fes_exclusive_scan_i32': (_int, [_p, _p, _i64, _p]),
    'fes_element_matrices': (_int, [_int, _int, _int, _int, _p, _i64, _p, C.POINTER(Coeffs), _p, _p]),
    'fes_element_volumes': (_int, [_int, _int, _p, _i64, _p, _p, _p]),
    'fes_assemble': (_int, [_p, _int, _int, _int, _p, _p, C.POINTER(Coeffs), _p, _p]),
    'fes_assemble_precomputed': (_int, [_p, _p, _p, _dbl, _p, _p]),
    'fes_assemble_host': (_int, [_p, _int, _int, _int, _p, _p, _i64, C.POINTER(Coeffs), _p, _p, _p, _p]),
    'fes_spmv': (_int, [_i64, _p, _p, _p, _p, _p, _p]),
    'fes_csr_add_subpattern': (_int, [_i64, _p, _p, _p, _p, _p, _p, _dbl, _p]),
    'fes_pow': (_int, [_p, _dbl, _i64, _p, _p]),
    'fes_solver_create': (_int, [_i64, _p, _p, _p, _p, _p, C.POINTER(_p)]),
    'fes_solver_create_host': (_int, [_i64, _p, _p, _p, C.POINTER(_p)]),
    'fes_solver_destroy': (None, [_p]),
    'fes_solver_operator_bytes': (_i64, [_p]),
    'fes_solver_use_plan': (_int, [_p, _p]),
    'fes_solver_refresh': (_int, [_p, _p]),
    'fes_solver_solve': (_int, [_p, _p, _p, _dbl, _i64, _int, C.POINTER(_i64), C.POINTER(_dbl), _p]),
    'fes_solver_solve_host': (_int, [_p, _p, _p, _dbl, _i64, C.POINTER(_i64), C.POINTER(_dbl)]),
    'fes_lift_rhs': (_int, [_p, _p, _p, _p, _p]),
    'fes_comm_create': (_int, [_int, _int, _i64, C.POINTER(_p)]),
    'fes_comm_handle': (_int, [_p, _p]),
    'fes_comm_connect': (_int, [_p, _p]),
    'fes_comm_destroy': (None, [_p]),
    'fes_solver_set_dist': (_int, [_p, _p, _i64, _i64, _p, _p, _p, _i64]),
    'fes_filter_create': (_int, [_p, _i64, _int, _p, _int, _dbl, _p, C.POINTER(_p)]),
    'fes_filter_destroy': (None, [_p]),
    'fes_filter_nnz': (_i64, [_p]),
    'fes_filter_export_csr': (_int, [_p, _p, _p, _p]),
    'fes_filter_apply': (_int, [_p, _p, _p, _p]),
    'fes_elastic_energy': (_int, [_int, _p, _i64, _p, _p, _dbl, _dbl, _p, _p, _p, _p, _p]),
    'fes_simp_sensitivity': (_int, [_p, _p, _dbl, _dbl, _i64, _p, _p]),
    'fes_oc_update': (_int, [_p, _p, _p, _i64, _dbl, _dbl, _int, _dbl, _p, C.POINTER(_int), _p]),
    'fes_dot': (_int, [_p, _p, _i64, C.POINTER(_dbl), _p]),
    'fes_node_scatter': (_int, [_p, _p, _int, _p, _p]),
    'fes_quadrature_size': (_int, [_int, _int]),
    'fes_quadrature_points': (_int, [_int, _int, _int, _p, _i64, _p, _p, _p, _p]),
    'fes_element_loads': (_int, [_int, _int, _int, _p, _i64, _p, _p, _int, _p, _p]),
    'fes_weighted_gradient_matrices': (_int, [_int, _int, _int, _int, _p, _i64, _p, _p, _p, _p, _p]),
    'fes_elastic_fields': (_int, [_int, _p, _i64, _p, _p, _p, _dbl, _dbl, _int, _p, _p, _p, _p]),
    'fes_elastic_stress_field': (_int, [_int, _p, _i64, _p, _p, _p, _dbl, _dbl, _int, _p, _p]),
    'fes_scalar_gradient': (_int, [_int, _int, _int, _p, _i64, _p, _p, _p, _p]),
    'fes_von_mises_gradient': (_int, [_int, _p, _i64, _p, _p, _p, _dbl, _dbl, _int, _p, _p, _p]),
}

_lib = None


def lib():
    """The loaded library; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                '(fes_b200 has no CPU fallback)')
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError('fes_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def check(status):
    """Map a C status to the exception the reference raises for the same condition."""
    if status == OK:
        return
    msg = lib().fes_last_error().decode('utf-8', 'replace')
    if status == ERR_VALUE:
        raise ValueError(msg)
    if status == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device (or host) pointer of a contiguous tensor / numpy array; None -> NULL."""
    if t is None:
        return C.c_void_p(0)
    if isinstance(t, torch.Tensor):
        assert t.is_contiguous()
        return C.c_void_p(t.data_ptr())
    assert t.flags['C_CONTIGUOUS']
    return C.c_void_p(t.ctypes.data)


def launch_count():
    return int(lib().fes_launch_count())
