"""A CSR operator resident in HBM: int32 indptr/indices shared with the plan that built it,
fp64 values.  Lazy host views give the scipy `csr_array` the reference returns (int64 index
arrays, fp64 data) for parity checks and drop-in callers."""
import numpy as np
import torch
from scipy.sparse import csr_array

from . import _lib


class DeviceCSR:
    def __init__(self, shape, indptr, indices, data, host_pattern=None):
        self.shape = tuple(shape)
        self.indptr_d, self.indices_d, self.data_d = indptr, indices, data
        self._host_pattern = host_pattern   # callable -> (indptr int64, indices int64), cached by the plan
        self._plan = None                   # the ScatterPlan owning indptr/indices, when there is one

    @property
    def nnz(self):
        return int(self.data_d.numel())

    @property
    def device(self):
        return self.data_d.device

    def host_pattern(self):
        if self._host_pattern is not None:
            return self._host_pattern()
        return (self.indptr_d.cpu().numpy().astype(np.int64), self.indices_d.cpu().numpy().astype(np.int64))

    def to_scipy(self):
        indptr, indices = self.host_pattern()
        return csr_array((self.data_d.cpu().numpy(), indices, indptr), shape=self.shape)

    def toarray(self):
        return self.to_scipy().toarray()

    @property
    def data(self):
        return self.data_d.cpu().numpy()

    @property
    def indptr(self):
        return self.host_pattern()[0]

    @property
    def indices(self):
        return self.host_pattern()[1]

    def matvec(self, x):
        """y = A x on the device (x: device fp64 tensor or ndarray -> same kind back)."""
        host = not isinstance(x, torch.Tensor)
        xd = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).to(self.device) if host else x.contiguous()
        y = torch.empty(self.shape[0], dtype=torch.float64, device=self.device)
        _lib.check(_lib.lib().fes_spmv(self.shape[0], _lib.ptr(self.indptr_d), _lib.ptr(self.indices_d),
                                       _lib.ptr(self.data_d), _lib.ptr(xd), _lib.ptr(y), _lib.stream()))
        return y.cpu().numpy() if host else y

    __matmul__ = matvec

    def _like(self, data):
        out = DeviceCSR(self.shape, self.indptr_d, self.indices_d, data, self._host_pattern)
        out._plan = self._plan
        return out

    def scaled(self, alpha):
        return self._like(self.data_d * alpha)

    def add_subpattern(self, other, alpha=1.0):
        """self + alpha*other where other's pattern is a subset of self's; returns a new DeviceCSR."""
        out = self.data_d.clone()
        _lib.check(_lib.lib().fes_csr_add_subpattern(
            self.shape[0], _lib.ptr(self.indptr_d), _lib.ptr(self.indices_d), _lib.ptr(out),
            _lib.ptr(other.indptr_d), _lib.ptr(other.indices_d), _lib.ptr(other.data_d), float(alpha), _lib.stream()))
        return self._like(out)

    @classmethod
    def from_scipy(cls, A, device):
        A = csr_array(A)
        if not A.has_sorted_indices:
            A = A.sorted_indices()
        return cls(A.shape,
                   torch.as_tensor(A.indptr.astype(np.int32)).to(device),
                   torch.as_tensor(A.indices.astype(np.int32)).to(device),
                   torch.as_tensor(np.ascontiguousarray(A.data, dtype=np.float64)).to(device))
