"""Host<->device plumbing shared by the façade classes: NumPy in -> NumPy out, torch CUDA tensors
in -> torch out (no copies), modulator tables cached on the device.  torch is used for device
memory and streams only; all arithmetic happens in libb200phy."""
import ctypes as C

import numpy as np

from . import _lib


def is_torch(x):
    return type(x).__module__.startswith('torch')


def to_device(x, np_dtype):
    """-> (contiguous CUDA tensor, was_numpy).  np_dtype: numpy dtype to cast host inputs to."""
    torch = _lib.torch_cuda()
    if is_torch(x):
        t = x if x.is_cuda else x.cuda()
        tdt = getattr(torch, np.dtype(np_dtype).name)
        if t.dtype != tdt:
            t = t.to(tdt)
        return t.contiguous(), False
    a = np.ascontiguousarray(np.asarray(x), dtype=np_dtype)
    return torch.from_numpy(a).cuda(), True


def from_device(t, was_numpy):
    return t.cpu().numpy() if was_numpy else t


def complex_np(dtype):
    return np.complex64 if dtype == _lib.F32 else np.complex128


def real_np(dtype):
    return np.float32 if dtype == _lib.F32 else np.float64


def dtype_of_samples(x, default=_lib.F64):
    """complex64/float32 inputs compute in f32; everything else in the reference's complex128."""
    dt = x.dtype
    name = str(dt).replace('torch.', '')
    if name in ('complex64', 'float32'):
        return _lib.F32
    if name in ('complex128', 'float64'):
        return _lib.F64
    return default


class ModemTables:
    """Per-modulator cache of the constellation table on the device, one per dtype."""

    def __init__(self):
        self._cache = {}

    def get(self, kind, symbols, dtype):
        torch = _lib.torch_cuda()
        key = (dtype, torch.cuda.current_device())
        ent = self._cache.get(key)
        if ent is None:
            tab = np.ascontiguousarray(np.asarray(symbols, dtype=complex_np(dtype)))
            ent = torch.from_numpy(tab).cuda()
            self._cache[key] = ent
        m = _lib.Modem(kind, int(np.asarray(symbols).size), ent.data_ptr())
        return m, ent

    def clear(self):
        self._cache = {}
