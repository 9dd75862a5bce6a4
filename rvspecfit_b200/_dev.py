"""Device-memory plumbing: PyTorch owns HBM buffers and streams, nothing else.
Every compute step is a C-ABI call into librvs_b200.so (see _cabi.py)."""
import ctypes

import numpy as np

from . import _cabi


# bytes moved between host and device by this process: [host->device, device->host]
# (bench.py reports them per step for the end-to-end figure)
IO_BYTES = [0, 0]


def torch_mod():
    return _cabi.require_cuda()


def device():
    torch = torch_mod()
    return torch.device('cuda', torch.cuda.current_device())


def upload(a, dtype):
    """numpy -> device tensor (contiguous, given numpy dtype)."""
    torch = torch_mod()
    a = np.ascontiguousarray(a, dtype=dtype)
    IO_BYTES[0] += a.nbytes
    return torch.from_numpy(a).to(device(), non_blocking=False)


def upload_concat(arrays, dtype):
    """Concatenate host arrays straight into pinned staging memory and start the
    copy to the device: (host view, device tensor).  One pass over the data on the
    host and a DMA at full PCIe rate instead of a pageable copy."""
    torch = torch_mod()
    tdt = {np.float64: torch.float64, np.bool_: torch.bool, np.int64: torch.int64}[dtype]
    n = int(sum(len(a) for a in arrays))
    pin = torch.empty((n,), dtype=tdt, pin_memory=True)
    host = pin.numpy()
    if n:
        np.concatenate(arrays, out=host)
    IO_BYTES[0] += host.nbytes
    return host, pin.to(device(), non_blocking=True)


def empty(shape, dtype):
    torch = torch_mod()
    tdt = {np.float64: torch.float64, np.float32: torch.float32, np.int32: torch.int32,
           np.int64: torch.int64}[dtype]
    return torch.empty(shape, dtype=tdt, device=device())


def zeros(shape, dtype):
    t = empty(shape, dtype)
    t.zero_()
    return t


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def hptr(a):
    """Raw host pointer of a numpy array."""
    return ctypes.c_void_p(a.ctypes.data)


def stream():
    torch = torch_mod()
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def download(t):
    IO_BYTES[1] += t.numel() * t.element_size()
    return t.detach().cpu().numpy()
