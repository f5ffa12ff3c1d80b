"""torch.Tensor <-> omx_array descriptors.  Arrays on this path ARE torch CUDA tensors:
PyTorch supplies device memory and streams (plumbing), never arithmetic."""
import ctypes

import torch

from . import _lib

_DT = {torch.float32: _lib.OMX_FLOAT32, torch.float16: _lib.OMX_FLOAT16, torch.bfloat16: _lib.OMX_BFLOAT16,
       torch.bool: _lib.OMX_BOOL, torch.int32: _lib.OMX_INT32}
_TD = {v: k for k, v in _DT.items()}
_CAI = {_lib.OMX_FLOAT32: ("<f4", torch.float32), _lib.OMX_FLOAT16: ("<f2", torch.float16),
        _lib.OMX_BFLOAT16: ("<u2", torch.bfloat16)}


def desc(t):
    """Borrowed descriptor of a CUDA tensor (no copy; strides preserved)."""
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        raise _lib.Exception_(f"expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise _lib.Exception_("arrays on the B200 attention path must live in device memory (got a CPU tensor); "
                              "there is no CPU fallback")
    if t.dtype not in _DT:
        raise _lib.Exception_(f"unsupported dtype {t.dtype}")
    if t.dim() > _lib.OMX_MAX_NDIM:
        raise _lib.Exception_("too many dimensions")
    a = _lib.OmxArray()
    a.data = t.data_ptr()
    a.dtype = _DT[t.dtype]
    a.ndim = t.dim()
    for i in range(t.dim()):
        a.shape[i] = t.shape[i]
        a.strides[i] = t.stride(i)
    return a


def ref(a):
    return None if a is None else ctypes.byref(a)


class _Blob:
    """Exposes library-owned device memory through __cuda_array_interface__ and keeps `owner`
    (the cache wrapper) alive for as long as a tensor view exists."""

    def __init__(self, a, owner):
        typestr, _ = _CAI[a.dtype]
        item = 4 if a.dtype == _lib.OMX_FLOAT32 else 2
        self.owner = owner
        self.__cuda_array_interface__ = {
            "shape": tuple(int(a.shape[i]) for i in range(a.ndim)),
            "strides": tuple(int(a.strides[i]) * item for i in range(a.ndim)),
            "typestr": typestr,
            "data": (int(a.data), False),
            "version": 3,
        }


def view(a, owner, device):
    """omx_array (library-owned memory) -> torch tensor view, zero-copy."""
    shape = tuple(int(a.shape[i]) for i in range(a.ndim))
    _, tdt = _CAI[a.dtype]
    if 0 in shape or not a.data:
        return torch.empty(shape, dtype=tdt, device=device)
    t = torch.as_tensor(_Blob(a, owner), device=device)
    return t.view(torch.bfloat16) if a.dtype == _lib.OMX_BFLOAT16 else t


def stream_ptr(stream=None):
    """`#[default_device]` (mlx-internal-macros/src/lib.rs:83-94): default to the current stream."""
    if stream is None:
        stream = torch.cuda.current_stream()
    return ctypes.c_void_p(stream.cuda_stream)
