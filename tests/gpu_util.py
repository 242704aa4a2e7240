"""Helpers for the -m gpu tests: torch owns device memory, the product is called through the C-ABI."""
import ctypes as C
import numpy as np
import torch

from zetaray_b200 import lib, check


def dev(arr):
    """numpy -> CUDA tensor (bytes preserved)."""
    a = np.ascontiguousarray(arr)
    t = torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).cuda()
    return t


def dptr(t):
    return C.c_void_p(t.data_ptr())


def host(t, dtype, shape=None):
    a = t.detach().cpu().numpy().view(dtype)
    return a.reshape(shape) if shape is not None else a


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
