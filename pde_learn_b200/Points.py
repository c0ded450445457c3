"""
Generate_Points(Bounds, Num_Points, Device) -> [Num_Points, d] float32, uniform in prod_j [a_j, b_j].
Same signature as the reference (Code/Points.py:7-59), sampled on the device with a counter-based
Philox4x32-10 generator instead of a Python double loop over random.uniform.

Extra keywords: `Seed` (default: a process-wide counter, so successive calls differ like the reference's
unseeded generator) and `First_Index` (global row index of the first sample: a rank that owns rows
[a, b) of a batch passes First_Index=a and gets exactly those rows, for any number of ranks).
"""
import ctypes as C
import itertools

import numpy
import torch

from . import _lib as L

_CALLS = itertools.count(1)


def Generate_Points(Bounds: numpy.ndarray, Num_Points: int, Device: torch.device = None, Seed: int = None,
                    First_Index: int = 0) -> torch.Tensor:
    B = numpy.asarray(Bounds, dtype=numpy.float64)
    assert B.ndim == 2 and B.shape[1] == 2
    d = int(B.shape[0])
    for j in range(d):
        assert B[j, 0] <= B[j, 1]                               # Points.py:41-42
    dev = torch.device(Device) if Device is not None else torch.device("cuda")
    if dev.type != "cuda":
        raise RuntimeError("Generate_Points samples on the GPU; Device must be a CUDA device (no CPU path)")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    out = torch.empty((int(Num_Points), d), dtype=torch.float32, device=dev)
    seed = int(Seed) if Seed is not None else (0x9E3779B97F4A7C15 * next(_CALLS)) & 0xFFFFFFFFFFFFFFFF
    flat = numpy.ascontiguousarray(B.reshape(-1))
    with torch.cuda.device(dev):
        L.check(L.lib().pdl_sample_points(flat.ctypes.data_as(C.POINTER(C.c_double)), d, int(Num_Points),
                                          C.c_uint64(seed), C.c_uint64(int(First_Index)), out.data_ptr(),
                                          torch.cuda.current_stream(dev).cuda_stream))
    return out
