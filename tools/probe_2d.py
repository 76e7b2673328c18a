#!/usr/bin/env python
"""How fast is a STRIDED host -> device copy (cudaMemcpy2DAsync from pinned memory, rows of `width` bytes at a 512-byte pitch)?
Decides whether the GP host call should send only the column prefixes the kernels read (upper triangle) instead of whole
128x128 matrices.  Prints useful GB/s (width * rows / time) per width; width 512 is the contiguous reference."""
import ctypes
import sys

import torch

rt = ctypes.CDLL("libcudart.so")
rt.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t,
                                 ctypes.c_int, ctypes.c_void_p]
total = 1 << 30
h = torch.empty(total, dtype=torch.uint8).pin_memory()
d = torch.empty(total, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
pitch = 512
rows = total // pitch
for width in (512, 384, 256, 128, 64):
    ts = []
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = rt.cudaMemcpy2DAsync(d.data_ptr(), pitch, h.data_ptr(), pitch, width, rows, 1, st)
        assert rc == 0, rc
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts[1:])
    print(f"width {width:4d} B of pitch {pitch}: {ms:8.2f} ms  useful {width * rows / ms / 1e6:7.2f} GB/s  (time vs contiguous {ms / (total / 55e9 * 1e3):.2f}x of a 55 GB/s copy)")
