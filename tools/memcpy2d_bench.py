"""Development aid: device->host rate of 2-D copies (the column-range Jacobian
fetch of opty_colloc_set_d2h_columns) as a function of the column-range width.
Source: [rows][pitch] float64 on the device, destination pinned, same pitch."""
import time

import numpy as np
import torch
from cuda.bindings import runtime as rt

rows, pitch_cols = 20000, 846
src = torch.zeros(rows * pitch_cols, dtype=torch.float64, device='cuda')
dst = torch.zeros(rows * pitch_cols, dtype=torch.float64).pin_memory()
err, stream = rt.cudaStreamCreate()
pitch = pitch_cols * 8
for width_cols in (1, 2, 4, 8, 16, 32, 64, 128, 256, 423, 846):
    for nranges in (1, 4):
        w = width_cols * 8
        if nranges * width_cols > pitch_cols:
            continue
        def go():
            for r in range(nranges):
                off = r * (pitch_cols // nranges) * 8
                rt.cudaMemcpy2DAsync(dst.data_ptr() + off, pitch,
                                     src.data_ptr() + off, pitch, w, rows,
                                     rt.cudaMemcpyKind.cudaMemcpyDeviceToHost,
                                     stream)
            rt.cudaStreamSynchronize(stream)
        go()
        t0 = time.perf_counter()
        n = 10
        for _ in range(n):
            go()
        dt = (time.perf_counter() - t0) / n
        mb = nranges * w * rows / 1e6
        print('width %4d cols x %d ranges: %7.3f ms  %7.2f MB  %6.2f GB/s' % (
            width_cols, nranges, 1e3 * dt, mb, mb / dt / 1e3), flush=True)
