#!/usr/bin/env python
"""FP64 tensor-pipe (DMMA) throughput of this box, measured with the dense LU's own trailing-update
kernel (MEASURED_PEAKS.json has no FP64 figure), next to torch.matmul in
float64 (cuBLAS) on the same shapes.  One JSON line per shape.

    python benchmarks/fp64_peak.py
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from nodal_b200 import _lib  # noqa: E402
from nodal_b200.device import Device  # noqa: E402

dev = Device.get(0)
p = dev.ptr
for M, N, K in ((8192, 8192, 8192), (16256, 16256, 128), (8192, 8192, 128)):
    ld = max(N, K)
    A = torch.rand(M, ld, dtype=torch.float64, device="cuda")
    B = torch.rand(max(K, 1), ld, dtype=torch.float64, device="cuda")
    Cm = torch.zeros(M, ld, dtype=torch.float64, device="cuda")
    ms = C.c_double(0.0)
    reps = 3 if K > 1000 else 20
    _lib.check(dev.lib.nodal_dgemm_sub_profile(dev.ctx, p(Cm), p(A), p(B), M, N, K, ld, reps, C.byref(ms), dev.stream()),
               "nodal_dgemm_sub_profile")
    flops = 2.0 * M * N * K
    a, b = A[:, :K].contiguous(), B[:K, :N].contiguous()
    torch.matmul(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b)
    e1.record()
    torch.cuda.synchronize()
    ms_blas = e0.elapsed_time(e1) / reps
    print(json.dumps({"shape": [M, N, K], "lu_gemm_kernel_ms": ms.value, "lu_gemm_kernel_tflops": flops / ms.value / 1e9,
                      "cublas_dgemm_ms": ms_blas, "cublas_dgemm_tflops": flops / ms_blas / 1e9}), flush=True)
