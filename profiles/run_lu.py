"""Short driver for ncu: one dense LU solve (config C3 family) at n = 8192."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from nodal_b200.device import Device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = Device.get(0)
rng = np.random.default_rng(0)
A = rng.standard_normal((n, n))
b = rng.standard_normal(n)
G = dev.to_device(A)
rhs = dev.to_device(b)
for _ in range(2):
    x, info = dev.lu_solve(G.clone(), rhs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
work = G.clone()
e0.record()
x, info = dev.lu_solve(work, rhs)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
xs = x.cpu().numpy()
print(f"n={n} lu_solve {ms:.2f} ms  {(2/3*n**3 + 2*n*n)/ms/1e9:.2f} TFLOP/s  status={info['status']} "
      f"backward_err={np.linalg.norm(A @ xs - b) / (np.linalg.norm(A) * np.linalg.norm(xs)):.2e}")
