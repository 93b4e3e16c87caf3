"""Development aid: where the end-to-end callback time goes (config 2)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import workloads
from opty_b200 import ConstraintCollocator

w = workloads.n_link_pendulum(10, 10000)
col = ConstraintCollocator(*w.collocator_args(), **w.collocator_kwargs())
con_f = col.generate_constraint_function(); jac_f = col.generate_jacobian_function()
h = col._evaluator.handle
free = w.free(col.num_free)
frees = [free + 1e-3 * i for i in range(4)]
def t(fn, n=50):
    fn(0)
    t0 = time.perf_counter()
    for i in range(n): fn(i)
    return 1e3 * (time.perf_counter() - t0) / n
print('upload_free (changed)       %.3f ms' % t(lambda i: h.upload_free(frees[i % 4])))
print('upload_free (unchanged)     %.3f ms' % t(lambda i: h.upload_free(frees[0])))
print('eval_device sync            %.3f ms' % t(lambda i: h.eval_device(sync=True)))
print('handle.constraints changed  %.3f ms' % t(lambda i: h.constraints(frees[i % 4])))
print('handle.jacobian same free   %.3f ms' % t(lambda i: (h.constraints(frees[i % 4]), h.jacobian(frees[i % 4]))[1]), '(includes the constraints call above)')
print('python con_f changed        %.3f ms' % t(lambda i: con_f(frees[i % 4])))
print('python con_f + jac_f        %.3f ms' % t(lambda i: (con_f(frees[i % 4]), jac_f(frees[i % 4]))))
print('d2h ranges', col._evaluator.d2h_ranges)
import ctypes
# raw D2H rates via torch for comparison
import torch
x = torch.empty(41_915_808 // 8, dtype=torch.float64, device='cuda')
y = torch.empty_like(x, device='cpu').pin_memory()
torch.cuda.synchronize()
def cp(i):
    y.copy_(x, non_blocking=True); torch.cuda.synchronize()
print('torch contiguous D2H 41.9MB %.3f ms' % t(cp))
