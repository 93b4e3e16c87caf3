"""Ad-hoc first GPU check (development aid, not part of the test-suite)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import workloads
from opty_b200 import ConstraintCollocator

def compare(a, b, K=None, name=''):
    d = np.abs(a - b)
    rel = d / np.maximum(np.abs(b), 1e-300)
    big = np.abs(b) > 1e-8
    print(name, 'max abs', d.max(), 'max rel(|b|>1e-8)', rel[big].max() if big.any() else 0,
          'exact frac', (a == b).mean(), 'zero pattern equal', np.array_equal(a == 0, b == 0), flush=True)

opts = json.loads(os.environ.get('OPTY_OPTS', '{}'))
N = int(os.environ.get('OPTY_N', 10000))
t0 = time.time()
w = workloads.n_link_pendulum(10, N)
col = ConstraintCollocator(*w.collocator_args(), **w.collocator_kwargs(), cuda_options=opts, show_compile_output=bool(os.environ.get('OPTY_V')))
con_f = col.generate_constraint_function(); jac_f = col.generate_jacobian_function()
print('setup', time.time() - t0, 'groups', col._evaluator.parts, 'cache hit', col._evaluator.cache_hit, flush=True)
free = w.free(col.num_free)
t0 = time.time(); con = con_f(free); t1 = time.time(); jac = jac_f(free); t2 = time.time()
print('first call con %.3f ms jac %.3f ms' % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), 'kernel ms', col._evaluator.handle.last_kernel_ms())
if N == 10000:
    ref = np.load(os.path.join(ROOT, 'oracle/_ref/cfg2_probe.npz'))
    assert np.array_equal(ref['free'], free)
    compare(con, ref['con'], name='con vs reference')
    compare(np.array(jac), ref['jac'], name='jac vs reference')
else:
    from host_harness import host_evaluate
    c2, j2 = host_evaluate(col, free)
    compare(con, c2, name='con vs host harness'); compare(np.array(jac), j2, name='jac vs host harness')
h = col._evaluator.handle
ms = []
for i in range(30):
    h.eval_device(sync=True); ms.append(h.last_kernel_ms())
ms = np.array(ms[5:])
B = 8 * (col.num_free + col.num_eom * (N - 1) + (N - 1) * col.num_eom * col._evaluator.program.P)
print('kernel ms median %.4f min %.4f ; algorithmic bytes %.3f MB ; %.1f GB/s (median)' % (np.median(ms), ms.min(), B / 1e6, B / np.median(ms) / 1e6))
# e2e timing
for rep in range(3):
    f2 = free.copy(); f2[0] += 1e-3 * (rep + 1)
    t0 = time.perf_counter(); c = con_f(f2); t1 = time.perf_counter(); j = jac_f(f2); t2 = time.perf_counter()
    print('e2e con %.3f ms jac %.3f ms -> %.1f evals/s' % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, 1.0 / (t2 - t0)))
rows, cols = col.jacobian_indices()
print('indices', rows.shape, rows[:5], cols[:5], rows.dtype)
