import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import workloads
from opty_b200 import ConstraintCollocator
w = workloads.n_link_pendulum(10, 40, seed=7)
one = ConstraintCollocator(*w.collocator_args(), **w.collocator_kwargs())
many = ConstraintCollocator(*w.collocator_args(), **w.collocator_kwargs(), devices=[0, 0, 0])
free = w.free(one.num_free)
jf, j1 = many.generate_jacobian_function(), one.generate_jacobian_function()
cf, c1 = many.generate_constraint_function(), one.generate_constraint_function()
for k, point in enumerate((free, free * 1.01, free * 1.01, free)):
    c_a, c_b = cf(point), c1(point)
    a, b = np.array(jf(point)), np.array(j1(point))
    bad = np.nonzero(a != b)[0]
    print('point', k, 'con equal', np.array_equal(c_a, c_b), 'jac mismatches', len(bad))
    if len(bad):
        K = 22 * 46
        for e in bad[:12]:
            print('   entry', e, 'node', e // K, 'row', (e % K) // 46, 'col', e % 46, a[e], b[e])
        print('   nodes', sorted(set(bad // K))[:20], 'rows', sorted(set((bad % K) // 46)))
