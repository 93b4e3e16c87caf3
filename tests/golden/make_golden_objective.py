"""Golden values for the objective / gradient path: the REFERENCE's
``opty.utils.create_objective_function`` (opty/utils.py:329-470) run on seeded
inputs.  Build container only:

    python tests/golden/make_golden_objective.py
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'baseline', 'stubs'))
sys.path.insert(0, '/root/reference')

import cases  # noqa: E402
from opty.utils import create_objective_function  # noqa: E402

if __name__ == '__main__':
    out = {}
    for name, make in cases.objective_cases().items():
        c = make()
        obj, grad = create_objective_function(
            c['objective'], c['states'], c['inputs'], c['params'], c['N'],
            c['h'], integration_method=c['method'], time_symbol=c['t'])
        out[name + '_value'] = np.float64(obj(c['free']))
        g = np.asarray(grad(c['free']), dtype=float)
        if g.size > 5000:
            # large cases: every 9th entry plus the parameter tail
            idx = np.unique(np.concatenate((np.arange(0, g.size, 9),
                                            np.arange(g.size - 4, g.size))))
            out[name + '_grad_index'] = idx
            g = g[idx]
        out[name + '_grad'] = g
        print(name, out[name + '_value'], g.shape)
    np.savez_compressed(os.path.join(HERE, 'objective_cases.npz'), **out)
