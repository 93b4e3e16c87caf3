"""Build container only: evaluates a few rows of the 50-link discrete EOM
(BASELINE config 5) at constraint node 0 with SymPy's arbitrary-precision
``evalf`` -- independent of the tape / emitter -- and stores them as a golden
fixture for tools/config5.py / tests."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import sympy as sm
import workloads
from opty_b200 import ConstraintCollocator

N = 50000
t0 = time.time()
w = workloads.n_link_pendulum(50, N)
col = ConstraintCollocator(*w.collocator_args(), **w.collocator_kwargs())
print('setup', time.time() - t0, flush=True)
free = w.free(col.num_free)
n, q = col.num_states, col.num_unknown_input_trajectories
subs = {}
for j, (si, sn) in enumerate(zip(col.current_discrete_state_symbols, col.next_discrete_state_symbols)):
    subs[si] = sm.Float(free[j * N + 0], 30); subs[sn] = sm.Float(free[j * N + 1], 30)
for j, (ui, un) in enumerate(zip(col.current_unknown_discrete_specified_symbols, col.next_unknown_discrete_specified_symbols)):
    subs[ui] = sm.Float(free[(n + j) * N + 0], 30); subs[un] = sm.Float(free[(n + j) * N + 1], 30)
for p, v in col.known_parameter_map.items():
    subs[p] = sm.Float(v, 30)
subs[col.time_interval_symbol] = sm.Float(col.node_time_interval, 30)
rows = [0, 25, 51, 52, 60, 75, 90, 101]
vals = []
for r in rows:
    t0 = time.time()
    v = col.discrete_eom[r].xreplace(subs).evalf(30)
    vals.append(float(v)); print(r, float(v), time.time() - t0, flush=True)
np.savez(os.path.join(ROOT, 'tests', 'golden', 'cfg5_pendulum50_node0_rows.npz'), rows=np.array(rows), values=np.array(vals), free_head=free[:8])
