"""Development aid: runs a few device-resident evaluations of config 2 with
the cuda_options given as JSON in OPTY_OPTS (for ncu captures)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads
from opty_b200 import ConstraintCollocator
opts = json.loads(os.environ.get('OPTY_OPTS', '{}'))
opts.setdefault('out_ring', 4)
w = workloads.n_link_pendulum(10, 10000)
col = ConstraintCollocator(*w.collocator_args(), **w.collocator_kwargs(), cuda_options=opts)
col.generate_constraint_function()
h = col._evaluator.handle
h.upload_free(w.free(col.num_free))
n = int(os.environ.get('OPTY_REPS', 12))
print('ms per eval', h.time_device_evals(n) / n)
