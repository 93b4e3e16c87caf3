# compute-sanitizer runs of the hot path: memcheck and racecheck
# on BASELINE config 1 (backward Euler, odd K: warp copy-out path), on the
# small config-2 smoke and a three-tile problem with free parameters (both the
# row-stationary kernel: TMA input windows, tile stores, bulk copies) and on a
# grid-kernel problem with instance constraints.  Summaries go to gpurun_out/.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
cat > /tmp/sanitize_case.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np
import workloads
from opty_b200 import ConstraintCollocator
for make in (lambda: workloads.pendulum_swing_up(51),
             lambda: workloads.n_link_pendulum(10, 40, seed=7),
             lambda: workloads.n_link_pendulum_torques(5, 700),
             lambda: workloads.n_link_pendulum_periodic(4, 200)):
    w = make()
    col = ConstraintCollocator(*w.collocator_args(), **w.collocator_kwargs(), device=0)
    free = w.free(col.num_free)
    con = col.generate_constraint_function()(free)
    jac = np.array(col.generate_jacobian_function()(free))
    rows, cols = col.jacobian_indices()
    print(w.name, len(con), len(jac), float(np.abs(jac).sum()))
    col.close()
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/sanitize_case.py > gpurun_out/r03_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|pendulum|vyasa" gpurun_out/r03_sanitizer_$tool.log | tail -6
done
