"""Development aid: register-pressure proxy of the emitted equation bodies --
peak number of simultaneously live temporaries in the emission order of
codegen._BodyWriter -- for the n-link pendulum (python tools/liveness.py LINKS)."""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads
from opty_b200 import ConstraintCollocator, ir
from opty_b200.codegen import _BodyWriter
links = int(sys.argv[1])
w = workloads.n_link_pendulum(links, 2000)
col = ConstraintCollocator(*w.collocator_args(), **w.collocator_kwargs(), cuda_options={'use_index': False})
prog = col._build_program()
T = prog.tape
P = prog.P
for j in (prog.M // 2 + 1, prog.M - 1):
    # emission order of one equation's body (residual + P partials), as codegen does it
    bw = _BodyWriter(prog, 'main', {})
    order = []
    emit = bw._emit
    def rec(i, emit=emit, order=order):
        order.append(i); emit(i)
    bw._emit = rec
    n = prog.M          # states (square systems: n = M)
    cols = list(range(P))
    if len(sys.argv) > 2 and sys.argv[2] == 'bylink':
        # partials with respect to the same state (current / next value) together
        half = n // 2
        cols = []
        for k in range(half):
            cols += [k, n + k, half + k, n + half + k]
        cols += [c for c in range(P) if c not in cols]
    outs = [prog.con[j]] + [prog.jac[j][c] for c in cols]
    for o in outs:
        bw.need(o)
    if len(sys.argv) > 2 and sys.argv[2] == 'tapeorder':
        # operation-major: nodes in creation order of the tape (forward mode
        # creates the tangent nodes of an operation right after it)
        order = sorted(order)
    pos = {nid: k for k, nid in enumerate(order)}
    last = {}
    for k, nid in enumerate(order):
        for o in T.operands(nid):
            if o in pos:
                last[o] = k
    for o in outs:
        if o in pos:
            last[o] = max(last.get(o, 0), pos[o])
    # live count over time
    delta = [0] * (len(order) + 2)
    for nid, k in pos.items():
        e = last.get(nid, k)
        delta[k] += 1; delta[e + 1] -= 1
    live = 0; peak = 0
    for d in delta:
        live += d; peak = max(peak, live)
    long_lived = sum(1 for nid, k in pos.items() if last.get(nid, k) - k > 2000)
    inputs_used = len({o for nid in order for o in T.operands(nid) if T.op[o] == ir.VIN})
    print('links %d row %d: %d ops, peak live temporaries %d, temporaries living > 2000 ops: %d, distinct inputs read %d' % (
        links, j, len(order), peak, long_lived, inputs_used), flush=True)
