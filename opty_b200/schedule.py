"""Register-pressure-aware ordering of one straight-line kernel body.

A body evaluates a set of tape outputs (residuals and Jacobian partials of a
few equations of motion) for one collocation node per lane.  The generated
code is straight-line float64 arithmetic that every warp walks exactly once,
so what bounds it on the SM is (i) how many warps fit -- i.e. registers per
thread -- and (ii) for large models whether the live values fit in the
register file at all: emitted output by output (residual first, then the
partials in column order, temporaries depth-first) the heaviest equation of
the 50-link chain keeps 1 301 float64 values alive, ptxas spills 3.2 KB per
thread and the kernel moves 7x its algorithmic bytes through DRAM
(profiles/sweeps_r01.md).  The reference has the same structure on the CPU
(one C function with thousands of ``z_k`` temporaries, opty/utils.py:483-494)
where the 1 KB-per-thread problem does not exist.

This module orders the operations instead of the outputs:

* **target-driven depth-first evaluation**: the next output is always the one
  that needs the fewest operations that are not yet computed; its expression
  tree is evaluated heavier-subtree-first (Sethi-Ullman order);
* **eager consumers**: any operation all of whose operands are available and
  that is the last use of at least one of them is issued immediately -- it
  cannot increase the number of live values;
* **sums accumulate in arrival order** (``reassociate=True``): every maximal
  tree of additions / subtractions whose interior nodes have a single use is
  treated as one n-ary sum with ONE accumulator; a term is folded into the
  accumulator as soon as it exists instead of waiting for its neighbour in the
  printed association order.  This is what lets the big sums of a multibody
  equation (the residual itself, the partial with respect to the equation's
  own coordinate) be interleaved with the partials that share their terms;
* **rematerialisation**: values that cost at most ``inline_cost`` operations
  from the always-available leaves are never kept in a register; values that
  cost at most ``remat_cost`` are recomputed at a use that comes more than
  ``remat_gap`` events after the previous one.  Loads of the node's inputs
  (shared-memory slice of the trajectory matrix, derived rows) are values like
  any other: reloaded after a long gap instead of pinning a register.

With ``reassociate=False`` the operations are exactly the tape's operations
(association order of the reference's C printer, lowering.py), only their
order changes, so results are bit-identical to the plain emission order.

The peak number of simultaneously live values falls from 126 to ~10-25 for
the heaviest 10-link equation, from 296 to ~25-50 at 20 links and from 1 301
to ~90 at 50 links (tools/liveness.py), for 10-15 % more operations.
"""

from . import ir

_SUM_OPS = (ir.ADD, ir.SUB, ir.NEG)

# event kinds
LOAD = 'load'      # ('load', leaf id)               register copy of an input
OP = 'op'          # ('op', node id)                 ordinary operation
ACC = 'acc'        # ('acc', sum id, term index, first)  fold one term into the sum
OUT = 'out'        # ('out', output index)


class BodyDag(object):
    """Scheduling view of one body: register values (loads, operations, sums)
    and the register values each of them reads, after flattening of sums and
    inlining of cheap values."""

    def __init__(self, tape, outputs, stop=frozenset(), reassociate=True,
                 inline_cost=3, inline_edges=frozenset()):
        T = self.tape = tape
        # one output slot may hold several values (two adjacent Jacobian
        # columns are stored with one 16-byte shared-memory store)
        self.outputs = [tuple(o) if isinstance(o, (tuple, list)) else (o,)
                        for o in outputs]
        outputs = [r for slot in self.outputs for r in slot]
        self.stop = stop
        op, varying = T.op, T.varying
        a_, b_, c_ = T.a, T.b, T.c
        CONST, VIN, UIN = ir.CONST, ir.VIN, ir.UIN

        def kind_of(i):
            """0 free leaf (literal / node-invariant), 1 load, 2 operation"""
            o = op[i]
            if o == VIN or i in stop:
                return 1
            if o <= UIN or not varying[i]:
                return 0
            return 2
        self.kind_of = kind_of

        def operands(i):
            if c_[i] >= 0:
                return (a_[i], b_[i], c_[i])
            if b_[i] >= 0:
                return (a_[i], b_[i])
            return (a_[i],)
        self.operands = operands

        # reachable operations and loads
        ops = set()
        loads = set()
        stack = list(outputs)
        while stack:
            i = stack.pop()
            k = kind_of(i)
            if k == 0:
                continue
            if k == 1:
                loads.add(i)
                continue
            if i in ops:
                continue
            ops.add(i)
            stack.extend(operands(i))
        nodes = sorted(ops)
        uses = {}
        single_consumer = {}
        for v in nodes:
            for o in operands(v):
                if o in ops:
                    uses[o] = uses.get(o, 0) + 1
                    single_consumer[o] = v
        out_set = set(outputs)
        for o in outputs:
            if o in ops:
                uses[o] = uses.get(o, 0) + 1

        # ---- sums ------------------------------------------------------
        interior = set()
        if reassociate:
            for v in nodes:
                if op[v] in _SUM_OPS and uses.get(v, 0) == 1 and \
                        v not in out_set and \
                        op[single_consumer[v]] in _SUM_OPS:
                    interior.add(v)
        terms = {}
        if reassociate:
            for v in nodes:
                if v in interior or op[v] not in (ir.ADD, ir.SUB):
                    continue
                tl = []
                stack = [(1, v)]
                while stack:
                    sg, x = stack.pop()
                    if (x == v or x in interior) and op[x] in _SUM_OPS:
                        if op[x] == ir.NEG:
                            stack.append((-sg, a_[x]))
                        else:
                            stack.append((sg if op[x] == ir.ADD else -sg,
                                          b_[x]))
                            stack.append((sg, a_[x]))
                    else:
                        tl.append((sg, x))
                terms[v] = tl
        self.terms = terms
        self.interior = interior
        values = [v for v in nodes if v not in interior]

        def children(v):
            if v in terms:
                return [t for _, t in terms[v]]
            return operands(v)
        self.children = children

        # ---- cost of recomputing a value from the leaves -----------------
        # (distinct operations + loads in its cone, weighted with the
        # emitter's issue-slot estimate: a division or a sine is not "one
        # operation" when it comes to recomputing it)
        ordered = sorted(loads) + nodes
        pos = {v: k for k, v in enumerate(ordered)}
        heavy_mask = 0
        extra = {}
        for v in nodes:
            c = ir.OP_COST.get(op[v], 1)
            if c > 1:
                heavy_mask |= 1 << pos[v]
                extra[pos[v]] = int(c) - 1
        cone = {}
        for v in sorted(loads):
            cone[v] = 1 << pos[v]
        self.leaf_cost = {}
        for v in nodes:
            m = 1 << pos[v]
            for o in operands(v):
                if o in cone:
                    m |= cone[o]
            cone[v] = m
        for v, m in cone.items():
            cost = m.bit_count()
            hv = m & heavy_mask
            while hv:
                low = hv & -hv
                cost += extra[low.bit_length() - 1]
                hv ^= low
            self.leaf_cost[v] = cost
        del cone

        cheap = set(v for v in values if self.leaf_cost[v] <= inline_cost)
        self.cheap = cheap
        self.inline_edges = inline_edges
        reg = sorted(loads) + [v for v in values if v not in cheap]
        self.reg = reg
        regset = self.regset = set(reg)
        self.loads = loads

        def reg_operands(consumer, roots):
            """Register values read when ``consumer`` evaluates ``roots``,
            looking through cheap values and inlined edges (one entry per
            textual occurrence)."""
            out = []
            stack = list(roots)
            while stack:
                x = stack.pop()
                if x in regset:
                    if (consumer, x) in inline_edges:
                        if x not in loads:
                            stack.extend(children(x))
                        # an inlined load reads nothing
                    else:
                        out.append(x)
                elif x in cheap or x in interior:
                    stack.extend(children(x))
            return out
        self.rops = {}
        self.term_rops = {}     # sum -> per term: register values it reads
        for v in reg:
            if v in loads:
                self.rops[v] = []
            elif v in terms:
                per_term = [reg_operands(v, [t]) for _, t in terms[v]]
                self.term_rops[v] = per_term
                self.rops[v] = [o for tr in per_term for o in tr]
            else:
                self.rops[v] = reg_operands(v, children(v))
        self.out_rops = [reg_operands(('out', k), list(slot))
                         for k, slot in enumerate(self.outputs)]
        # register values nobody reads any more (every use recomputes them
        # inline) are not computed at all
        count = dict.fromkeys(reg, 0)
        for v in reg:
            for o in self.rops[v]:
                count[o] += 1
        for ro in self.out_rops:
            for o in ro:
                count[o] += 1
        dead = [v for v in reg if count[v] == 0]
        removed = set()
        while dead:
            v = dead.pop()
            removed.add(v)
            for o in self.rops[v]:
                count[o] -= 1
                if count[o] == 0:
                    dead.append(o)
        if removed:
            self.reg = reg = [v for v in reg if v not in removed]
            regset -= removed
            for v in removed:
                del self.rops[v]
                self.term_rops.pop(v, None)
            self.loads = loads = loads - removed

    # -- operation count of the code that will be emitted -------------------
    def inline_cost_of(self, consumer, x, memo=None):
        """Operations + loads spent on evaluating ``x`` inline inside
        ``consumer``."""
        if self.kind_of(x) == 0:
            return 0
        if x in self.regset and (consumer, x) not in self.inline_edges:
            return 0
        if self.kind_of(x) == 1:
            return 1
        kids = self.children(x)
        own = max(len(kids) - 1, 1) if x in self.terms else 1
        return own + sum(self.inline_cost_of(consumer, k) for k in kids)

    def emitted_ops(self):
        total = 0
        for v in self.reg:
            if v in self.loads:
                total += 1
                continue
            kids = self.children(v)
            own = max(len(kids) - 1, 1) if v in self.terms else 1
            total += own + sum(self.inline_cost_of(v, k) for k in kids)
        for k, slot in enumerate(self.outputs):
            for o in slot:
                total += self.inline_cost_of(('out', k), o)
        return total


def _schedule_dag(dag, phases):
    reg = dag.reg
    rops = dag.rops
    out_rops = dag.out_rops
    loads = dag.loads
    is_sum = dag.terms
    n_out = len(dag.outputs)
    idx = {v: k for k, v in enumerate(reg)}

    term_rops = dag.term_rops
    consumers = {v: [] for v in reg}       # ordinary operations reading v
    term_consumers = {v: [] for v in reg}  # (sum, term index) reading v
    term_missing = {}
    for v in reg:
        if v in is_sum:
            tm = []
            for j, tr in enumerate(term_rops[v]):
                distinct = set(tr)
                tm.append(len(distinct))
                for o in distinct:
                    term_consumers[o].append((v, j))
            term_missing[v] = tm
        else:
            for o in set(rops[v]):
                consumers[o].append(v)
    out_of = {}
    for k, ro in enumerate(out_rops):
        for o in set(ro):
            out_of.setdefault(o, []).append(k)
    rem_use = dict.fromkeys(reg, 0)
    for v in reg:
        for o in rops[v]:
            rem_use[o] += 1
    for ro in out_rops:
        for o in ro:
            rem_use[o] += 1
    cone = {}
    for v in reg:
        m = 1 << idx[v]
        for o in rops[v]:
            m |= cone[o]
        cone[v] = m
    out_cone = []
    for ro in out_rops:
        m = 0
        for o in ro:
            m |= cone[o]
        out_cone.append(m)

    state = {'done_mask': 0, 'live': 0, 'peak': 0}
    computed = set()
    started = set()
    # operations: distinct operands not yet computed; sums: terms (that read
    # at least one register value) not yet folded
    missing = {}
    for v in reg:
        if v in is_sum:
            missing[v] = sum(1 for m in term_missing[v] if m > 0)
        else:
            missing[v] = len(set(rops[v]))
    out_missing = [len(set(ro)) for ro in out_rops]
    out_done = [False] * n_out
    out_open = [False] * n_out
    events = []
    eager = []

    def use(o, count=1):
        rem_use[o] -= count
        if rem_use[o] == 0:
            state['live'] -= 1

    def store(k):
        out_done[k] = True
        events.append((OUT, k))
        for o in out_rops[k]:
            use(o)

    def born():
        state['live'] += 1
        if state['live'] > state['peak']:
            state['peak'] = state['live']

    def finish(v):
        computed.add(v)
        state['done_mask'] |= 1 << idx[v]
        for k in out_of.get(v, ()):
            out_missing[k] -= 1
            if out_missing[k] == 0 and out_open[k] and not out_done[k]:
                store(k)
        for c in consumers[v]:
            missing[c] -= 1
            if missing[c] == 0:
                note(c)
        for s, j in term_consumers[v]:
            term_missing[s][j] -= 1
            if term_missing[s][j] == 0:
                accumulate(s, j)

    def accumulate(s, j):
        first = s not in started
        if first:
            started.add(s)
            born()
        events.append((ACC, s, j, first))
        for o in term_rops[s][j]:
            use(o)
        missing[s] -= 1
        if missing[s] == 0:
            if rem_use[s] == 0:
                state['live'] -= 1
            finish(s)

    def kills(v):
        seen = {}
        for o in rops[v]:
            seen[o] = seen.get(o, 0) + 1
        return sum(1 for o, c in seen.items() if rem_use[o] == c)

    def note(v):
        if v in computed or missing[v] != 0 or v in is_sum:
            return
        if kills(v) >= 1:
            eager.append(v)

    def do_op(v):
        events.append((LOAD if v in loads else OP, v))
        born()
        touched = set(rops[v])
        for o in rops[v]:
            use(o)
        if rem_use[v] == 0:
            state['live'] -= 1
        finish(v)
        for o in touched:
            if rem_use[o] > 0:
                for c in consumers[o]:
                    if c not in computed:
                        note(c)

    def drain():
        while eager:
            v = eager.pop()
            if v in computed or missing[v] != 0:
                continue
            if kills(v) >= 1:
                do_op(v)

    def compute(root):
        stack = [(root, False)]
        while stack:
            v, expanded = stack.pop()
            if v in computed:
                continue
            if expanded:
                if v in is_sum:
                    # a sum whose register terms are all there has been
                    # finished by its last accumulate; one without any
                    # register term is an ordinary value
                    if v not in computed:
                        do_op(v)
                else:
                    do_op(v)
                drain()
                continue
            stack.append((v, True))
            dm = state['done_mask']
            kids = [o for o in set(rops[v]) if o not in computed]
            kids.sort(key=lambda o: ((cone[o] & ~dm).bit_count(), -o))
            for o in kids:          # lightest pushed first, heaviest on top
                stack.append((o, False))

    if phases is None:
        phases = [list(range(n_out))]
    for ph in phases:
        for k in ph:
            out_open[k] = True
            if out_missing[k] == 0 and not out_done[k]:
                store(k)
        pending = [k for k in ph if not out_done[k]]
        while pending:
            dm = state['done_mask']
            best = min(pending,
                       key=lambda k: ((out_cone[k] & ~dm).bit_count(), k))
            roots = sorted(set(out_rops[best]),
                           key=lambda o: (-(cone[o] & ~dm).bit_count(), o))
            for o in roots:
                compute(o)
            drain()
            pending = [k for k in pending if not out_done[k]]
    assert state['live'] == 0 and all(out_done), 'scheduler left values live'
    return events, state['peak']


def _use_times(dag, events):
    born = {}
    times = {}
    for t, e in enumerate(events):
        kind = e[0]
        if kind in (OP, LOAD):
            v = e[1]
            born[v] = t
            for o in set(dag.rops[v]):
                times.setdefault(o, []).append((t, v))
        elif kind == ACC:
            born.setdefault(e[1], t)
            for o in set(dag.term_rops[e[1]][e[2]]):
                times.setdefault(o, []).append((t, e[1]))
        else:
            for o in set(dag.out_rops[e[1]]):
                times.setdefault(o, []).append((t, (OUT, e[1])))
    return born, times


class Schedule(object):
    """Result of :func:`schedule_body`: ``events`` in emission order, the
    :class:`BodyDag` they refer to (which values are registers, which are
    inlined where), the model's peak live count and the operation count."""

    def __init__(self, dag, events, peak, deferred=()):
        self.dag = dag
        self.events = events
        self.peak_live = peak
        self.num_ops = dag.emitted_ops()
        self.deferred = frozenset(deferred)   # outputs moved to the last phase


def _live_profile(dag, events):
    """Number of live register values after every event."""
    rem = dict.fromkeys(dag.reg, 0)
    for v in dag.reg:
        for o in dag.rops[v]:
            rem[o] += 1
    for ro in dag.out_rops:
        for o in ro:
            rem[o] += 1
    live = 0
    started = set()
    prof = []
    for e in events:
        kind = e[0]
        if kind in (OP, LOAD):
            v = e[1]
            live += 1
            peak_here = live
            reads = dag.rops[v]
            if rem[v] == 0:
                live -= 1
        elif kind == ACC:
            s = e[1]
            peak_here = live
            if s not in started:
                started.add(s)
                live += 1
                peak_here = live
            reads = dag.term_rops[s][e[2]]
        else:
            peak_here = live
            reads = dag.out_rops[e[1]]
        for o in reads:
            rem[o] -= 1
            if rem[o] == 0:
                live -= 1
        prof.append(peak_here)
    return prof


def schedule_body(tape, outputs, stop=frozenset(), phases=None,
                  reassociate=True, inline_cost=2, remat_cost=24,
                  live_budget=40, max_gap=4096, min_gap=24, max_passes=40,
                  deferrable=(), defer_fraction=0.2):
    """Orders the evaluation of ``outputs`` (tape ids, or tuples of tape ids
    for slots that hold several values).

    ``stop``: tape ids that are inputs of this body (derived rows written by
    the pre-pass kernel).  ``phases``: lists of output indices; the outputs of
    one phase are stored before those of the next (a phase is what fits in
    the shared-memory staging buffers at once); ``None`` = no constraint.

    Rematerialisation is driven by ``live_budget``: while the schedule keeps
    more values alive than that at some point, idle intervals (previous use
    .. next use) longer than ``gap`` events that span such a point and belong
    to values costing at most ``remat_cost`` operations are cut -- the later
    use recomputes the value inline; ``gap`` starts at ``max_gap`` and
    shrinks by 30 % whenever no such interval is left, down to ``min_gap``.
    A body that fits the budget is not touched, so small models pay nothing.

    ``deferrable``: output indices that may be moved to the LAST phase
    (the emitter then stores them straight to global memory instead of
    through their phase's staging tile).  An output of an early phase whose
    cone is more than ``defer_fraction`` of the whole body -- the partial of
    a multibody equation with respect to its own coordinate collects a term
    from every body of the chain -- would force all of those terms to be
    computed before the phase can be flushed and again, or kept alive, for
    the later phases; moved to the end it accumulates while the other
    partials are produced.  The moved indices are in ``Schedule.deferred``.
    """
    dag = BodyDag(tape, outputs, stop, reassociate, inline_cost)
    deferred = []
    if phases is not None and len(phases) > 1 and deferrable:
        last = set(phases[-1])
        total = max(1, len(dag.reg))
        memo = {}

        def cone_size(k):
            seen = set()
            stack = list(dag.out_rops[k])
            while stack:
                x = stack.pop()
                if x in seen:
                    continue
                seen.add(x)
                stack.extend(dag.rops.get(x, ()))
            return len(seen)
        for k in deferrable:
            if k not in last and cone_size(k) > defer_fraction * total:
                deferred.append(k)
        if deferred:
            moved = set(deferred)
            phases = [[k for k in ph if k not in moved] for ph in phases]
            phases[-1] = phases[-1] + deferred
        del memo
    events, peak = _schedule_dag(dag, phases)
    edges = set()
    gap = max_gap
    passes = 0
    while remat_cost > 0 and peak > live_budget and passes < max_passes:
        prof = _live_profile(dag, events)
        # prefix count of over-budget events: the interval (a, b] spans one
        # iff over[b + 1] - over[a + 1] > 0
        over = [0]
        for x in prof:
            over.append(over[-1] + (1 if x > live_budget else 0))
        born, times = _use_times(dag, events)
        found = set()
        for v, uses in times.items():
            if dag.leaf_cost.get(v, 1 << 30) > remat_cost:
                continue
            prev = born[v]
            for t, consumer in uses:
                if t - prev > gap:
                    if over[t + 1] - over[prev + 1] > 0:
                        found.add((consumer, v))
                else:
                    prev = t
        found -= edges
        if not found:
            if gap <= min_gap:
                break
            gap = max(min_gap, int(gap * 0.7))
            continue
        edges |= found
        passes += 1
        dag = BodyDag(tape, outputs, stop, reassociate, inline_cost,
                      frozenset(edges))
        events, peak = _schedule_dag(dag, phases)
    return Schedule(dag, events, peak, deferred)


def plain_order(tape, outputs, stop=frozenset()):
    """The emission order without scheduling (outputs in the given order,
    temporaries depth-first just before their first use, inputs read in
    place): what the emitter did before this module existed; kept for
    ``tools/liveness.py`` and as the ``schedule=False`` option."""
    dag = BodyDag(tape, outputs, stop, reassociate=False, inline_cost=0)
    # inputs are read in place: no load events, not register values
    dag.regset -= dag.loads
    dag.reg = [v for v in dag.reg if v not in dag.loads]
    for v in dag.reg:
        dag.rops[v] = [o for o in dag.rops[v] if o not in dag.loads]
    dag.out_rops = [[o for o in ro if o not in dag.loads]
                    for ro in dag.out_rops]
    dag.loads = set()
    done = set()
    events = []
    for k, slot in enumerate(dag.outputs):
        for root in slot:
            if root not in dag.regset or root in done:
                continue
            stack = [(root, False)]
            while stack:
                v, expanded = stack.pop()
                if v in done:
                    continue
                if expanded:
                    events.append((OP, v))
                    done.add(v)
                    continue
                stack.append((v, True))
                for o in reversed(dag.operands(v)):
                    if o in dag.regset and o not in done:
                        stack.append((o, False))
        events.append((OUT, k))
    return Schedule(dag, events, peak_live(dag, events))


def peak_live(dag, events):
    """Replays ``events`` and returns the peak number of live register
    values (consistency check of the scheduler's own count)."""
    rem = dict.fromkeys(dag.reg, 0)
    for v in dag.reg:
        for o in dag.rops[v]:
            rem[o] += 1
    for ro in dag.out_rops:
        for o in ro:
            rem[o] += 1
    live = peak = 0
    started = set()

    def use(o, c=1):
        nonlocal live
        rem[o] -= c
        if rem[o] == 0:
            live -= 1
    for e in events:
        kind = e[0]
        if kind in (OP, LOAD):
            v = e[1]
            live += 1
            peak = max(peak, live)
            for o in dag.rops[v]:
                use(o)
            if rem[v] == 0:
                live -= 1
        elif kind == ACC:
            s, j = e[1], e[2]
            if s not in started:
                started.add(s)
                live += 1
                peak = max(peak, live)
            for o in dag.term_rops[s][j]:
                use(o)
        else:
            for o in dag.out_rops[e[1]]:
                use(o)
    return peak
