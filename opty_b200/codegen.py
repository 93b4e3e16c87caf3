"""SymPy -> CUDA-C emitter: turns a :class:`CollocationProgram` into the source
of one sm_100a module.

This is the replacement for the C / Cython text templates of
``opty.utils.ufuncify_matrix`` (opty/utils.py:483-546, 743-818).  The module
contains

``opty_colloc_inv``
    single-thread kernel that evaluates the node-invariant sub-expressions
    (functions of the parameters and the time interval only) into a table that
    the host runtime copies to ``__constant__`` memory,

``opty_colloc_pre``
    pre-pass, one thread per node: evaluates the expensive sub-expressions
    that several output groups share (sines / cosines of the joint angles at
    the 10-link pendulum) once per node into *derived rows* appended to the
    trajectory matrix,

``opty_colloc_eval``
    the hot kernel, in one of two skeletons.  Row-stationary persistent
    kernel (``persistent=2``, the default where a body's input window fits in
    shared memory): one 8-warp block per SM, one equation row per group, a
    static schedule (:func:`stationary_schedule`) that keeps a row on the
    same SMs and makes all rows walk the node tiles at the same pace, the
    next item's input window prefetched by TMA, rows with node-invariant
    partials written as constant runs by bulk copies.  Grid kernel
    (``persistent=0``): grid = (node tiles, output groups).  A warp owns 32
    consecutive collocation nodes (lane = node) of one output group (a
    contiguous range of EOM rows); the block stages the tile's slice of the
    trajectory matrix in shared memory with TMA tile loads, every warp runs
    the group's straight-line float64 code -- ordered by the register-pressure
    scheduler of :mod:`opty_b200.schedule` --, writes the residuals eom-major
    and streams the node-major Jacobian block through per-warp shared-memory
    staging buffers that are drained by TMA tile stores, one *phase* (a few
    column runs of one equation) at a time,

``opty_module_info``
    the kernel geometry as a table of integers that the host runtime reads
    from the loaded module, so that no tuning parameter crosses the C-ABI.

The skeleton (staging, buffers, TMA, flush) is hand written in
``csrc/colloc_kernel.cuh``; only the arithmetic bodies and sizes come from
here.
"""

import os

from . import ir, schedule

EMITTER_VERSION = 8

_HERE = os.path.dirname(os.path.abspath(__file__))
KERNEL_HEADER = os.path.join(_HERE, 'csrc', 'colloc_kernel.cuh')

# layout of ``opty_module_info`` (mirrored in csrc/runtime.cu)
INFO_MAGIC = 0x4f505459   # 'OPTY'
INFO_WORDS = 40
MAX_MAPS = 8

PLAIN_LIVE_LIMIT = 120

SCHEDULE_DEFAULTS = {
    'schedule': 'auto',     # False: plain emission order (output by output);
                            # 'auto': schedule a group only if its plain
                            # order keeps more than PLAIN_LIVE_LIMIT values
                            # alive (it then cannot stay in registers).  At
                            # the 10-link pendulum the plain order fits 252
                            # registers and is 2-3 % faster than any
                            # scheduled variant (profiles/r02j_*); from 20
                            # links on it spills
    'reassociate': True,    # sums accumulate in arrival order
    'live_budget': 56,      # float64 values the schedule may keep alive
    'inline_cost': 2,       # values this cheap are never kept in a register
    'remat_cost': 24,       # values this cheap may be recomputed after a gap
    'volatile_loads': 'auto',  # input loads the compiler may not merge:
                            # 'auto' = for bodies of more than 3 000
                            # operations (they cannot afford the registers a
                            # merged load pins; small bodies gain more from
                            # the shorter instruction stream)
    'load_ahead': 0,        # input loads are emitted this many statements
                            # before the schedule's place for them
    'fence_every': 0,       # > 0: a warp-level memory fence after this many
                            # statements.  ptxas hoists global loads far
                            # ahead of their use to overlap their latency;
                            # on bodies of 10^4 statements that undoes the
                            # schedule and spills.  The fence bounds how far
                            # a load can move
}


def _lit(v):
    if v != v:
        return '__longlong_as_double(0x7ff8000000000000LL)'
    if v in (float('inf'), float('-inf')):
        return ('' if v > 0 else '-') + \
            '__longlong_as_double(0x7ff0000000000000LL)'
    s = repr(float(v))
    if 'e' not in s and '.' not in s:
        s += '.0'
    return s


_CMP = {ir.LT: '<', ir.LE: '<=', ir.EQ: '==', ir.NE: '!=', ir.AND: '&&',
        ir.OR: '||'}


def _op_expr(T, i, r):
    """C expression of tape operation ``i``; ``r(operand id)`` gives the text
    of an operand.  Returns ``(type, expression)``."""
    o = T.op[i]
    a, b, c = T.a[i], T.b[i], T.c[i]
    typ = 'double'
    if o == ir.NEG:
        e = '-{}'.format(r(a))
    elif o == ir.ADD:
        e = '{} + {}'.format(r(a), r(b))
    elif o == ir.SUB:
        e = '{} - {}'.format(r(a), r(b))
    elif o == ir.MUL:
        e = '{} * {}'.format(r(a), r(b))
    elif o == ir.DIV:
        e = '{} / {}'.format(r(a), r(b))
    elif o in ir.UNARY_MATH:
        e = '{}({})'.format(ir.OP_NAMES[o], r(a))
    elif o in (ir.POW, ir.ATAN2, ir.MIN, ir.MAX):
        e = '{}({}, {})'.format(ir.OP_NAMES[o], r(a), r(b))
    elif o == ir.SIGN:
        e = 'opty_sign({})'.format(r(a))
    elif o == ir.SEL:
        e = '({} ? {} : {})'.format(r(a), r(b), r(c))
    elif o in ir.BOOL_OPS:
        typ = 'bool'
        if o == ir.NOT:
            e = '!{}'.format(r(a))
        else:
            e = '({} {} {})'.format(r(a), _CMP[o], r(b))
    else:
        raise NotImplementedError(ir.OP_NAMES[o])
    return typ, e


class _BodyWriter(object):
    """Emits straight-line code for tape nodes on demand (depth-first from the
    outputs, so temporaries are defined close to their first use): the
    single-thread invariants kernel (``mode='inv'``) and the pre-pass kernel
    (``mode='pre'``, trajectory values from global memory)."""

    def __init__(self, prog, mode):
        self.prog = prog
        self.T = prog.tape
        self.mode = mode
        self.varying_ctx = mode != 'inv'
        self.done = set()
        self.lines = []
        self.num_ops = 0

    def ref(self, i):
        T = self.T
        o = T.op[i]
        if o == ir.CONST:
            return _lit(T.val[i])
        if self.varying_ctx:
            if o == ir.VIN:
                slot = T.a[i]
                return '{}({})'.format('GB' if slot & 1 else 'GA', slot >> 1)
            if not T.varying[i]:
                return 'CI({})'.format(self.prog.inv_index[i])
            return 'v{}'.format(i)
        if o == ir.UIN:
            return 'uni[{}]'.format(T.a[i])
        return 'w{}'.format(i)

    def need(self, root):
        """Makes sure ``root`` (and what it depends on) has been emitted."""
        T = self.T
        op_, a_, b_, c_ = T.op, T.a, T.b, T.c
        varying = T.varying
        done = self.done
        vctx = self.varying_ctx

        def is_leaf(i):
            if op_[i] <= ir.UIN:
                return True
            if vctx and not varying[i]:
                return True
            return i in done

        if is_leaf(root):
            return
        stack = [(root, False)]
        while stack:
            i, expanded = stack.pop()
            if i in done:
                continue
            if expanded:
                typ, e = _op_expr(T, i, self.ref)
                self.lines.append('const {} {} = {};'.format(
                    typ, ('v{}' if vctx else 'w{}').format(i), e))
                self.num_ops += 1
                done.add(i)
                continue
            stack.append((i, True))
            # push so that operand ``a`` is emitted first
            for o in (c_[i], b_[i], a_[i]):
                if o >= 0 and not is_leaf(o):
                    stack.append((o, False))


# ---------------------------------------------------------------------------
# phases: which columns of the node block are staged together
# ---------------------------------------------------------------------------
def _split_even_runs(col0, length, max_len, even):
    """Cuts the column run ``[col0, col0+length)`` into pieces of at most
    ``max_len`` columns.  With ``even`` every piece but possibly the last has
    an even length whose half is odd where possible: the 16-byte
    shared-memory stores of 8 consecutive lanes (row pitch 8*w bytes) then
    fall into distinct bank groups."""
    out = []
    c = col0
    end = col0 + length
    while c < end:
        w = min(max_len, end - c)
        if even and w > 2 and w % 2 == 0 and (w // 2) % 2 == 0:
            w -= 2
        if even and w % 2 and end - c > w:
            w -= 1
        out.append((c, w))
        c += w
    return out


def row_phases(col0, ncols, tile_cols, pair=None, even=True):
    """Phases of the column run ``[col0, col0+ncols)`` (one equation row, or
    two rows when ``P`` is odd): a list of phases, each a list of sub-tiles
    ``(first column, width)`` whose widths add up to at most ``tile_cols``.

    A row wider than ``tile_cols`` is cut so that the partial with respect to
    state ``k`` at the current node (column ``k``) and at the adjacent node
    (column ``pair + k``) land in the same phase: they share almost all of
    their operations (both discretisations replace ``x`` and ``x'`` by linear
    combinations of the two, opty/direct_collocation.py:2143-2156)."""
    if ncols <= tile_cols:
        return [_split_even_runs(col0, ncols, tile_cols, even)]
    phases = []
    if pair and 2 * pair <= ncols and not (even and pair % 2):
        w = max(2, tile_cols // 2)
        if even:
            w -= w % 2
            if w > 2 and (w // 2) % 2 == 0:
                w -= 2
        for k0 in range(0, pair, w):
            k1 = min(pair, k0 + w)
            phases.append([(col0 + k0, k1 - k0),
                           (col0 + pair + k0, k1 - k0)])
        tail0, tail = 2 * pair, ncols - 2 * pair
        if tail:
            used = sum(wd for _, wd in phases[-1])
            if used + tail <= tile_cols:
                phases[-1].extend(_split_even_runs(col0 + tail0, tail,
                                                   tile_cols, even))
            else:
                for piece in _split_even_runs(col0 + tail0, tail, tile_cols,
                                              even):
                    phases.append([piece])
        return phases
    return [[piece] for piece in _split_even_runs(col0, ncols, tile_cols,
                                                  even)]


def choose_tile_cols(requested, P, even=True):
    """Columns of one staging buffer.  ``'auto'``: one whole equation row (two
    when P is odd, so that a buffer starts at an even column) if that is at
    most 64 columns, otherwise 52 (two paired runs of 26)."""
    unit = P if (P % 2 == 0 or not even) else 2 * P
    if requested == 'auto':
        return unit if unit <= 64 else 52
    c = max(2, int(requested))
    if even:
        c -= c % 2
    return max(2, c)


# ---------------------------------------------------------------------------
# group bodies
# ---------------------------------------------------------------------------
class _GroupLayout(object):
    """Output slots, phases and sub-tiles of one output group."""

    def __init__(self, prog, gc0, gc1, tile_cols, pair, tma_store, tile_bufs,
                 map_index):
        P = prog.P
        T = prog.tape
        even = bool(tma_store)
        unit = P if (P % 2 == 0 or not even) else 2 * P
        self.slots = []        # (kind, ...) parallel to ``roots``
        self.roots = []
        self.phases = []       # per phase: dict(buf, subtiles, slots)
        assert gc0 % P == 0 and gc1 % P == 0
        c = gc0
        while c < gc1:
            ncols = min(unit, gc1 - c)
            rp = row_phases(c, ncols, tile_cols,
                            pair if ncols == P else None, even)
            for pi, subtiles in enumerate(rp):
                t = len(self.phases)
                ph = {'buf': t % tile_bufs, 'subtiles': [], 'slots': []}
                off = 0
                for (s0, w) in subtiles:
                    mi = map_index(w) if tma_store else 0
                    ph['subtiles'].append((mi, off, w, s0))
                    col = s0
                    while col < s0 + w:
                        pairable = (w % 2 == 0 and (col - s0) % 2 == 0 and
                                    col + 1 < s0 + w)
                        j0, k0 = divmod(col, P)
                        if pairable:
                            j1, k1 = divmod(col + 1, P)
                            self.slots.append(('js2', t, ph['buf'], off, w,
                                               col - s0, col))
                            self.roots.append((prog.jac[j0][k0],
                                               prog.jac[j1][k1]))
                            col += 2
                        else:
                            self.slots.append(('js1', t, ph['buf'], off, w,
                                               col - s0, col))
                            self.roots.append((prog.jac[j0][k0],))
                            col += 1
                        ph['slots'].append(len(self.slots) - 1)
                    off += 32 * w
                self.phases.append(ph)
                # the residual of a row is due with the row's last phase
                if pi == len(rp) - 1 and prog.con:
                    for j in range(c // P, (c + ncols) // P):
                        self.slots.append(('con', t, j))
                        self.roots.append((prog.con[j],))
                        ph['slots'].append(len(self.slots) - 1)
            c += ncols
        self.stored = sum(w for ph in self.phases
                          for (_, _, w, _) in ph['subtiles'])
        del T


class _ScheduledWriter(object):
    """Turns the events of a :class:`schedule.Schedule` into CUDA-C lines."""

    def __init__(self, prog, derived_index, sched):
        self.prog = prog
        self.T = prog.tape
        self.derived_index = derived_index
        self.sched = sched
        self.dag = sched.dag

    def leaf(self, i):
        T = self.T
        o = T.op[i]
        if o == ir.CONST:
            return _lit(T.val[i])
        if o == ir.VIN:
            slot = T.a[i]
            return '{}({})'.format('XB' if slot & 1 else 'XA', slot >> 1)
        if i in self.derived_index:
            return 'XD({})'.format(self.derived_index[i])
        if not T.varying[i]:
            return 'CI({})'.format(self.prog.inv_index[i])
        raise AssertionError('not a leaf: {}'.format(i))

    def expr(self, consumer, i):
        """Text of value ``i`` as read inside ``consumer``."""
        dag = self.dag
        kind = dag.kind_of(i)
        if kind == 0:
            return self.leaf(i)
        if i in dag.regset and (consumer, i) not in dag.inline_edges:
            return 'v{}'.format(i)
        if kind == 1:
            return self.leaf(i)
        if i in dag.terms:
            return '(' + self.sum_text(consumer, dag.terms[i]) + ')'
        _, e = _op_expr(self.T, i, lambda o: self.expr(consumer, o))
        return '(' + e + ')'

    def sum_text(self, consumer, terms):
        out = []
        for sg, t in terms:
            e = self.expr(consumer, t)
            if not out:
                out.append(e if sg > 0 else '-' + e)
            else:
                out.append((' + ' if sg > 0 else ' - ') + e)
        return ''.join(out)

    def lines_for(self, layout, fence_every=0, load_ahead=0):
        lines, num_ops, loads = self._lines_for(layout, fence_every)
        if load_ahead > 0:
            # input loads issued `load_ahead` statements before the place the
            # schedule gave them: with direct input loads (large models) a
            # load is an L2 access of several hundred cycles that nothing
            # else in the warp covers
            for i in loads:
                j = max(0, i - load_ahead)
                if j < i:
                    lines.insert(j, lines.pop(i))
        return lines, num_ops

    def _lines_for(self, layout, fence_every=0):
        dag = self.dag
        T = self.T
        lines = []
        loads = []
        slots = layout.slots
        phases = layout.phases
        deferred = self.sched.deferred
        left = [sum(1 for k in ph['slots']
                    if slots[k][0] != 'con' and k not in deferred)
                for ph in phases]
        drained = False
        begun = [False] * len(phases)
        num_ops = 0
        since_fence = 0
        for e in self.sched.events:
            kind = e[0]
            if fence_every and kind != schedule.OUT:
                since_fence += 1
                if since_fence >= fence_every:
                    lines.append('OPTY_FENCE();')
                    since_fence = 0
            if kind == schedule.LOAD:
                v = e[1]
                loads.append(len(lines))
                lines.append('const double v{} = {};'.format(v, self.leaf(v)))
            elif kind == schedule.OP:
                v = e[1]
                if v in dag.terms:
                    lines.append('const double v{} = {};'.format(
                        v, self.sum_text(v, dag.terms[v])))
                else:
                    typ, ex = _op_expr(T, v, lambda o, v=v: self.expr(v, o))
                    lines.append('const {} v{} = {};'.format(typ, v, ex))
                num_ops += 1
            elif kind == schedule.ACC:
                s, j, first = e[1], e[2], e[3]
                terms = dag.terms[s]
                if first:
                    # terms that read no register value are folded in here
                    start = [terms[j]] + [
                        t for k, t in enumerate(terms)
                        if k != j and not dag.term_rops[s][k]]
                    lines.append('double v{} = {};'.format(
                        s, self.sum_text(s, start)))
                else:
                    sg, t = terms[j]
                    lines.append('v{} {}= {};'.format(
                        s, '+' if sg > 0 else '-', self.expr(s, t)))
                num_ops += 1
            else:
                k = e[1]
                slot = slots[k]
                consumer = (schedule.OUT, k)
                vals = [self.expr(consumer, r) for r in dag.outputs[k]]
                if slot[0] == 'con':
                    lines.append('OPTY_CON({}, {});'.format(slot[2], vals[0]))
                    continue
                if k in deferred:
                    # stored straight to global memory, after the tile store
                    # of the phase that owns the column has completed
                    if not drained:
                        lines.append('OPTY_DRAIN_WRITES();')
                        drained = True
                    for c, v in enumerate(vals):
                        lines.append('OPTY_JG({}, {});'.format(slot[6] + c, v))
                    continue
                t = slot[1]
                if not begun[t]:
                    begun[t] = True
                    lines.append('OPTY_PHASE_BEGIN({});'.format(t))
                if slot[0] == 'js2':
                    lines.append('OPTY_JS2({}, {}, {}, {}, {}, {});'.format(
                        slot[2], slot[3], slot[4], slot[5], vals[0], vals[1]))
                else:
                    lines.append('OPTY_JS1({}, {}, {}, {}, {});'.format(
                        slot[2], slot[3], slot[4], slot[5], vals[0]))
                left[t] -= 1
                if left[t] == 0:
                    lines.append('OPTY_FLUSH_BEGIN()')
                    for (mi, off, w, col0) in phases[t]['subtiles']:
                        lines.append('OPTY_TSTORE({}, {}, {}, {}, {})'.format(
                            mi, phases[t]['buf'], off, w, col0))
                    lines.append('OPTY_FLUSH_END()')
        assert all(v == 0 for v in left)
        lines.append('OPTY_DRAIN();')
        return lines, num_ops, loads


_INPUT_REF = __import__('re').compile(r'\bX([ABD])\((\d+)\)')


def stationary_schedule(costs, n_tiles, n_slots, strided=True):
    """Static work assignment of the row-stationary kernel.  ``costs[g]`` is
    the estimated time of one item (one node tile of a block x group ``g``);
    every group has ``n_tiles`` items.  Returns one list of segments
    ``(group, first tile, tiles)`` per slot (= resident block).

    Heavy groups (at least a quarter of the most expensive one) are laid out
    group by group, tile by tile along the slots in proportion to their cost,
    so that a block keeps one body -- its instructions stay in the SM's
    instruction cache from one item, and one launch, to the next.  The light
    groups (store-only rows such as ``x' = v``) are then dealt out item by
    item to the least loaded slots and interleaved with the heavy items, which
    spreads their output over all SMs and over the whole launch."""
    import heapq
    G = len(costs)
    cmax = max(costs) if costs else 1.0
    heavy = [g for g in range(G) if costs[g] >= 0.25 * cmax]
    light = [g for g in range(G) if costs[g] < 0.25 * cmax]
    heavy.sort(key=lambda g: -costs[g])
    load = [0.0] * n_slots
    hv = [[] for _ in range(n_slots)]
    if strided and 0 < len(heavy) <= n_slots:
        # whole slots per group (one more slot to the group whose busiest
        # slot is the busiest, until none is left), tile t of a group on the
        # group's slot t mod S, in step t div S: every group then walks the
        # node tiles at the same pace, and the pieces of a node's Jacobian
        # row reach memory within about one item time of each other -- the
        # store stream is 15-20 % faster that way (tools/write_path_bench.cu,
        # variants I/J/K)
        S = {g: 1 for g in heavy}
        worst = [(-(-(-n_tiles // S[g])) * costs[g], g) for g in heavy]
        heapq.heapify(worst)
        for _ in range(n_slots - len(heavy)):
            _, g = heapq.heappop(worst)
            S[g] += 1
            heapq.heappush(worst, (-(-(-n_tiles // S[g])) * costs[g], g))
        base = 0
        for g in heavy:
            for t in range(n_tiles):
                s_ = base + t % S[g]
                hv[s_].append([g, t, 1])
                load[s_] += costs[g]
            base += S[g]
    else:
        total_heavy = float(sum(costs[g] for g in heavy) * n_tiles) or 1.0
        per_slot = total_heavy / n_slots
        cum = 0.0
        for g in heavy:
            for t in range(n_tiles):
                s = min(n_slots - 1, int((cum + 0.5 * costs[g]) / per_slot))
                cum += costs[g]
                load[s] += costs[g]
                if hv[s] and hv[s][-1][0] == g and \
                        hv[s][-1][1] + hv[s][-1][2] == t:
                    hv[s][-1][2] += 1
                else:
                    hv[s].append([g, t, 1])
    lt = [[] for _ in range(n_slots)]
    heap = [(load[s], s) for s in range(n_slots)]
    heapq.heapify(heap)
    for g in light:
        for t in range(n_tiles):
            ld, s = heapq.heappop(heap)
            lt[s].append([g, t, 1])
            heapq.heappush(heap, (ld + costs[g], s))
    out = []
    for s in range(n_slots):
        # heavy items one by one, a light item after each
        hitems = [(g, t0 + k) for g, t0, n in hv[s] for k in range(n)]
        litems = [(g, t) for g, t, _ in lt[s]]
        seq = []
        per = -(-len(litems) // max(len(hitems), 1))
        li = 0
        for it in hitems:
            seq.append(it)
            seq.extend(litems[li:li + per])
            li += per
        seq.extend(litems[li:])
        segs = []
        for g, t in seq:
            if segs and segs[-1][0] == g and segs[-1][1] + segs[-1][2] == t:
                segs[-1][2] += 1
            else:
                segs.append([g, t, 1])
        out.append([tuple(x) for x in segs])
    return out


# shared state of the worker processes that emit group bodies in parallel
# (set before the pool forks; the tape is far too large to pickle per task)
_WORK = {}


def _emit_group(g):
    prog = _WORK['prog']
    opts = _WORK['opts']
    gc0, gc1 = _WORK['groups'][g]
    widths = _WORK['widths']

    def map_index(w):
        return widths.index(w)
    layout = _GroupLayout(prog, gc0, gc1, _WORK['tile_cols'], _WORK['pair'],
                          _WORK['tma_store'], _WORK['tile_bufs'], map_index)
    stop = _WORK['stop']
    phase_lists = [ph['slots'] for ph in layout.phases]
    use_scheduler = opts['schedule']
    plain = None
    if use_scheduler == 'auto':
        plain = schedule.plain_order(prog.tape, layout.roots, stop)
        use_scheduler = plain.peak_live > PLAIN_LIVE_LIMIT
    if use_scheduler:
        sched = schedule.schedule_body(
            prog.tape, layout.roots, stop, phases=phase_lists,
            reassociate=opts['reassociate'], inline_cost=opts['inline_cost'],
            remat_cost=opts['remat_cost'], live_budget=opts['live_budget'],
            deferrable=[] if _WORK.get('no_defer') else
            [k for k, sl in enumerate(layout.slots) if sl[0] != 'con'])
    else:
        # plain order: slots in layout order (phases are contiguous there)
        sched = plain or schedule.plain_order(prog.tape, layout.roots, stop)
    writer = _ScheduledWriter(prog, _WORK['derived_index'], sched)
    lines, num_ops = writer.lines_for(layout, int(opts['fence_every']),
                                      int(opts.get('load_ahead', 0)))
    meta = {'rows': [gc0 // prog.P, gc1 // prog.P], 'cols': [gc0, gc1],
            'col0': gc0, 'ncols': layout.stored, 'ops': sched.num_ops,
            'statements': num_ops, 'peak_live': sched.peak_live,
            'phases': len(layout.phases), 'deferred': len(sched.deferred),
            'scheduled': bool(use_scheduler)}
    return lines, meta


def group_widths(prog, groups, tile_cols, pair, tma_store):
    """Distinct sub-tile widths of all groups (one output tensor map each)."""
    widths = []
    P = prog.P
    even = bool(tma_store)
    unit = P if (P % 2 == 0 or not even) else 2 * P
    for gc0, gc1 in groups:
        c = gc0
        while c < gc1:
            ncols = min(unit, gc1 - c)
            for subtiles in row_phases(c, ncols, tile_cols,
                                       pair if ncols == P else None, even):
                for _, w in subtiles:
                    if w not in widths:
                        widths.append(w)
            c += ncols
    return sorted(widths)


def emit_module(prog, groups, method, tile_cols='auto', warps_per_block=2,
                min_blocks_per_sm=4, tma_load=True, tma_store=True,
                derived=(), debug_nostore=False, tile_bufs=2,
                only_groups=None, with_aux=True, pair=None,
                schedule_options=None, workers=1, persistent=False,
                num_sms=148, num_nodes=None, blocks_per_sm=1, const_rows=(),
                const_head_pct=(35, 50, 15, 70), store_hint=0, fused_pre=False,
                item_cost=16000, strided_schedule=True, const_pre_pct=0,
                tile_major=False, const_kernel=False):
    """Returns ``(source_text, meta)`` for ``prog`` split into ``groups``
    (list of ``(c0, c1)`` column ranges of the flattened ``M*P`` node block,
    whole equations each; a group also owns the residuals of its rows).
    ``derived`` lists the tape ids that the pre-pass kernel evaluates once
    per node into derived rows.  ``pair``: number of states ``n`` when the
    columns ``k`` and ``n + k`` of a row are the partials with respect to the
    same state at the two nodes of the stencil (collocation programs), else
    None.

    Large problems are compiled as several modules in parallel: with
    ``only_groups = (g0, g1)`` the module contains the bodies and the main
    kernel of groups ``g0 .. g1-1`` only; ``with_aux=False`` leaves out the
    invariants and pre-pass kernels (the first module carries them)."""
    T = prog.tape
    M, P, K, R = prog.M, prog.P, prog.K, prog.R
    tma_store = bool(tma_store) and K % 2 == 0
    C = choose_tile_cols(tile_cols, P, even=tma_store)
    persistent = int(persistent)
    stationary = persistent == 2
    if stationary:
        if num_nodes is None or not tma_store or int(tma_load) != 1:
            raise ValueError('the row-stationary kernel needs the node count, '
                             'TMA stores and staged (row-major) input')
        if only_groups is not None:
            raise ValueError('the row-stationary kernel is one module')
    if warps_per_block > 4 and warps_per_block % 4:
        raise ValueError('warps_per_block above 4 must be a multiple of 4')
    tile_bufs = int(tile_bufs)
    if tile_bufs not in (1, 2):
        raise ValueError('tile_bufs must be 1 or 2')
    opts = dict(SCHEDULE_DEFAULTS)
    opts.update(schedule_options or {})
    ninv = len(prog.inv_nodes)
    # constant rows (row-stationary kernel): equations whose partials are all
    # literals or node-invariant values (x' = v and the like).  Their part of
    # every node's Jacobian row is the same run of numbers: the invariants
    # kernel appends it to its table, the main kernel copies it to every node
    # with plain 16-byte stores, the pre-pass evaluates their residuals.
    const_rows = sorted(const_rows)
    const_runs = []         # (first column, doubles, offset into the values)
    const_entries = []
    for j in const_rows:
        if const_runs and const_runs[-1][0] + const_runs[-1][1] == j * P:
            const_runs[-1][1] += P
        else:
            const_runs.append([j * P, P, len(const_entries)])
        const_entries.extend(prog.jac[j])
    if const_rows and not (tma_store and P % 2 == 0):
        raise ValueError('constant rows need TMA stores and an even number '
                         'of partials per equation')
    if const_rows and not stationary and len(const_entries) * 8 > 200 * 1024:
        raise ValueError('the constant runs do not fit in shared memory')
    cval0 = ninv + (ninv & 1)
    if const_rows:
        ninv = cval0 + len(const_entries)
    derived = list(derived)
    derived_index = {nid: k for k, nid in enumerate(derived)}
    D = len(derived)
    groups = [tuple(g) for g in groups]
    g0, g1 = only_groups if only_groups is not None else (0, len(groups))
    widths = group_widths(prog, groups, C, pair, tma_store)
    if tma_store and len(widths) > MAX_MAPS:
        raise ValueError('The module needs {} distinct staging tile widths, '
                         'at most {} are supported.'.format(
                             len(widths), MAX_MAPS))
    if not tma_store:
        map_widths = [2]        # unused descriptor slot
    else:
        map_widths = widths

    out = []
    w = out.append
    w('// generated by opty_b200.codegen (emitter v{}); do not edit'.format(
        EMITTER_VERSION))
    w('#define OPTY_NMAPS {}'.format(len(map_widths)))
    w('#define OPTY_M {}'.format(M))
    w('#define OPTY_P {}'.format(P))
    w('#define OPTY_K {}'.format(K))
    w('#define OPTY_R {}'.format(R))
    w('#define OPTY_D {}'.format(D))
    w('#define OPTY_TILE_DOUBLES {}'.format(32 * C))
    w('#define OPTY_NGROUPS {}'.format(g1 - g0))
    # (the constant table holds the invariants the bodies read, not the
    # constant column runs behind them)
    w('#define OPTY_NINV {}'.format(max(cval0 if const_rows else ninv, 1)))
    w('#define OPTY_NUNI {}'.format(max(prog.num_uniform, 1)))
    w('#define OPTY_WARPS {}'.format(warps_per_block))
    w('#define OPTY_MIN_BLOCKS {}'.format(min_blocks_per_sm))
    w('#define OPTY_TMA_LOAD {}'.format(int(tma_load)))
    w('#define OPTY_TMA_STORE {}'.format(1 if tma_store else 0))
    w('#define OPTY_NBUF {}'.format(tile_bufs))
    w('#define OPTY_STORE_HINT {}'.format(int(store_hint)))
    if tile_major and not persistent:
        w('#define OPTY_TILE_MAJOR 1')
    vol = opts['volatile_loads']
    if vol == 'auto':
        stop_set = set(derived) or None
        vol = max(prog.range_cost(c0, c1, stop_set)
                  for c0, c1 in groups[g0:g1]) > 3000
    w('#define OPTY_VOLATILE_LOADS {}'.format(1 if vol else 0))
    if persistent:
        w('#define OPTY_PERSISTENT {}'.format(persistent))
        w('#define OPTY_SM_TABLE {}'.format(num_sms))
    if stationary:
        w('#define OPTY_NSLOTS {}'.format(num_sms * blocks_per_sm))
        w('#define OPTY_BPS {}'.format(blocks_per_sm))
        # placeholder, replaced once the bodies' input windows are known
        w('@@XROWS_MAX@@')
    w('#define OPTY_PRE_GROUPS {}'.format(
        len(set(T.a[nid] for nid in derived)) + (1 if const_rows else 0) +
        (8 if (const_rows and with_aux and int(const_pre_pct) > 0) else 0)))
    w('#define OPTY_NCRUNS {}'.format(len(const_runs)))
    w('#define OPTY_NCONST {}'.format(len(const_entries)))
    if const_runs and not stationary and not const_kernel:
        # the grid kernel's blocks send the runs themselves
        w('#define OPTY_GRID_CONST 1')
        w('#define OPTY_G0 {}'.format(g0))
        w('#define OPTY_NGROUPS_ALL {}'.format(len(groups)))
    chead = list(const_head_pct) if isinstance(
        const_head_pct, (list, tuple)) else [const_head_pct, 50, 15, 70]
    w('#define OPTY_CONST_HEAD_PCT {}'.format(int(chead[0])))
    w('#define OPTY_CONST_FIRST_PCT {}'.format(int(chead[1])))
    w('#define OPTY_CONST_ITEM_PCT {}'.format(int(chead[2])))
    w('#define OPTY_CONST_ALIGN_PCT {}'.format(
        int(chead[3]) if len(chead) > 3 else 100))
    tick_pct = int(chead[4]) if len(chead) > 4 else 0
    ticks_per_body = int(chead[5]) if len(chead) > 5 else 3
    if not (stationary and const_runs):
        tick_pct = 0
    w('#define OPTY_CONST_TICK_PCT {}'.format(tick_pct))
    if debug_nostore:
        w('#define OPTY_DEBUG_NOSTORE {}'.format(int(debug_nostore)))
    w('#include "colloc_kernel.cuh"')
    w('')

    # ---- invariants kernel -------------------------------------------
    inv_ops = 0
    if with_aux:
        bw = _BodyWriter(prog, 'inv')
        for nid in prog.inv_nodes:
            bw.need(nid)
        for nid in const_entries:
            bw.need(nid)
        w('extern "C" __global__ void opty_colloc_inv('
          'const double* __restrict__ uni, double* __restrict__ inv)')
        w('{')
        w('  if (threadIdx.x != 0 || blockIdx.x != 0) return;')
        for line in bw.lines:
            w('  ' + line)
        for k, nid in enumerate(prog.inv_nodes):
            w('  inv[{}] = {};'.format(k, bw.ref(nid)))
        for k, nid in enumerate(const_entries):
            w('  inv[{}] = {};'.format(cval0 + k, bw.ref(nid)))
        w('}')
        w('')
        inv_ops = bw.num_ops

    # ---- pre-pass kernel: derived rows, grid.y = groups of derived rows ---
    # rows that share their argument (sin / cos of the same angle) stay
    # together so that the argument is computed once
    pre_chunks = []
    by_arg = {}
    for k, nid in enumerate(derived):
        by_arg.setdefault(T.a[nid], []).append(k)
    for ks in by_arg.values():
        pre_chunks.append(ks)
    const_pre_pct = int(const_pre_pct) if (const_runs and with_aux) else 0
    pre_groups = len(pre_chunks) + (1 if const_rows else 0) + \
        (8 if const_pre_pct > 0 else 0)
    pre_ops = 0
    if const_runs:
        w('__device__ const int opty_crun[OPTY_NCRUNS][3] = {{{}}};'.format(
            ', '.join('{{{}, {}, {}}}'.format(c0, n // 2, off)
                      for c0, n, off in const_runs)))
    if const_pre_pct > 0:
        out.insert(out.index('#include "colloc_kernel.cuh"'),
                   '#define OPTY_CONST_PRE_PCT {}'.format(const_pre_pct))
    fused_pre = bool(fused_pre) and stationary and with_aux and \
        0 < pre_groups <= num_sms * blocks_per_sm
    if fused_pre:
        out.insert(out.index('#include "colloc_kernel.cuh"'),
                   '#define OPTY_FUSED_PRE 1')
    if with_aux:
        case_lines = []
        for pg, ks in enumerate(pre_chunks):
            bw = _BodyWriter(prog, 'pre')
            for k in ks:
                bw.need(derived[k])
                bw.lines.append('OPTY_DRV({}, {});'.format(
                    k, bw.ref(derived[k])))
            case_lines.append('    case {}: {{'.format(pg))
            case_lines.extend('      ' + line for line in bw.lines)
            case_lines.append('    } break;')
            pre_ops += bw.num_ops
        if const_rows:
            bw = _BodyWriter(prog, 'pre')
            # all loads before the first store: the compiler cannot tell that
            # the residual rows do not alias the trajectory matrix
            for j in const_rows:
                bw.need(prog.con[j])
            for j in const_rows:
                bw.lines.append('OPTY_PCON({}, {});'.format(
                    j, bw.ref(prog.con[j])))
            case_lines.append('    case {}: {{'.format(len(pre_chunks)))
            case_lines.extend('      ' + line for line in bw.lines)
            case_lines.append('    } break;')
            pre_ops += bw.num_ops
        w('extern "C" __global__ void __launch_bounds__(OPTY_PRE_THREADS)')
        w('opty_colloc_pre(const OptyParams p)')
        w('{')
        w('  OPTY_PRE_BEGIN();')
        w('  switch (opty_pg) {')
        for line in case_lines:
            w(line)
        w('    default: break;')
        w('  }')
        w('  OPTY_PRE_END()')
        w('}')
        w('')
        if const_runs and not stationary and const_kernel:
            w('extern "C" __global__ void __launch_bounds__(256)')
            w('opty_colloc_const(const OptyParams p)')
            w('{')
            w('  OPTY_CONST_KERNEL_BODY()')
            w('}')
            w('')
        if fused_pre:
            # the same cases as a device function: phase 0 of the
            # row-stationary kernel (one case per block)
            w('static __device__ __forceinline__ void opty_pre_case('
              'const OptyParams& p, const int* nodes, const int opty_pg)')
            w('{')
            w('  switch (opty_pg) {')
            for line in case_lines:
                if line.startswith('    case '):
                    w(line)
                    w('      _Pragma("unroll") for (int u_ = 0; u_ < '
                      'OPTY_PRE_ILP; ++u_) {')
                    w('      const int node = nodes[u_];')
                    w('      const double* xg = p.traj + node;')
                    w('      double* drv = p.traj + (long long)OPTY_R * p.ldt '
                      '+ node;')
                elif line.startswith('    } break;'):
                    w('      }')
                    w(line)
                else:
                    w(line)
            w('    default: break;')
            w('  }')
            w('}')
            w('')

    if stationary:
        # do the input windows fit?  (before the bodies are scheduled: the
        # rows a group reads follow from the tape)
        stop_set = set(derived)
        rows_max = 1
        for gc0, gc1 in groups:
            used = set()
            for i in prog._reachable_stop(prog.range_roots(gc0, gc1),
                                          stop_set):
                if T.op[i] == ir.VIN:
                    used.add(T.a[i] >> 1)
                elif i in derived_index:
                    used.add(R + derived_index[i])
            if used:
                rows_max = max(rows_max, max(used) + 1 - min(used))
        threads = 32 * warps_per_block
        xseg = min(threads, 128)
        need = warps_per_block * tile_bufs * 32 * C * 8 + 128 + \
            2 * (threads // xseg) * (-(-(rows_max * (xseg + 2) * 8) // 128)
                                     * 128) + \
            (-(-len(const_entries) * 8 // 128) * 128)
        if need > 227 * 1024:
            raise ValueError(
                'the row-stationary kernel would need {} bytes of shared '
                'memory per block ({} staging buffers of {} columns for {} '
                'warps, two input buffers of {} rows)'.format(
                    need, tile_bufs, C, warps_per_block, rows_max))

    # ---- group bodies (scheduled, possibly in parallel) ------------------
    _WORK.update(prog=prog, opts=opts, groups=groups, widths=map_widths,
                 tile_cols=C, pair=pair, tma_store=tma_store,
                 tile_bufs=tile_bufs, stop=frozenset(derived),
                 derived_index=derived_index, no_defer=stationary)
    todo = list(range(g0, g1))
    workers = max(1, min(int(workers), len(todo)))
    if workers > 1:
        import multiprocessing
        with multiprocessing.get_context('fork').Pool(workers) as pool:
            results = pool.map(_emit_group, todo, chunksize=1)
    else:
        results = [_emit_group(g) for g in todo]
    _WORK.clear()
    group_meta = []
    for g, (lines, gm) in zip(todo, results):
        if stationary:
            # input window of the body: the contiguous range of trajectory /
            # derived rows it reads (staged per item, csrc/colloc_kernel.cuh)
            used = set()
            for line in lines:
                for kind, idx in _INPUT_REF.findall(line):
                    used.add(int(idx) + (R if kind == 'D' else 0))
            row0 = min(used) if used else 0
            gm['xrow0'] = row0
            gm['xrows'] = (max(used) + 1 - row0) if used else 1
            w('#undef OPTY_XROW0')
            w('#define OPTY_XROW0 {}'.format(row0))
        w('static __device__ __forceinline__ void opty_group_{}('
          'const OptyCtx& ctx)'.format(g))
        w('{')
        if tick_pct > 0 and len(lines) > 8 * ticks_per_body:
            # a trickle of constant runs: ticks at even distances, in front
            # of a definition (never inside a flush sequence)
            marks = [len(lines) * (k + 1) // (ticks_per_body + 1)
                     for k in range(ticks_per_body)]
            lines = list(lines)
            for m in reversed(marks):
                while m < len(lines) and not (
                        lines[m].startswith('const double') or
                        lines[m].startswith('double ')):
                    m += 1
                lines.insert(m, 'OPTY_TICK();')
        for line in lines:
            w('  ' + line)
        w('}')
        w('')
        group_meta.append(gm)

    # blockIdx.y -> group: most expensive groups are launched first
    order = sorted(range(len(group_meta)),
                   key=lambda g: -(20 * group_meta[g]['ops'] +
                                   43 * group_meta[g]['ncols']))
    w('__device__ const int opty_group_order[OPTY_NGROUPS] = {{{}}};'.format(
        ', '.join(str(g) for g in order)))
    if persistent == 1:
        # slot = position in opty_group_order.  SMs are dealt out to the
        # slots in proportion to the slots' cost; every block of an SM starts
        # on the SM's slot
        costs = [20 * group_meta[g]['ops'] + 43 * group_meta[g]['ncols']
                 for g in order]
        total = float(sum(costs)) or 1.0
        table = []
        acc = 0.0
        slot = 0
        for sm in range(num_sms):
            target = (sm + 0.5) / num_sms * total
            while slot < len(costs) - 1 and acc + costs[slot] < target:
                acc += costs[slot]
                slot += 1
            table.append(slot)
        w('__device__ const int opty_sm_group[OPTY_SM_TABLE] = {{{}}};'.format(
            ', '.join(str(v) for v in table)))
        w('__device__ const int opty_slot_cost[OPTY_NGROUPS] = {{{}}};'.format(
            ', '.join(str(max(1, int(c // 100))) for c in costs)))
    smem_bytes = 0
    if stationary:
        threads = 32 * warps_per_block
        n_tiles = -(-int(num_nodes) // threads)
        n_slots = num_sms * blocks_per_sm
        # measured at the 10-link pendulum: an item takes 1.3 us + 1.4 ns per
        # operation (profiles/r02z_*)
        costs = [20.0 * gm['ops'] + 43.0 * gm['ncols'] + float(item_cost)
                 for gm in group_meta]
        sched = stationary_schedule(costs, n_tiles, n_slots,
                                    strided=bool(strided_schedule))
        starts, segs = [0], []
        for slot in sched:
            segs.extend(slot)
            starts.append(len(segs))
        xrows_max = max(gm['xrows'] for gm in group_meta)
        out[out.index('@@XROWS_MAX@@')] = \
            '#define OPTY_XROWS_MAX {}'.format(xrows_max)
        w('__device__ const int opty_group_xrow0[OPTY_NGROUPS] = {{{}}};'.format(
            ', '.join(str(gm['xrow0']) for gm in group_meta)))
        w('__device__ const int opty_group_xrows[OPTY_NGROUPS] = {{{}}};'.format(
            ', '.join(str(gm['xrows']) for gm in group_meta)))
        w('__device__ const int opty_group_odd[OPTY_NGROUPS] = {{{}}};'.format(
            ', '.join(str(gm['phases'] & 1) for gm in group_meta)))
        # (constant memory while the tables are small: the first item of a
        # block waits for them)
        space = '__constant__' if len(segs) <= 2048 else '__device__ const'
        w('{} int opty_sched_slot[OPTY_NSLOTS + 1] = {{{}}};'
          .format(space, ', '.join(str(v) for v in starts)))
        w('{} int4 opty_sched_seg[{}] = {{{}}};'.format(space, 
            max(len(segs), 1),
            ', '.join('{{{}, {}, {}, {}}}'.format(
                g, t0, nt, group_meta[g]['xrow0'] |
                (group_meta[g]['xrows'] << 16)) for g, t0, nt in segs) or
            '{0, 0, 0, 0}'))
        xseg = min(threads, 128)
        xbuf = (threads // xseg) * (
            -(-(xrows_max * (xseg + 2) * 8) // 128) * 128)
        smem_bytes = warps_per_block * tile_bufs * 32 * C * 8 + 2 * xbuf + 128
        if const_runs:
            smem_bytes += -(-len(const_entries) * 8 // 128) * 128
        if smem_bytes > 227 * 1024:
            raise ValueError(
                'the row-stationary kernel needs {} bytes of shared memory per '
                'block ({} staging buffers of {} columns for {} warps, two '
                'input buffers of {} rows); use fewer warps, one staging '
                'buffer or narrower tiles'.format(
                    smem_bytes, tile_bufs, C, warps_per_block, xrows_max))
    w('')
    info = [0] * INFO_WORDS
    info[0] = INFO_MAGIC
    info[1] = EMITTER_VERSION
    info[2] = warps_per_block
    info[3] = g1 - g0
    info[4] = D
    info[5] = pre_groups
    info[6] = int(tma_load)
    info[7] = 1 if tma_store else 0
    info[8] = tile_bufs
    info[9] = 32 * C
    info[10] = len(map_widths)
    for k, mw in enumerate(map_widths):
        info[11 + k] = mw
    info[19] = ninv
    info[20] = R
    info[21] = M
    info[22] = P
    info[23] = 1 if with_aux else 0
    info[24] = persistent
    info[25] = min_blocks_per_sm
    if stationary:
        info[26] = smem_bytes
        info[27] = num_sms * blocks_per_sm
        info[28] = n_tiles
        info[29] = cval0 if const_rows else 0
        info[30] = 1 if fused_pre else 0
        info[31] = xrows_max
    if const_runs:
        info[29] = cval0
        info[32] = len(const_entries)
        info[33] = 1 if (const_kernel and not stationary) else 0
    w('extern "C" __device__ const int opty_module_info[{}] = {{{}}};'.format(
        INFO_WORDS, ', '.join(str(v) for v in info)))
    w('')
    w('extern "C" __global__ void __launch_bounds__(OPTY_THREADS, '
      'OPTY_MIN_BLOCKS)')
    w('opty_colloc_eval(const __grid_constant__ OptyTmaps tm, '
      'const OptyParams p)')
    w('{')
    w('  OPTY_KERNEL_BEGIN()')
    w('  switch (opty_g) {')
    for g in range(g0, g1):
        w('    case {}: opty_group_{}(ctx); break;'.format(g - g0, g))
    w('    default: break;')
    w('  }')
    w('  OPTY_KERNEL_END()')
    w('}')
    w('')

    meta = {
        'emitter_version': EMITTER_VERSION,
        'M': M, 'P': P, 'K': K, 'R': R, 'C': C, 'D': D,
        'num_groups': g1 - g0,
        'groups': group_meta,
        'group_range': [g0, g1],
        'map_widths': list(map_widths),
        'num_inv': ninv,
        'num_uniform': prog.num_uniform,
        'inv_ops': inv_ops,
        'pre_ops': pre_ops,
        'pre_groups': pre_groups,
        'group_order': order,
        'warps_per_block': warps_per_block,
        'min_blocks_per_sm': min_blocks_per_sm,
        'tma_load': int(tma_load),
        'tma_store': bool(tma_store),
        'tile_bufs': tile_bufs,
        'persistent': persistent,
        'smem_bytes': smem_bytes,
        'const_rows': list(const_rows),
        'fused_pre': bool(fused_pre),
        'method': method,
        'schedule': opts,
        'entry_kind': prog.entry_kind(),
        'stats': prog.stats(),
    }
    return '\n'.join(out), meta
