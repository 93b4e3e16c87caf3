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
    the hot kernel, persistent: one block slot per SM x ``min_blocks``.  A
    warp owns 32 consecutive collocation nodes (lane = node) of one output
    group (a contiguous range of EOM rows) per tile; it stages the tile's
    slice of the trajectory matrix in shared memory with TMA tile loads, runs
    the group's straight-line float64 code, writes the residuals eom-major and
    streams the node-major Jacobian block through a double-buffered
    shared-memory tile that is drained by TMA tile stores.

The skeleton (staging, tiles, TMA, flush) is hand written in
``csrc/colloc_kernel.cuh``; only the arithmetic bodies and sizes come from
here.
"""

import os

from . import ir

EMITTER_VERSION = 5

_HERE = os.path.dirname(os.path.abspath(__file__))
KERNEL_HEADER = os.path.join(_HERE, 'csrc', 'colloc_kernel.cuh')


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


class _BodyWriter(object):
    """Emits straight-line code for tape nodes on demand (depth-first from the
    outputs, so temporaries are defined close to their first use).

    ``mode``: ``'inv'`` single-thread invariants kernel, ``'pre'`` pre-pass
    kernel (trajectory values from global memory), ``'main'`` group bodies
    (trajectory values and derived rows from the staged shared-memory tile).
    """

    def __init__(self, prog, mode, derived_index=None):
        self.prog = prog
        self.T = prog.tape
        self.mode = mode
        self.varying_ctx = mode != 'inv'
        self.derived_index = derived_index or {}
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
                if self.mode == 'pre':
                    return '{}({})'.format('GB' if slot & 1 else 'GA',
                                           slot >> 1)
                return '{}({})'.format('XB' if slot & 1 else 'XA', slot >> 1)
            if not T.varying[i]:
                return 'CI({})'.format(self.prog.inv_index[i])
            if self.mode == 'main' and i in self.derived_index:
                return 'XD({})'.format(self.derived_index[i])
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
        derived = self.derived_index if self.mode == 'main' else {}

        def is_leaf(i):
            o = op_[i]
            if o <= ir.UIN:
                return True
            if vctx and not varying[i]:
                return True
            if i in derived:
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
                self._emit(i)
                done.add(i)
                continue
            stack.append((i, True))
            # push so that operand ``a`` is emitted first
            for o in (c_[i], b_[i], a_[i]):
                if o >= 0 and not is_leaf(o):
                    stack.append((o, False))

    def _emit(self, i):
        T = self.T
        o = T.op[i]
        r = self.ref
        a, b, c = T.a[i], T.b[i], T.c[i]
        name = ('v{}' if self.varying_ctx else 'w{}').format(i)
        typ = 'const double'
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
            typ = 'const bool'
            sym = {ir.LT: '<', ir.LE: '<=', ir.EQ: '==', ir.NE: '!=',
                   ir.AND: '&&', ir.OR: '||'}
            if o == ir.NOT:
                e = '!{}'.format(r(a))
            else:
                e = '({} {} {})'.format(r(a), sym[o], r(b))
        else:
            raise NotImplementedError(ir.OP_NAMES[o])
        self.lines.append('{} {} = {};'.format(typ, name, e))
        self.num_ops += 1


def choose_tile_cols(requested):
    """Tile width C (doubles) of the Jacobian staging tile.  C/2 must be odd
    so that the 16-byte shared-memory stores of 8 consecutive lanes (row pitch
    8*C bytes) fall into distinct bank groups."""
    c = max(2, int(requested))
    c -= c % 2
    if (c // 2) % 2 == 0:
        c -= 2
    return max(2, c)


def emit_module(prog, groups, method, tile_cols=30, warps_per_block=2,
                min_blocks_per_sm=4, tma_load=True, tma_store=True,
                derived=(), debug_nostore=False, tile_bufs=2, debug_reps=1,
                const_runs=(), only_groups=None, with_aux=True,
                persistent=False, tile_major=False,
                persistent_block_stores=True):
    """Returns ``(source_text, meta)`` for ``prog`` split into ``groups``
    (list of ``(c0, c1)`` column ranges of the flattened ``M*P`` node block;
    a group also owns the residuals of the rows that start inside it).  ``derived`` lists the tape ids
    that the pre-pass kernel evaluates once per node into derived rows.
    ``const_runs`` lists ``(col0, length)`` column runs of the node block
    whose entries are the same for every node: the group bodies skip them
    and the runtime's replicator kernel copies one shared-memory image of
    them into every node row with TMA tile stores.

    Large problems are compiled as several modules in parallel: with
    ``only_groups = (g0, g1)`` the module contains the bodies and the main
    kernel of groups ``g0 .. g1-1`` only (group and store-segment indices
    inside it are local); ``with_aux=False`` leaves out the invariants and
    pre-pass kernels (the first module carries them).

    ``persistent=True`` emits the persistent variant of the main kernel
    (``csrc/colloc_persistent.cuh``): one block per SM bound to one group,
    warps looping over node tiles along a host-made schedule, the pre-pass
    as phase 0 of the same launch."""
    T = prog.tape
    M, P, K, R = prog.M, prog.P, prog.K, prog.R
    C = choose_tile_cols(tile_cols)
    if warps_per_block > 4 and warps_per_block % 4 and not persistent:
        raise ValueError('warps_per_block above 4 must be a multiple of 4')
    ninv = len(prog.inv_nodes)
    derived = list(derived)
    derived_index = {nid: k for k, nid in enumerate(derived)}
    D = len(derived)

    const_runs = sorted(const_runs)
    carved = [False] * K
    for a, ln in const_runs:
        assert a % 2 == 0 and ln % 2 == 0 and tma_store
        for c in range(a, a + ln):
            carved[c] = True
    # store segments: maximal runs of non-carved columns inside one group;
    # each gets its own TMA descriptor (box C x 32, clipped at the segment end)
    segments = []          # (col0, ncols)
    group_segments = []    # per group: list of segment ids
    for (gc0, gc1) in groups:
        ids = []
        c = gc0
        end = gc1
        while c < end:
            if carved[c]:
                c += 1
                continue
            e = c
            while e < end and not carved[e]:
                e += 1
            ids.append(len(segments))
            segments.append((c, e - c))
            c = e
        group_segments.append(ids)
    seg_of_col = {}
    for sid, (a, ln) in enumerate(segments):
        for c in range(a, a + ln):
            seg_of_col[c] = sid
    ncc = sum(ln for _, ln in const_runs)
    g0, g1 = only_groups if only_groups is not None else (0, len(groups))
    seg_first = group_segments[g0][0] if group_segments[g0] else 0
    local_segs = [sid for g in range(g0, g1) for sid in group_segments[g]]
    assert local_segs == list(range(seg_first, seg_first + len(local_segs)))

    out = []
    w = out.append
    w('// generated by opty_b200.codegen (emitter v{}); do not edit'.format(
        EMITTER_VERSION))
    w('#define OPTY_NSEGS {}'.format(max(len(local_segs), 1)))
    w('#define OPTY_M {}'.format(M))
    w('#define OPTY_P {}'.format(P))
    w('#define OPTY_K {}'.format(K))
    w('#define OPTY_R {}'.format(R))
    w('#define OPTY_D {}'.format(D))
    w('#define OPTY_C {}'.format(C))
    w('#define OPTY_NGROUPS {}'.format(g1 - g0))
    w('#define OPTY_NINV {}'.format(max(ninv, 1)))
    w('#define OPTY_NUNI {}'.format(max(prog.num_uniform, 1)))
    if persistent:
        # the base skeleton is used with its per-warp geometry
        w('#define OPTY_WARPS 1')
        w('#define OPTY_PWARPS {}'.format(warps_per_block))
        if not persistent_block_stores:
            w('#define OPTY_PERSIST_BLOCK_STORES 0')
        w('#define OPTY_MIN_BLOCKS 1')
    else:
        w('#define OPTY_WARPS {}'.format(warps_per_block))
        w('#define OPTY_MIN_BLOCKS {}'.format(min_blocks_per_sm))
    w('#define OPTY_TMA_LOAD {}'.format(int(tma_load)))
    w('#define OPTY_TMA_STORE {}'.format(1 if tma_store else 0))
    w('#define OPTY_NBUF {}'.format(int(tile_bufs)))
    if tile_major:
        w('#define OPTY_TILE_MAJOR 1')
    if debug_nostore:
        w('#define OPTY_DEBUG_NOSTORE {}'.format(int(debug_nostore)))
    if debug_reps != 1:
        w('#define OPTY_DEBUG_REPS {}'.format(int(debug_reps)))
    w('#include "colloc_kernel.cuh"')
    if persistent:
        w('#include "colloc_persistent.cuh"')
    w('')

    # ---- invariants kernel -------------------------------------------
    inv_ops = 0
    if with_aux:
        bw = _BodyWriter(prog, 'inv')
        for nid in prog.inv_nodes:
            bw.need(nid)
        w('extern "C" __global__ void opty_colloc_inv('
          'const double* __restrict__ uni, double* __restrict__ inv)')
        w('{')
        w('  if (threadIdx.x != 0 || blockIdx.x != 0) return;')
        for line in bw.lines:
            w('  ' + line)
        for k, nid in enumerate(prog.inv_nodes):
            w('  inv[{}] = {};'.format(k, bw.ref(nid)))
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
    pre_groups = len(pre_chunks)
    # pattern of the constant runs (replicated into every node row by the
    # runtime's opty_replicate_kernel): literal value, or index into the
    # node-invariant table, in run order
    const_lit, const_inv = [], []
    for a, ln in const_runs:
        for c in range(a, a + ln):
            e = prog.jac[c // P][c % P]
            if T.op[e] == ir.CONST:
                const_lit.append(float(T.val[e]))
                const_inv.append(-1)
            else:
                assert not T.varying[e]
                const_lit.append(0.0)
                const_inv.append(prog.inv_index[e])
    pre_ops = 0
    if with_aux:
        pre_cases = []
        for pg, ks in enumerate(pre_chunks):
            bw = _BodyWriter(prog, 'pre')
            for k in ks:
                bw.need(derived[k])
                bw.lines.append('OPTY_DRV({}, {});'.format(
                    k, bw.ref(derived[k])))
            pre_cases.append('    case {}: {{'.format(pg))
            pre_cases.extend('      ' + line for line in bw.lines)
            pre_cases.append('    } break;')
            pre_ops += bw.num_ops
        if persistent:
            # phase 0 of the persistent kernel calls the same bodies
            w('static __device__ void opty_pre_unit(const OptyParams& p, '
              'const int node, const int opty_pg)')
            w('{')
            w('  const double* xg = p.traj + node;')
            w('  double* drv = p.traj + (long long)OPTY_R * p.ldt + node;')
            w('  switch (opty_pg) {')
            for line in pre_cases:
                w(line)
            w('    default: break;')
            w('  }')
            w('}')
            w('')
        w('extern "C" __global__ void __launch_bounds__(OPTY_PRE_THREADS)')
        w('opty_colloc_pre(const OptyParams p)')
        w('{')
        w('  OPTY_PRE_BEGIN();')
        if persistent:
            w('  opty_pre_unit(p, node, opty_pg);')
            w('  (void)xg; (void)drv;')
        else:
            w('  switch (opty_pg) {')
            for line in pre_cases:
                w(line)
            w('    default: break;')
            w('  }')
        w('}')
        w('')

    # ---- group bodies --------------------------------------------------
    group_meta = []
    for g, (gc0, gc1) in enumerate(groups):
        if not g0 <= g < g1:
            continue
        bw = _BodyWriter(prog, 'main', derived_index)
        body = bw.lines
        w('static __device__ __forceinline__ void opty_group_{}('
          'const OptyCtx& ctx)'.format(g))
        w('{')
        state = {'seg': None, 'cc': 0, 'chunk': 0, 'pending': None,
                 'stored': 0}

        def flush():
            sid = state['seg']
            cc = state['cc']
            ncols_in_chunk = cc % C or C
            q = (cc - 1) // C
            body.append('OPTY_FLUSH({}, {}, {}, {}, {});'.format(
                sid - seg_first, q, state['chunk'] % tile_bufs,
                segments[sid][0], ncols_in_chunk))
            state['chunk'] += 1

        def close_segment():
            assert state['pending'] is None
            if state['seg'] is not None and state['cc'] % C != 0:
                flush()
            state['seg'] = None
            state['cc'] = 0

        for col in range(gc0, gc1):
            j, k = divmod(col, P)
            if k == 0 and prog.con:
                bw.need(prog.con[j])
                body.append('OPTY_CON({}, {});'.format(
                    j, bw.ref(prog.con[j])))
            if carved[col]:
                continue
            sid = seg_of_col[col]
            if sid != state['seg']:
                close_segment()
                state['seg'] = sid
            seg_ncols = segments[sid][1]
            e = prog.jac[j][k]
            bw.need(e)
            cc = state['cc']
            tc = cc % C
            buf = state['chunk'] % tile_bufs
            pending = state['pending']
            if pending is None and tc % 2 == 0 and tc + 1 < C and \
                    cc + 1 < seg_ncols:
                state['pending'] = (tc, bw.ref(e))
            elif pending is not None:
                body.append('OPTY_JS2({}, {}, {}, {});'.format(
                    buf, pending[0], pending[1], bw.ref(e)))
                state['pending'] = None
            else:
                body.append('OPTY_JS1({}, {}, {});'.format(
                    buf, tc, bw.ref(e)))
            state['cc'] = cc + 1
            state['stored'] += 1
            if state['cc'] % C == 0 and state['pending'] is None:
                flush()
        close_segment()
        body.append('OPTY_DRAIN();')
        for line in body:
            w('  ' + line)
        w('}')
        w('')
        group_meta.append({'rows': [gc0 // P, -(-gc1 // P)],
                           'cols': [gc0, gc1], 'col0': gc0,
                           'ncols': state['stored'],
                           'segments': group_segments[g],
                           'ops': bw.num_ops, 'chunks': state['chunk']})

    # blockIdx.y -> group: most expensive groups are launched first
    order = sorted(range(len(group_meta)),
                   key=lambda g: -(20 * group_meta[g]['ops'] +
                                   43 * group_meta[g]['ncols']))
    w('__device__ const int opty_group_order[OPTY_NGROUPS] = {{{}}};'.format(
        ', '.join(str(g) for g in order)))
    w('')
    if persistent:
        w('extern "C" __global__ void __launch_bounds__(OPTY_PTHREADS, 1)')
        w('opty_colloc_eval(const __grid_constant__ OptyTmaps tm, '
          'const OptyParams p, const OptyPersist ps)')
        w('{')
        w('  OPTY_PERSIST_BEGIN()')
        w('  OPTY_PERSIST_LOOP_BEGIN()')
        w('  switch (opty_g) {')
        for g in range(g0, g1):
            w('    case {}: opty_group_{}(ctx); break;'.format(g - g0, g))
        w('    default: break;')
        w('  }')
        w('  OPTY_PERSIST_LOOP_END()')
        w('  OPTY_PERSIST_END()')
        w('}')
        w('')
    else:
        w('extern "C" __global__ void __launch_bounds__(OPTY_THREADS, '
          'OPTY_MIN_BLOCKS)')
        w('opty_colloc_eval(const __grid_constant__ OptyTmaps tm, '
          'const OptyParams p)')
        w('{')
        w('  OPTY_KERNEL_BEGIN()')
        if debug_reps != 1:
            # measurement aid: the same tile is evaluated several times,
            # passes after the first find the group body in the instruction
            # caches
            w('#pragma unroll 1')
            w('  for (int opty_rep = 0; opty_rep < OPTY_DEBUG_REPS; '
              '++opty_rep)')
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
        'segment_range': [seg_first, seg_first + len(local_segs)],
        'segments': [list(sg) for sg in segments],
        'const_runs': [list(cr) for cr in const_runs],
        'const_image_doubles': ncc,
        'const_lit': const_lit,
        'const_inv': const_inv,
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
        'tile_bufs': int(tile_bufs),
        'persistent': bool(persistent),
        'persistent_block_stores': bool(persistent_block_stores),
        'method': method,
        'entry_kind': prog.entry_kind(),
        'stats': prog.stats(),
    }
    return '\n'.join(out), meta
