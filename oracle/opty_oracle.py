"""CPU oracle for the collocation constraint + Jacobian path of csu-hmc/opty.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module, and only as the checker / the CPU baseline -- never as part of the
product path (``opty_b200/`` does not import it and has no CPU fallback).

It is a restatement, from the reference's sources, of what opty does between a
SymPy equations-of-motion matrix and the numbers IPOPT sees:

    symbol sorting        opty/direct_collocation.py:1904-2035
    discrete symbols      opty/direct_collocation.py:2037-2118
    discretisation        opty/direct_collocation.py:2120-2156
    argument / wrt order  opty/direct_collocation.py:2345-2364, 2713-2747
    forward Jacobian      opty/utils.py:82-228
    C code generation     opty/utils.py:61-79, 483-494, 743-757
    node loop             opty/utils.py:500-529 (Cython there, plain C here)
    free-vector parsing   opty/utils.py:277-326
    fixed/free merging    opty/direct_collocation.py:2891-2926
    callback wrappers     opty/direct_collocation.py:2382-2446, 2816-2887,
                          2928-3001
    COO index loop        opty/direct_collocation.py:2450-2690
    instance constraints  opty/direct_collocation.py:2158-2282

The generated C is compiled with ``gcc -O2`` (the reference's effective flags;
``-fopenmp`` is added for ``parallel=True`` exactly as opty/utils.py:735-736
does) and driven through ctypes instead of Cython.

Parity is PINNED: ``tests/test_oracle.py`` checks this module against golden
vectors that ``tests/golden/make_golden.py`` produced by running the reference
itself (``/root/reference``, backend='cython') in the build container, and
against the reference's hand-computed known-answer tests
(opty/tests/test_direct_collocation.py:791-866, 1127-1177, ...).  On BASELINE
config 2 the oracle's residuals and Jacobian are bit-identical to the
reference's.
"""

import ctypes
import hashlib
import os
import subprocess
from collections import Counter

import numpy as np
import sympy as sm
import sympy.physics.mechanics as me
from sympy.printing.c import C99CodePrinter
from sympy.utilities.iterables import numbered_symbols

_HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(_HERE, '_build')


# ---------------------------------------------------------------------------
# SymPy -> C text  (opty/utils.py:61-79)
# ---------------------------------------------------------------------------
class _UnderscorePrinter(C99CodePrinter):
    """Every symbol / function name gets a trailing underscore."""

    def _print_Symbol(self, expr):
        return super()._print_Symbol(expr) + '_'

    def _print_Function(self, expr):
        return super()._print_Function(expr) + '_'


def _c(expr, assign_to=None):
    return _UnderscorePrinter({}).doprint(expr, assign_to)


# ---------------------------------------------------------------------------
# forward-mode symbolic Jacobian over the CSE graph  (opty/utils.py:82-228)
# ---------------------------------------------------------------------------
class _ForwardJacobian(object):
    """Restatement of ``opty.utils._forward_jacobian``.

    State: ``sub_of`` maps an expression to the symbol that replaces it,
    ``expr_of`` maps a replacement symbol to its (reduced) expression.
    """

    def __init__(self, expr, wrt):
        if not isinstance(expr, sm.ImmutableDenseMatrix) or expr.shape[1] != 1:
            raise NotImplementedError('column ImmutableDenseMatrix required')
        if not isinstance(wrt, sm.ImmutableDenseMatrix) or wrt.shape[1] != 1:
            raise NotImplementedError('column ImmutableDenseMatrix required')
        self.expr = expr
        self.wrt = wrt
        self.symbols = numbered_symbols(prefix='z', cls=sm.Symbol,
                                        exclude=expr.free_symbols, real=True)
        self.sub_of = {}
        self.expr_of = {}

    def register(self, node):
        """Returns ``(symbol, reduced expression)`` for ``node``, creating a
        new replacement when the node has not been seen (utils.py:89-105)."""
        if node in self.sub_of:
            s = self.sub_of[node]
            return s, self.expr_of[s]
        if node in self.expr_of:
            return node, self.expr_of[node]
        if isinstance(node, sm.Tuple):
            return None, None
        if not node.free_symbols:
            return node, node
        s = next(self.symbols)
        reduced = node.xreplace(self.sub_of)
        self.expr_of[s] = reduced
        self.sub_of[node] = s
        return s, reduced

    def run(self):
        wrt_list = list(self.wrt.args[2])
        P = len(wrt_list)

        # 1. CSE of the expression; every sub-node of every replacement gets
        #    its own symbol (utils.py:141-151)
        repl, reduced = sm.cse(self.expr.args[2], self.symbols, order='none')
        for s, sub in repl:
            self.expr_of[s] = sub.xreplace(self.sub_of)
            self.sub_of[sub] = s
            for node in sm.postorder_traversal(sub):
                self.register(node)
        for red in reduced:
            for node in red:
                self.register(node)
        reduced_matrix = sm.ImmutableDenseMatrix(reduced).xreplace(self.sub_of)
        ordered = list(self.expr_of.items())

        # 2. seeds (utils.py:159-165)
        total = {}
        for i, w in enumerate(wrt_list):
            row = [sm.S.Zero] * P
            row[i] = sm.S.One
            total[w] = sm.ImmutableDenseMatrix([row])

        # 3. chain rule in definition order (utils.py:167-185)
        zero_row = sm.ImmutableDenseMatrix.zeros(1, P)
        for s, sub in ordered:
            acc = zero_row
            for fs in sm.ordered(sub.free_symbols):
                _, partial = self.register(sub.diff(fs))
                acc += partial * total.get(fs, zero_row)
            total[s] = sm.ImmutableDenseMatrix(
                [[self.register(a)[0] for a in acc]])
        jac = sm.ImmutableDenseMatrix.vstack(
            *[total[e] for e in reduced_matrix])

        # 4. keep only what the Jacobian entries need (utils.py:187-209)
        needed = set()
        stack = [e for e in jac if e.free_symbols]
        while stack:
            e = stack.pop()
            if e in needed or e in self.wrt:
                continue
            kids = list(sm.ordered(self.expr_of.get(e, e).free_symbols))
            for kid in kids:
                if kid not in needed and kid not in self.wrt:
                    stack.append(kid)
            needed.add(e)
        dense = {s: sub for s, sub in self.expr_of.items() if s in needed}

        # 5. inline plain symbols and single-use temporaries (utils.py:211-223)
        uses = Counter(sm.ordered(jac.free_symbols))
        for sub in dense.values():
            uses.update(sm.ordered(sub.free_symbols))
        kept = {}
        inlined = {}
        for s, sub in dense.items():
            if isinstance(sub, sm.Symbol) or uses[s] == 1:
                inlined[s] = sub.xreplace(inlined)
            else:
                kept[s] = sub.xreplace(inlined)
        return list(kept.items()), [jac.xreplace(inlined)]


def forward_jacobian(expr, wrt):
    return _ForwardJacobian(expr, wrt).run()


# ---------------------------------------------------------------------------
# C module generation + loading  (opty/utils.py:483-529, 743-757)
# ---------------------------------------------------------------------------
_C_TEMPLATE = """\
// oracle_code_hash={code_hash}
#include <math.h>

static void eval_matrix(double matrix[{size}],
{scalar_args})
{{
{body}
}}

void eval_matrix_loop(long n, double* matrix, const double* const* arrays,
                      const double* consts)
{{
    long i;
    {pragma}
    for (i = 0; i < n; i++) {{
        eval_matrix(matrix + i*{size},
{call_args});
    }}
}}
"""


def _gcc():
    # the image's default $CC wrapper cannot link -fopenmp
    return '/usr/bin/gcc' if os.path.exists('/usr/bin/gcc') else 'gcc'


def compile_matrix_function(args, expr, const=(), parallel=False,
                            build_dir=None):
    """Restatement of ``ufuncify_matrix`` (opty/utils.py:639-928).

    ``expr`` is a Matrix or the ``(replacements, [matrix])`` pair that cse /
    forward_jacobian return.  Returns ``f(matrix, *args)`` with the
    reference's calling convention: ``matrix`` is a C-contiguous float64
    ``(n, rows*cols)`` array that is filled and returned reshaped to ``(n,
    rows, cols)``; non-const args are float64 arrays of length n, const args
    floats.
    """
    if hasattr(expr, 'shape'):
        rows, cols = expr.shape
        sub_exprs, mats = sm.cse(expr, sm.numbered_symbols('z_'),
                                 order='none')
    else:
        sub_exprs, mats = expr
        rows, cols = mats[0].shape
    size = rows * cols
    lines = ['double ' + _c(sub, sym) for sym, sub in sub_exprs]
    lines.append(_c(mats[0], sm.MatrixSymbol('matrix', rows, cols)))
    body = '    ' + '\n    '.join('\n'.join(lines).split('\n'))

    const = tuple(const)
    arr_idx = 0
    con_idx = 0
    call = []
    for a in args:
        if a in const:
            call.append('consts[{}]'.format(con_idx))
            con_idx += 1
        else:
            call.append('arrays[{}][i]'.format(arr_idx))
            arr_idx += 1
    pad = ' ' * 20
    code_hash = hashlib.sha256(
        ('const={}parallel={}'.format(const, parallel) + body).encode()
    ).hexdigest()
    src = _C_TEMPLATE.format(
        code_hash=code_hash, size=size,
        scalar_args=',\n'.join(pad + 'double ' + _c(a) for a in args),
        body=body,
        pragma='#pragma omp parallel for' if parallel else '',
        call_args=',\n'.join(pad + c for c in call))

    build_dir = build_dir or BUILD_DIR
    os.makedirs(build_dir, exist_ok=True)
    stem = os.path.join(build_dir, 'oracle_' + code_hash[:24])
    so_path = stem + '.so'
    if not os.path.exists(so_path):
        with open(stem + '.c', 'w') as f:
            f.write(src)
        cmd = [_gcc(), '-O2', '-fPIC', '-shared', '-o', so_path + '.tmp',
               stem + '.c', '-lm']
        if parallel:
            cmd.insert(1, '-fopenmp')
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise ImportError('oracle C compilation failed:\n' + proc.stderr)
        os.replace(so_path + '.tmp', so_path)
    lib = ctypes.CDLL(so_path)
    dptr = ctypes.POINTER(ctypes.c_double)
    lib.eval_matrix_loop.argtypes = [ctypes.c_long, dptr,
                                     ctypes.POINTER(dptr), dptr]
    lib.eval_matrix_loop.restype = None
    is_const = [a in const for a in args]

    def loop(matrix, *values):
        assert matrix.flags['C_CONTIGUOUS'] and matrix.dtype == np.float64
        n = matrix.shape[0]
        arrays = []
        consts = []
        for flag, v in zip(is_const, values):
            if flag:
                consts.append(float(v))
            else:
                v = np.ascontiguousarray(v, dtype=np.float64)
                assert v.shape == (n,)
                arrays.append(v)
        arr_ptrs = (dptr * max(len(arrays), 1))(
            *[a.ctypes.data_as(dptr) for a in arrays])
        cvals = np.array(consts if consts else [0.0])
        lib.eval_matrix_loop(n, matrix.ctypes.data_as(dptr), arr_ptrs,
                             cvals.ctypes.data_as(dptr))
        return matrix.reshape(n, rows, cols)

    loop.source = src
    loop.source_path = stem + '.c'
    return loop


# ---------------------------------------------------------------------------
# free-vector parsing  (opty/utils.py:277-326)
# ---------------------------------------------------------------------------
def parse_free(free, n, q, N, variable_duration=False):
    states = free[:n * N].reshape((n, N))
    specified = None
    if q:
        specified = free[n * N:(n + q) * N]
        if q > 1:
            specified = specified.reshape((q, N))
    if variable_duration:
        return states, specified, free[(n + q) * N:-1], free[-1]
    return states, specified, free[(n + q) * N:]


def _sorted_by_name(seq):
    # opty/utils.py:473-480
    seq = list(seq)
    try:
        seq.sort(key=lambda x: x.name)
    except AttributeError:
        seq.sort(key=lambda x: x.__class__.__name__)
    return seq


class OracleCollocator(object):
    """CPU oracle with the reference ConstraintCollocator's constructor
    arguments (opty/direct_collocation.py:1406-1411) and the three products
    of the hot path: ``constraints(free)``, ``jacobian(free)``,
    ``jacobian_indices()``."""

    def __init__(self, equations_of_motion, state_symbols,
                 num_collocation_nodes, node_time_interval,
                 known_parameter_map={}, known_trajectory_map={},
                 instance_constraints=None, time_symbol=None,
                 integration_method='backward euler', parallel=False,
                 build_dir=None):
        self.eom = equations_of_motion
        # the reference redirects the global default time symbol
        # (direct_collocation.py:1490-1494); find_dynamicsymbols needs it
        if time_symbol is not None:
            me.dynamicsymbols._t = time_symbol
        self.t = me.dynamicsymbols._t
        self.states = tuple(state_symbols)
        self.state_derivs = tuple(s.diff(self.t) for s in self.states)
        self.N = num_collocation_nodes
        self.node_time_interval = node_time_interval
        self.variable_duration = isinstance(node_time_interval, sm.Symbol)
        self.h_sym = node_time_interval if self.variable_duration else \
            sm.Symbol('h_opty', real=True)
        self.par_map = known_parameter_map
        self.traj_map = known_trajectory_map
        self.instance_constraints = instance_constraints
        self.method = integration_method
        self.parallel = parallel
        self.build_dir = build_dir
        self.M = self.eom.shape[0]
        self.n = len(self.states)

        # parameters: known in the given order, unknown sorted by name
        # (direct_collocation.py:1954-1973)
        pars = set(self.eom.free_symbols)
        pars.discard(self.t)
        self.known_pars = tuple(self.par_map.keys())
        self.unknown_pars = tuple(
            _sorted_by_name(pars.difference(self.known_pars)))
        self.pars = self.known_pars + self.unknown_pars
        self.r = len(self.unknown_pars)

        # trajectories (direct_collocation.py:1988-2035)
        non_states = me.find_dynamicsymbols(self.eom).difference(
            set(self.states) | set(self.state_derivs))
        self.known_trajs = tuple(self.traj_map.keys())
        self.unknown_trajs = tuple(
            _sorted_by_name(non_states.difference(self.known_trajs)))
        self.trajs = self.known_trajs + self.unknown_trajs
        self.q = len(self.unknown_trajs)

        self.num_free = ((self.n + self.q) * self.N + self.r +
                         int(self.variable_duration))

        # discrete symbols (direct_collocation.py:2070-2118)
        def tag(funcs, suffix):
            return tuple(sm.Symbol(f.__class__.__name__ + suffix, real=True)
                         for f in funcs)
        self.xp, self.xi, self.xn = (tag(self.states, s) for s in 'pin')
        self.si, self.sn = tag(self.trajs, 'i'), tag(self.trajs, 'n')
        self.ui = tag(self.unknown_trajs, 'i')
        self.un = tag(self.unknown_trajs, 'n')

        # discretisation (direct_collocation.py:2143-2156)
        h = self.h_sym
        if self.method == 'backward euler':
            d_sub = {d: (i - p) / h
                     for d, i, p in zip(self.state_derivs, self.xi, self.xp)}
            f_sub = dict(zip(self.states + self.trajs, self.xi + self.si))
            self.discrete_eom = me.msubs(self.eom, d_sub, f_sub)
        elif self.method == 'midpoint':
            d_sub = {d: (n - i) / h
                     for d, i, n in zip(self.state_derivs, self.xi, self.xn)}
            x_sub = {d: (i + n) / 2
                     for d, i, n in zip(self.states, self.xi, self.xn)}
            u_sub = {d: (i + n) / 2
                     for d, i, n in zip(self.trajs, self.si, self.sn)}
            self.discrete_eom = me.msubs(self.eom, d_sub, x_sub, u_sub)
        else:
            raise ValueError(self.method)

        if self.method == 'backward euler':
            self.P = 2 * self.n + self.q + self.r + int(self.variable_duration)
        else:
            self.P = (2 * self.n + 2 * self.q + self.r +
                      int(self.variable_duration))

        self.o = 0
        if instance_constraints is not None:
            self.o = len(instance_constraints)
            self._prepare_instance_constraints()
        self._con_loop = None
        self._jac_loop = None

    # -- argument lists (direct_collocation.py:2345-2364, 2713-2747) --------
    def _args(self):
        if self.method == 'backward euler':
            return (self.xi + self.xp + self.si + self.pars + (self.h_sym,))
        return (self.xi + self.xn + self.si + self.sn + self.pars +
                (self.h_sym,))

    def _wrt(self):
        if self.method == 'backward euler':
            wrt = self.xi + self.xp + self.ui + self.unknown_pars
        else:
            wrt = self.xi + self.xn + self.ui + self.un + self.unknown_pars
        if self.variable_duration:
            wrt += (self.h_sym,)
        return wrt

    def _constraint_loop(self):
        if self._con_loop is None:
            self._con_loop = compile_matrix_function(
                self._args(), self.discrete_eom,
                const=self.pars + (self.h_sym,), parallel=self.parallel,
                build_dir=self.build_dir)
        return self._con_loop

    def _jacobian_loop(self):
        if self._jac_loop is None:
            partials = forward_jacobian(
                sm.ImmutableDenseMatrix(self.discrete_eom),
                sm.ImmutableDenseMatrix([list(self._wrt())]).T)
            self._jac_loop = compile_matrix_function(
                self._args(), partials, const=self.pars + (self.h_sym,),
                parallel=self.parallel, build_dir=self.build_dir)
            self._jac_buffer = np.empty((self.N - 1, self.M * self.P))
        return self._jac_loop

    def _loops(self):
        return self._constraint_loop(), self._jacobian_loop()

    # -- numeric argument assembly (direct_collocation.py:2411-2437) --------
    def _numeric_args(self, free):
        N = self.N
        if self.variable_duration:
            x, u, p_free, h = parse_free(free, self.n, self.q, N, True)
        else:
            x, u, p_free = parse_free(free, self.n, self.q, N)
            h = self.node_time_interval
        # known first, then unknown, as in _merge_fixed_free
        # (direct_collocation.py:2911-2926)
        traj_vals = []
        k = 0
        for s in self.trajs:
            if s in self.traj_map:
                v = self.traj_map[s]
                traj_vals.append(v(free) if callable(v) else np.asarray(v))
            else:
                traj_vals.append(u if u.ndim == 1 else u[k])
                k += 1
        par_vals = [float(self.par_map[s]) if s in self.par_map else
                    float(p_free[self.unknown_pars.index(s)])
                    for s in self.pars]
        if self.method == 'backward euler':
            cur, adj = slice(1, None), slice(None, -1)
        else:
            cur, adj = slice(None, -1), slice(1, None)
        vals = [row[cur] for row in x] + [row[adj] for row in x]
        vals += [tv[cur] for tv in traj_vals]
        if self.method == 'midpoint':
            vals += [tv[adj] for tv in traj_vals]
        vals = [np.ascontiguousarray(v) for v in vals]
        return vals + par_vals + [float(h)]

    def constraints(self, free):
        """``[eom_1 @ nodes, ..., eom_M @ nodes, c_1..c_o]``
        (direct_collocation.py:2444-2446, 2985-2991)."""
        free = np.asarray(free, dtype=float)
        con_loop = self._constraint_loop()
        result = np.empty((self.N - 1, self.M))
        vals = con_loop(result, *self._numeric_args(free))
        out = vals.reshape(self.N - 1, self.M).T.flatten()
        if self.o:
            out = np.hstack((out, self._instance_values(free)))
        return out

    def jacobian(self, free, copy=True):
        """Node-major partials then the instance entries
        (direct_collocation.py:2885-2887, 2985-2991).  With ``copy=False``
        and no instance constraints the result is a view of the persistent
        buffer, overwritten by the next call -- exactly what the reference
        returns (direct_collocation.py:2814, 2887); the default copies so
        that tests can hold several results."""
        free = np.asarray(free, dtype=float)
        jac_loop = self._jacobian_loop()
        vals = jac_loop(self._jac_buffer, *self._numeric_args(free)).ravel()
        if self.o:
            return np.hstack((vals, self._instance_jacobian_values(free)))
        return vals.copy() if copy else vals

    # -- COO structure: the reference's per-node loop -------------------------
    def jacobian_indices(self):
        """direct_collocation.py:2628-2688."""
        N, M, n, q = self.N, self.M, self.n, self.q
        tail = self.r + int(self.variable_duration)
        per_node = M * self.P
        nnz = (N - 1) * per_node
        if self.o:
            irows, icols = self._instance_indices()
            nnz += len(irows)
        rows = np.empty(nnz, dtype=int)
        cols = np.empty(nnz, dtype=int)
        for i in range(N - 1):
            r_idx = [j * (N - 1) + i for j in range(M)]
            if self.method == 'backward euler':
                c_idx = [j * N + i + 1 for j in range(n)]
                c_idx += [j * N + i for j in range(n)]
                c_idx += [n * N + j * N + i + 1 for j in range(q)]
            else:
                c_idx = [j * N + i for j in range(n)]
                c_idx += [j * N + i + 1 for j in range(n)]
                c_idx += [n * N + j * N + i for j in range(q)]
                c_idx += [n * N + j * N + i + 1 for j in range(q)]
            c_idx += [(n + q) * N + j for j in range(tail)]
            rows[i * per_node:(i + 1) * per_node] = np.repeat(r_idx,
                                                              len(c_idx))
            cols[i * per_node:(i + 1) * per_node] = np.tile(c_idx, M)
        if self.o:
            rows[-len(irows):] = irows
            cols[-len(icols):] = icols
        return rows, cols

    # -- instance constraints (direct_collocation.py:2158-2282) --------------
    def _prepare_instance_constraints(self):
        N = self.N
        funcs = set()
        for con in self.instance_constraints:
            funcs |= con.atoms(sm.Function)
        index = {}
        for f in funcs:
            if self.variable_duration:
                node = 0 if f.args[0] == 0 else int(f.args[0] / self.h_sym)
                if node not in range(N):
                    raise ValueError('instance time outside the node range')
            else:
                grid = np.linspace(0.0, self.node_time_interval * (N - 1),
                                   num=N)
                node = np.argmin(np.abs(grid - float(f.args[0])))
            base = f.__class__(self.t)
            if base in self.states:
                index[f] = node + self.states.index(base) * N
            elif base in self.unknown_trajs:
                index[f] = (node + self.n * N +
                            self.unknown_trajs.index(base) * N)
        self.instance_index = index
        vec = sm.DeferredVector('FREE')
        sub = {f: vec[i] for f, i in index.items()}
        known = list(self.par_map.keys())
        mods = [{'ImmutableMatrix': np.array}, 'numpy']
        self._inst_f = sm.lambdify(
            [vec] + known, [c.subs(sub) for c in self.instance_constraints],
            modules=mods)
        self._inst_jac = []
        for con in self.instance_constraints:
            wrt = list(con.atoms(sm.Function))
            jac = sm.Matrix([con]).jacobian(wrt).subs(sub)
            self._inst_jac.append((len(wrt),
                                   sm.lambdify([vec] + known, jac,
                                               modules=mods)))

    def _instance_values(self, free):
        return self._inst_f(free, *self.par_map.values())

    def _instance_jacobian_values(self, free):
        out = np.zeros(sum(c for c, _ in self._inst_jac))
        j = 0
        for count, f in self._inst_jac:
            out[j:j + count] = f(free, *self.par_map.values())
            j += count
        return out

    def _instance_indices(self):
        base = self.M * (self.N - 1)
        rows, cols = [], []
        for i, con in enumerate(self.instance_constraints):
            for f in con.atoms(sm.Function):
                rows.append(base + i)
                cols.append(self.instance_index[f])
        return np.array(rows, dtype=int), np.array(cols, dtype=int)
