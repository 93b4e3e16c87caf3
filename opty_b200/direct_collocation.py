"""``Problem`` / ``ConstraintCollocator`` with the constructor signatures and
the cyipopt callback surface of ``opty.direct_collocation`` (reference
opty/direct_collocation.py:93-567 and :1379-3015), evaluated by generated
sm_100a CUDA kernels instead of generated C + Cython.

Only the collocation hot path is implemented natively here: symbolic
transcription -> CUDA-C emitter -> kernels for the constraint residuals and the
Jacobian partials -> COO structure.  Plotting and initial-guess helpers of the
reference's ``Problem`` are outside the scope of this package.
"""

import logging
import os

import numpy as np
import sympy as sm
import sympy.physics.mechanics as me
from sympy.core.function import AppliedUndef

from . import build, codegen, runtime
from .program import CollocationProgram
from .utils import parse_free, sort_sympy

try:  # IPOPT is optional: every callback works without it, only solve() needs it
    import cyipopt
    _IpoptBase = cyipopt.Problem
except ImportError:  # pragma: no cover - depends on the environment
    cyipopt = None

    class _IpoptBase(object):
        """Stand-in base class used when cyipopt is not installed."""

        def __init__(self, n=None, m=None, lb=None, ub=None, cl=None, cu=None):
            self._nlp_dims = (n, m)
            self._nlp_options = {}

        def add_option(self, name, value):
            self._nlp_options[name] = value

        def solve(self, *args, **kwargs):
            raise ImportError('cyipopt (IPOPT) is not installed; the NLP '
                              'callbacks are available but solve() is not.')

__all__ = ['Problem', 'ConstraintCollocator']

logger = logging.getLogger(__name__)

_METHODS = ('backward euler', 'midpoint')

# longest chain of store phases one output group may hold (prepare_program_module)
MAX_PHASES_PER_GROUP = 48

DEFAULT_CUDA_OPTIONS = {
    'groups': 'auto',           # number of output groups (grid.y) or 'auto'
    'tile_cols': 'auto',        # columns of one staging buffer: a whole
                                # equation row up to 64 columns, else 52
    'tile_bufs': 1,             # staging buffers per warp (1 or 2; at config 2
                                # a second buffer costs a resident block per
                                # SM and 5 us, profiles/r02a_*)
    'warps_per_block': 'auto',  # 2, or 8 when there are enough node tiles x
                                # groups for >= 6 waves of 8-warp blocks (the
                                # warps of a block share every instruction
                                # fetch: large models are bound by that)
    'min_blocks_per_sm': 'auto',  # launch bound: 16 warps per SM when the
                                # scheduler is on (128 registers), 8 otherwise
    'fmad': True,               # FMA contraction (False: mul/add stay unfused
                                # like gcc -O2 on x86-64)
    'maxrregcount': None,
    'tma_load': True,           # input staging: True = TMA tile loads into
                                # shared memory, False = plain loads into
                                # shared memory, 'direct' = no staging
    'tma_store': True,
    'persistent': 'auto',       # 'auto': 'stationary' where it applies (below),
                                # else the grid kernel (False).
                                # True: code-stationary persistent main kernel:
                                # the resident blocks of an SM keep one group
                                # for as long as it has node tiles left (the
                                # body stays in the instruction cache) and pull
                                # tiles from per-group atomic counters.
                                # 'stationary': one block per SM, one equation
                                # row per group, static schedule that keeps a
                                # row on the same SMs, next item's input
                                # prefetched (csrc/colloc_kernel.cuh)
    'pre_pass': True,           # shared expensive sub-expressions once per node
    'schedule': 'auto',         # register-pressure scheduler (schedule.py):
                                # 'auto' = for groups whose plain order keeps
                                # more than 120 values alive; False: outputs
                                # in column order, temporaries depth-first
                                # before their first use
    'reassociate': True,        # sums accumulate their terms in arrival order
                                # (False: the association order of the
                                # reference's C printer is kept, results are
                                # then independent of grouping and tiling)
    'live_budget': 56,          # float64 values a body may keep alive before
                                # the scheduler starts recomputing cheap ones
    'inline_cost': 2,
    'remat_cost': 24,
    'volatile_loads': 'auto',   # input loads nvcc may not merge (large bodies)
    'load_ahead': 0,            # input loads this many statements early
    'fence_every': 0,           # warp-level memory fence every so many
                                # statements (bounds ptxas' load hoisting)
    'debug_nostore': False,     # measurement aid: skip Jacobian tile stores
    'out_ring': 2,              # device output sets to rotate through (2: an
                                # evaluation at a new point does not wait for
                                # the speculative Jacobian copy of the last one)
    'const_rows': True,         # equations whose partials are all
                                # node-invariant are written as constant
                                # column runs, not as groups (True: with the
                                # row-stationary kernel; 'grid': with the grid
                                # kernel too)
    'fused_pre': False,         # row-stationary kernel: derived rows and the
                                # residuals of constant rows as phase 0 of
                                # the main kernel instead of a pre-pass launch
                                # (slower: 256 threads per SM cannot hide the
                                # latency of the sine / cosine chains,
                                # profiles/r02y_*)
    'tile_major': 'auto',       # grid kernel: consecutive blocks of a launch
                                # are the groups of one node tile, so that the
                                # pieces of a node's Jacobian row are written
                                # at about the same time ('auto': with 8-warp
                                # blocks -- 50-link chain 7.5 -> 6.6 ms,
                                # config-4 stand-in 46.1 -> 42.7 us; with
                                # 2-warp blocks an SM would host four
                                # different bodies at a time: 26.9 -> 34.4 us
                                # at the 10-link pendulum)
    'const_kernel': False,      # grid kernel with constant rows: a separate
                                # kernel writes the runs (else the blocks of
                                # the main kernel do, with their tiles)
    'const_pre_pct': 0,         # per cent of the nodes whose constant runs the
                                # pre-pass kernel writes (plain stores)
    'strided_schedule': True,   # row-stationary schedule: whole slots per
                                # group, node tiles dealt round robin to them
                                # (all groups walk the tiles at the same pace)
    'item_cost': 16000,         # row-stationary schedule: fixed cost of an item
                                # in units of 1/20 operation
    'store_hint': 1,            # L2 policy of the Jacobian stores: 0 default,
                                # 1 evict_first (the output is read by nobody
                                # on the device; code and trajectory stay in
                                # L2), 2 evict_last
    'const_head_pct': (35, 50, 15, 70),  # constant runs of the row-stationary
                                # kernel: per cent of its early nodes a warp
                                # sends when the block starts / when its first
                                # input has arrived / with every later item
                                # (the rest after its last item); nodes from
                                # the 4th figure (per cent of all nodes) on
                                # are sent together with the tiles of their
                                # node tile instead
    'use_sympy_cse': True,
    'd2h_skip_constants': True,  # do not re-copy literal Jacobian columns
    'prefetch_jacobian': True,  # constraints() starts the Jacobian D2H early
    'use_index': True,          # set-up cache keyed by the symbolic inputs
    'compile_shards': 'auto',   # modules compiled in parallel (large problems)
    'target_warps': 148 * 16,
    'max_group_cost': 6000.0,
}


def _is_callable_value(v):
    return callable(v) and not isinstance(v, np.ndarray)


class ConstraintCollocator(object):
    """Generates the constraint function, its sparse Jacobian and the
    Jacobian's COO structure for a direct collocation NLP.

    Notation (reference opty/direct_collocation.py:1385-1398): N nodes, M
    equations of motion, n states, q unknown and k known input trajectories,
    r unknown parameters, s = 1 for a free node time interval, o instance
    constraints.  ``free`` has ``n*N + q*N + r + s`` entries, there are
    ``M*(N-1) + o`` constraints.

    The constructor takes the reference's arguments in the reference's order
    (opty/direct_collocation.py:1406-1411); ``parallel`` is accepted and
    ignored (the CUDA grid is the node loop), ``tmp_dir`` selects the
    directory of the compiled-module cache.

    Additional keyword arguments
    ----------------------------
    backend : 'cuda'
    device : int, CUDA device ordinal (default ``LOCAL_RANK`` or 0)
    devices : sequence of CUDA device ordinals (an ordinal may repeat: two
        shards then share that GPU).  The constraint nodes are split into one
        contiguous shard per entry, all driven by this one process; every shard copies its residuals and Jacobian block straight
        into its slice of ONE pinned host vector, so that ``constraints`` /
        ``jacobian`` return the whole problem's vectors like a single-device
        collocator does.
    node_range : (lo, hi), evaluate only constraint nodes ``lo <= i < hi`` of
        the ``N - 1`` (used for sharding the nodes over several processes)
    cuda_options : dict overriding ``DEFAULT_CUDA_OPTIONS``
    """

    def __init__(self, equations_of_motion, state_symbols,
                 num_collocation_nodes, node_time_interval,
                 known_parameter_map={}, known_trajectory_map={},
                 instance_constraints=None, time_symbol=None, tmp_dir=None,
                 integration_method='backward euler', parallel=False,
                 show_compile_output=False, backend='cuda', device=None,
                 node_range=None, cuda_options=None, devices=None):
        self._eom = equations_of_motion

        # the reference also redirects the global default time symbol
        # (opty/direct_collocation.py:1490-1494)
        if time_symbol is None:
            self._time_symbol = me.dynamicsymbols._t
        else:
            self._time_symbol = time_symbol
            me.dynamicsymbols._t = time_symbol

        self._state_symbols = tuple(state_symbols)
        if len(set(self._state_symbols)) != len(self._state_symbols):
            raise ValueError('State symbols must be unique.')

        if backend in ('cython', 'numpy'):
            raise NotImplementedError(
                'opty_b200 only ships the "cuda" backend; the "{}" backend '
                'belongs to the reference implementation.'.format(backend))
        if backend != 'cuda':
            raise ValueError('backend must be "cuda".')
        self._backend = backend

        self._state_derivative_symbols = tuple(
            s.diff(self._time_symbol) for s in self._state_symbols)
        self._num_collocation_nodes = int(num_collocation_nodes)

        self._node_time_interval = node_time_interval
        if isinstance(node_time_interval, sm.Symbol):
            self._variable_duration = True
            self._time_interval_symbol = node_time_interval
        else:
            self._variable_duration = False
            self._time_interval_symbol = sm.Symbol('h_opty', real=True)

        self._known_parameter_map = known_parameter_map
        self._known_trajectory_map = known_trajectory_map
        self._instance_constraints = instance_constraints
        self._tmp_dir = tmp_dir
        self._parallel = parallel
        self._show_compile_output = show_compile_output

        opts = dict(DEFAULT_CUDA_OPTIONS)
        if cuda_options:
            unknown = set(cuda_options) - set(opts)
            if unknown:
                raise ValueError('Unknown cuda_options: {}'.format(
                    sorted(unknown)))
            opts.update(cuda_options)
        self._cuda_options = opts
        self._devices = None
        if devices is not None:
            devices = [int(d) for d in devices]
            if not devices:
                raise ValueError('devices must be a non-empty sequence of '
                                 'CUDA device ordinals.')
            if device is not None and int(device) != devices[0]:
                raise ValueError('Give either device or devices.')
            device = devices[0]
            if len(devices) > 1:
                self._devices = devices
        if device is None:
            device = int(os.environ.get('LOCAL_RANK', 0))
        self._device = int(device)

        N = self._num_collocation_nodes
        if node_range is None:
            node_range = (0, N - 1)
        lo, hi = int(node_range[0]), int(node_range[1])
        if not (0 <= lo < hi <= N - 1):
            raise ValueError('node_range must satisfy 0 <= lo < hi <= N - 1.')
        self._node_range = (lo, hi)

        self._classify_parameters()
        self._classify_trajectories()
        self._num_free = ((self.num_states +
                           self.num_unknown_input_trajectories) * N +
                          self.num_unknown_parameters +
                          int(self._variable_duration))
        self._check_known_trajectories()

        if integration_method not in _METHODS:
            raise ValueError('{} is not a valid integration method.'.format(
                integration_method))
        self._integration_method = integration_method
        self._make_discrete_symbols()
        self._discretize_eom()

        self._num_constraints = self.num_eom * (N - 1)
        if instance_constraints is None:
            self._num_instance_constraints = 0
        else:
            self._num_instance_constraints = len(instance_constraints)
            self._num_constraints += self._num_instance_constraints
            self._index_instance_constraints()
            self.eval_instance_constraints = \
                self._instance_constraints_func()
            self.eval_instance_constraints_jacobian_values = \
                self._instance_constraints_jacobian_values_func()

        self._evaluator = None

    # ------------------------------------------------------------------
    # read-only attributes, same names as the reference
    # (opty/direct_collocation.py:1556-1892)
    # ------------------------------------------------------------------
    eom = property(lambda self: self._eom)
    discrete_eom = property(lambda self: self._discrete_eom)
    state_symbols = property(lambda self: self._state_symbols)
    state_derivative_symbols = property(
        lambda self: self._state_derivative_symbols)
    time_symbol = property(lambda self: self._time_symbol)
    time_interval_symbol = property(lambda self: self._time_interval_symbol)
    node_time_interval = property(lambda self: self._node_time_interval)
    num_collocation_nodes = property(
        lambda self: self._num_collocation_nodes)
    num_constraints = property(lambda self: self._num_constraints)
    num_free = property(lambda self: self._num_free)
    num_eom = property(lambda self: self._eom.shape[0])
    num_states = property(lambda self: len(self._state_symbols))
    num_instance_constraints = property(
        lambda self: self._num_instance_constraints)
    instance_constraints = property(lambda self: self._instance_constraints)
    integration_method = property(lambda self: self._integration_method)
    known_parameter_map = property(lambda self: self._known_parameter_map)
    known_trajectory_map = property(lambda self: self._known_trajectory_map)
    known_parameters = property(lambda self: self._known_parameters)
    unknown_parameters = property(lambda self: self._unknown_parameters)
    parameters = property(lambda self: self._parameters)
    num_known_parameters = property(lambda self: len(self._known_parameters))
    num_unknown_parameters = property(
        lambda self: len(self._unknown_parameters))
    num_parameters = property(lambda self: len(self._parameters))
    known_input_trajectories = property(
        lambda self: self._known_input_trajectories)
    unknown_input_trajectories = property(
        lambda self: self._unknown_input_trajectories)
    input_trajectories = property(lambda self: self._input_trajectories)
    num_known_input_trajectories = property(
        lambda self: len(self._known_input_trajectories))
    num_unknown_input_trajectories = property(
        lambda self: len(self._unknown_input_trajectories))
    num_input_trajectories = property(
        lambda self: len(self._input_trajectories))
    previous_discrete_state_symbols = property(lambda self: self._xp)
    current_discrete_state_symbols = property(lambda self: self._xi)
    next_discrete_state_symbols = property(lambda self: self._xn)
    current_known_discrete_specified_symbols = property(lambda self: self._ki)
    next_known_discrete_specified_symbols = property(lambda self: self._kn)
    current_unknown_discrete_specified_symbols = property(
        lambda self: self._ui)
    next_unknown_discrete_specified_symbols = property(lambda self: self._un)
    current_discrete_specified_symbols = property(
        lambda self: self._ki + self._ui)
    next_discrete_specified_symbols = property(
        lambda self: self._kn + self._un)
    parallel = property(lambda self: self._parallel)
    show_compile_output = property(lambda self: self._show_compile_output)
    tmp_dir = property(lambda self: self._tmp_dir)
    node_range = property(lambda self: self._node_range)
    device = property(lambda self: self._device)
    devices = property(lambda self: self._devices or [self._device])

    @integration_method.setter
    def integration_method(self, method):
        if method not in _METHODS:
            raise ValueError('{} is not a valid integration method.'.format(
                method))
        self._integration_method = method
        self._discretize_eom()
        self._evaluator = None

    # ------------------------------------------------------------------
    # symbol bookkeeping
    # ------------------------------------------------------------------
    @staticmethod
    def _split_known(everything, known):
        """Known symbols keep the order the user gave them in, unknown ones
        are sorted by name (opty/direct_collocation.py:1928-1952)."""
        everything = set(everything)
        known = tuple(known)
        if not everything:
            if known:
                raise ValueError('{} are not in the provided equations of '
                                 'motion.'.format(known))
            return (), ()
        return known, tuple(sort_sympy(everything.difference(known)))

    def _classify_parameters(self):
        # opty/direct_collocation.py:1954-1973
        symbols = set(self._eom.free_symbols)
        symbols.discard(self._time_symbol)
        known, unknown = self._split_known(symbols,
                                           self._known_parameter_map.keys())
        self._known_parameters = known
        self._unknown_parameters = unknown
        self._parameters = known + unknown

    def _classify_trajectories(self):
        # opty/direct_collocation.py:1988-2035
        state_like = set(self._state_symbols) | \
            set(self._state_derivative_symbols)
        others = me.find_dynamicsymbols(self._eom).difference(state_like)
        if any(isinstance(f, sm.Derivative) for f in others):
            raise ValueError('Too few state variables provided for state '
                             'time derivatives found in equations of motion.')
        self._deriv_in_knw_traj = False
        for f in others:
            if f.args == (self._time_symbol,):
                continue
            if len(f.args) > 1:
                raise ValueError('{} is a function of more than one '
                                 'variable.'.format(f))
            self._deriv_in_knw_traj = True
        names = [f.name for f in others]
        if len(set(names)) != len(names):
            raise ValueError('Repeated input trajectory variable fnames not '
                             'allowed: {}'.format(names))
        known, unknown = self._split_known(others,
                                           self._known_trajectory_map.keys())
        self._known_input_trajectories = known
        self._unknown_input_trajectories = unknown
        self._input_trajectories = known + unknown

    def _check_known_trajectories(self):
        # opty/direct_collocation.py:1975-1986
        N = self._num_collocation_nodes
        for sym, val in self._known_trajectory_map.items():
            if _is_callable_value(val):
                val = val(np.ones(self._num_free))
            if len(val) != N:
                raise ValueError('The known parameter {} is not length '
                                 '{}.'.format(sym, N))

    def _make_discrete_symbols(self):
        # naming scheme of opty/direct_collocation.py:2070-2118
        def tagged(funcs, tag):
            return tuple(sm.Symbol(f.__class__.__name__ + tag, real=True)
                         for f in funcs)
        self._xp = tagged(self._state_symbols, 'p')
        self._xi = tagged(self._state_symbols, 'i')
        self._xn = tagged(self._state_symbols, 'n')
        self._ki = tuple(self._discrete_known_input(f, 'i')
                         for f in self._known_input_trajectories)
        self._kn = tuple(self._discrete_known_input(f, 'n')
                         for f in self._known_input_trajectories)
        self._ui = tagged(self._unknown_input_trajectories, 'i')
        self._un = tagged(self._unknown_input_trajectories, 'n')

    def _discrete_known_input(self, f, tag):
        """Discrete stand-in of a known input (opty/direct_collocation.py:
        2080-2093): ``r(t) -> ri``; an implicit function of time ``r(x(t))``
        stays a function ``ri(xi)`` so that differentiation applies the chain
        rule; its user-supplied derivative ``dr/dx`` becomes the symbol
        ``dri_dxi``."""
        if isinstance(f, sm.Derivative):
            var, (wrt, _) = f.args
            return sm.Symbol('d{}{}_d{}{}'.format(
                var.__class__.__name__, tag, wrt.__class__.__name__, tag),
                real=True)
        if f.args[0] != self._time_symbol:
            inner = sm.Symbol(f.args[0].__class__.__name__ + tag, real=True)
            return sm.Function(f.__class__.__name__ + tag, real=True)(inner)
        return sm.Symbol(f.__class__.__name__ + tag, real=True)

    def _chain_rules(self):
        """``[(discrete function, discrete state symbol, derivative symbol)]``
        for every implicit known trajectory ``r(x(t))``: the seed
        ``d ri(xi) / d xi = dri_dxi`` of the forward-mode differentiation
        (the reference reaches the same through SymPy's unevaluated
        ``Derivative`` and the replacements of opty/direct_collocation.py:
        2284-2302, 2760-2793)."""
        rules = []
        known = self._known_input_trajectories
        for tag, symbols in (('i', self._ki), ('n', self._kn)):
            for f, disc in zip(known, symbols):
                if isinstance(f, sm.Derivative) or \
                        f.args[0] == self._time_symbol:
                    continue
                deriv = f.diff(f.args[0])
                if deriv not in known:
                    raise ValueError(
                        'The known trajectory {} is a function of {}; its '
                        'derivative {} must be in known_trajectory_map too.'
                        .format(f, f.args[0], deriv))
                rules.append((disc, disc.args[0],
                              symbols[known.index(deriv)]))
        return rules

    def _discretize_eom(self):
        """Backward Euler: x' -> (xi - xp)/h, x -> xi, u -> ui.  Midpoint:
        x' -> (xn - xi)/h, x -> (xi + xn)/2, u -> (ui + un)/2
        (opty/direct_collocation.py:2143-2156)."""
        logger.info('Discretizing the equations of motion.')
        h = self._time_interval_symbol
        x, xd = self._state_symbols, self._state_derivative_symbols
        u = self._input_trajectories
        ui = self._ki + self._ui
        un = self._kn + self._un
        if self._integration_method == 'backward euler':
            rates = {d: (c - p) / h for d, c, p in zip(xd, self._xi, self._xp)}
            values = dict(zip(x + u, self._xi + ui))
            self._discrete_eom = me.msubs(self._eom, rates, values)
        else:
            rates = {d: (nx - c) / h
                     for d, c, nx in zip(xd, self._xi, self._xn)}
            mid_x = {f: (c + nx) / 2
                     for f, c, nx in zip(x, self._xi, self._xn)}
            mid_u = {f: (c + nx) / 2 for f, c, nx in zip(u, ui, un)}
            self._discrete_eom = me.msubs(self._eom, rates, mid_x, mid_u)

    # ------------------------------------------------------------------
    # instance constraints: o scalar expressions, evaluated on the host
    # (opty/direct_collocation.py:2158-2282)
    # ------------------------------------------------------------------
    def _free_index_of(self, func):
        N = self._num_collocation_nodes
        arg = func.args[0]
        if self._variable_duration:
            if arg == 0:
                node = 0
            else:
                try:
                    node = int(arg / self._time_interval_symbol)
                except TypeError as err:
                    raise TypeError(
                        'Instance constraint {} is not a correct integer '
                        'multiple of the time interval.'.format(func)) from err
            if node not in range(N):
                raise ValueError(
                    'Instance constraint {} gives an index of {} which is not '
                    'between 0 and {}.'.format(func, node, N - 1))
        else:
            duration = self._node_time_interval * (N - 1)
            grid = np.linspace(0.0, duration, num=N)
            node = int(np.argmin(np.abs(grid - float(arg))))
        of_time = func.__class__(self._time_symbol)
        if of_time in self._state_symbols:
            return node + self._state_symbols.index(of_time) * N
        if of_time in self._unknown_input_trajectories:
            return (node + self.num_states * N +
                    self._unknown_input_trajectories.index(of_time) * N)
        return None

    def _index_instance_constraints(self):
        atoms = set()
        for con in self._instance_constraints:
            atoms |= con.atoms(sm.Function)
        self.instance_constraint_function_atoms = atoms
        self.instance_constraints_free_index_map = {
            f: self._free_index_of(f) for f in atoms}

    def _instance_lambdify(self, exprs):
        vec = sm.DeferredVector('FREE')
        subs = {f: vec[i] for f, i in
                self.instance_constraints_free_index_map.items()}
        known = list(self._known_parameter_map.keys())
        return sm.lambdify([vec] + known, [e.subs(subs) for e in exprs],
                           modules=[{'ImmutableMatrix': np.array}, 'numpy'])

    def _instance_constraints_func(self):
        f = self._instance_lambdify(self._instance_constraints)
        return lambda free: f(free, *self._known_parameter_map.values())

    def _instance_constraints_jacobian_indices(self):
        # entry order inside one constraint follows ``con.atoms`` exactly as
        # in the reference (opty/direct_collocation.py:2243-2249)
        base = self.num_eom * (self._num_collocation_nodes - 1)
        rows, cols = [], []
        for i, con in enumerate(self._instance_constraints):
            for f in con.atoms(sm.Function):
                rows.append(base + i)
                cols.append(self.instance_constraints_free_index_map[f])
        return np.array(rows, dtype=int), np.array(cols, dtype=int)

    def _instance_constraints_jacobian_values_func(self):
        partials = []
        for con in self._instance_constraints:
            partials.extend(con.diff(f) for f in con.atoms(sm.Function))
        f = self._instance_lambdify(partials)
        count = len(partials)

        def values(free):
            out = f(free, *self._known_parameter_map.values())
            return np.asarray(out, dtype=float).reshape(count)
        return values

    # ------------------------------------------------------------------
    # the CUDA evaluator
    # ------------------------------------------------------------------
    def _program_inputs(self):
        """Rows of the device trajectory matrix, the uniform arguments and the
        differentiation variables.

        Argument order of the reference's generated functions:
        opty/direct_collocation.py:2345-2364 (constraints) and :2713-2747
        (Jacobian, which also defines the ``wrt`` = Jacobian column order).
        """
        midpoint = self._integration_method == 'midpoint'
        rows = []
        if midpoint:
            rows += list(zip(self._xi, self._xn))
            rows += list(zip(self._ui, self._un))
            rows += list(zip(self._ki, self._kn))
            wrt = self._xi + self._xn + self._ui + self._un
        else:
            rows += list(zip(self._xp, self._xi))
            rows += [(None, s) for s in self._ui]
            rows += [(None, s) for s in self._ki]
            wrt = self._xi + self._xp + self._ui
        wrt += self._unknown_parameters
        if self._variable_duration:
            wrt += (self._time_interval_symbol,)
        uniform = list(self._parameters) + [self._time_interval_symbol]
        return rows, uniform, list(wrt)

    def _build_program(self):
        rows, uniform, wrt = self._program_inputs()
        return CollocationProgram(
            list(self.discrete_eom), rows, uniform, wrt,
            use_sympy_cse=self._cuda_options['use_sympy_cse'],
            chain_rules=self._chain_rules())

    def prepare_module(self):
        """Lowers, emits and compiles this problem's CUDA module without
        touching a GPU; returns the ``_PreparedModule``."""
        return _PreparedModule(self)

    def _build_evaluator(self):
        if self._evaluator is None:
            if self._devices:
                self._evaluator = _MultiDeviceEvaluator(self, self._devices)
            else:
                self._evaluator = _CudaEvaluator(self)
        return self._evaluator

    def close(self):
        """Releases the device and pinned host buffers."""
        if self._evaluator is not None:
            self._evaluator.close()
            self._evaluator = None

    def generate_constraint_function(self):
        """Returns ``f(free) -> ndarray, shape(M*(N-1) + o,)`` ordered
        ``[eom_1 @ nodes, ..., eom_M @ nodes, c_1..c_o]``
        (opty/direct_collocation.py:3003-3008, :127-132)."""
        logger.info('Generating constraint function.')
        ev = self._build_evaluator()
        return ev.constraints

    def generate_jacobian_function(self):
        """Returns ``f(free) -> ndarray, shape(nnz,)``: node-major
        ``[node][eom][wrt]`` partials followed by the instance-constraint
        entries (opty/direct_collocation.py:3010-3015, :2681-2688).  The
        array is a view of a persistent pinned buffer, valid until the next
        call (the reference returns a view of a persistent buffer too,
        opty/direct_collocation.py:2814, :2887)."""
        logger.info('Generating jacobian function.')
        ev = self._build_evaluator()
        return ev.jacobian

    def jacobian_indices(self):
        """COO row and column indices (int64) matching
        ``generate_jacobian_function``'s values; bit-equal to the reference's
        Python loop (opty/direct_collocation.py:2450-2690) but generated by a
        CUDA kernel."""
        lo, hi = self._node_range
        method = 1 if self._integration_method == 'midpoint' else 0
        rows, cols = runtime.jacobian_indices(
            self._device, self._num_collocation_nodes, lo, hi,
            self.num_states, self.num_unknown_input_trajectories,
            self.num_unknown_parameters, int(self._variable_duration),
            self.num_eom, method)
        if self._instance_constraints is not None and self._owns_instance():
            irows, icols = self._instance_constraints_jacobian_indices()
            rows = np.concatenate((rows, irows.astype(np.int64)))
            cols = np.concatenate((cols, icols.astype(np.int64)))
        return rows, cols

    def _owns_instance(self):
        """Instance constraints are appended only by the collocator that
        covers the whole node range (shards leave them to the gather step)."""
        return self._node_range == (0, self._num_collocation_nodes - 1)


def prepare_program_module(prog, num_nodes, method, opts, tmp_dir=None,
                           show_compile_output=False, pair=None):
    """Groups, emits and compiles the CUDA module of a
    :class:`CollocationProgram` for ``num_nodes`` evaluation nodes.  Needs nvcc
    but no GPU.  Returns ``(parts, derived, source, meta, cubin, cubin_path,
    cache_hit)``.

    ``persistent='auto'`` picks the row-stationary kernel for collocation
    programs whose equations have an even number of partials and whose input
    windows fit next to the staging buffers of an 8-warp block in shared
    memory (up to ~30 trajectory + derived rows: the 10-link pendulum), the
    grid kernel otherwise."""
    if opts['persistent'] == 'auto':
        eligible = (pair is not None and prog.P % 2 == 0 and
                    bool(opts['tma_store']) and opts['tma_load'] is True and
                    opts['groups'] == 'auto' and
                    opts['warps_per_block'] == 'auto' and
                    opts['compile_shards'] in ('auto', 1) and
                    prog.stats()['varying_cost'] < 40000)
        if eligible:
            try:
                return _prepare_program_module(
                    prog, num_nodes, method,
                    dict(opts, persistent='stationary', tile_bufs=1),
                    tmp_dir, show_compile_output, pair)
            except ValueError as err:
                logger.info('Row-stationary kernel not used: %s', err)
        opts = dict(opts, persistent=False)
    return _prepare_program_module(prog, num_nodes, method, opts, tmp_dir,
                                   show_compile_output, pair)


def _prepare_program_module(prog, num_nodes, method, opts, tmp_dir=None,
                            show_compile_output=False, pair=None):
    M = prog.M
    K = M * prog.P
    tma_store = bool(opts['tma_store']) and K % 2 == 0
    if opts['tma_load'] == 'direct':
        tma_load = 2
    else:
        tma_load = int(bool(opts['tma_load']) and prog.R <= 256)
    stationary = opts['persistent'] in ('stationary', 2)
    groups = opts['groups']
    if stationary and groups == 'auto':
        groups = M
    if groups == 'auto':
        node_warps = -(-num_nodes // 32)
        g_par = -(-int(opts['target_warps']) // node_warps)
        g_cost = int(np.ceil(prog.stats()['varying_cost'] /
                             float(opts['max_group_cost'])))
        groups = max(1, g_par, g_cost)
    groups = int(min(groups, M, runtime.OPTY_MAX_GROUPS))
    # with TMA stores a group starts at an even column (odd P: even row)
    align = 2 if tma_store else 1

    tile_cols = codegen.choose_tile_cols(opts['tile_cols'], prog.P,
                                         even=tma_store)
    unit_rows = 1 if (prog.P % 2 == 0 or not tma_store) else 2
    phases_per_unit = len(codegen.row_phases(
        0, unit_rows * prog.P, tile_cols, pair if unit_rows == 1 else None,
        tma_store))

    const_rows = []
    if opts['const_rows'] and unit_rows == 1 and tma_store and \
            pair is not None and (stationary or opts['const_rows'] == 'grid'):
        # (with the grid kernel only on request, 'grid': its blocks then send
        # the runs of a share of their tile's nodes -- or a separate kernel
        # does, `const_kernel` -- which measured the same as leaving the rows
        # to store-only groups at the 50-link chain, 6.4-6.6 ms, and slower at
        # the 10-link pendulum, profiles/r03z_*, r04b_*)
        kinds = prog.entry_kind()
        const_rows = [j for j in range(M)
                      if max(kinds[j * prog.P:(j + 1) * prog.P]) < 2]
        if len(const_rows) == M:
            const_rows = const_rows[1:]     # the main kernel needs a group
        if not stationary and len(const_rows) * prog.P * 8 > 200 * 1024:
            const_rows = []

    def drop_const(rows):
        # constant rows are not output groups: cut them out of the ranges
        out = []
        for r0, r1 in rows:
            start = None
            for j in range(r0, r1 + 1):
                if j < r1 and j not in const_rows:
                    if start is None:
                        start = j
                elif start is not None:
                    out.append((start, j))
                    start = None
        return out

    def column_parts(stop=None):
        if stationary and groups >= M // unit_rows:
            # one equation row (two when P is odd) per group
            rows = [(r, min(M, r + unit_rows)) for r in range(0, M, unit_rows)
                    if r not in const_rows]
            return [(r0 * prog.P, r1 * prog.P) for r0, r1 in rows]
        rows = prog.partition_rows(groups, col_align=align, stop=stop)
        if opts['groups'] == 'auto':
            # A warp has one tile store in flight per staging buffer, so a
            # group is a serial chain of its phases.  Balanced by operation
            # count alone, the 48 kinematic equations of the 50-link chain
            # (480 operations, 192 store-only phases) form ONE group that
            # takes 2.6x as long as the heaviest dynamic equation
            # (profiles/r02k_*): long chains are cut.
            cut = []
            for r0, r1 in rows:
                units = -(-(r1 - r0) // unit_rows)
                pieces = -(-units * phases_per_unit // MAX_PHASES_PER_GROUP)
                pieces = max(1, min(pieces, units))
                step = -(-units // pieces) * unit_rows
                r = r0
                while r < r1:
                    cut.append((r, min(r1, r + step)))
                    r += step
            if len(cut) <= runtime.OPTY_MAX_GROUPS:
                rows = cut
        if const_rows:
            rows = drop_const(rows)
        return [(r0 * prog.P, r1 * prog.P) for r0, r1 in rows]

    parts = column_parts()
    derived = []
    if opts['pre_pass'] and len(parts) > 1:
        derived = prog.select_derived(parts, max_rows=max(0, 256 - prog.R))
        if derived:
            # re-balance with the shared work taken out of the groups
            parts = column_parts(stop=set(derived))
    if prog.R + len(derived) > 256 and tma_load == 1:
        tma_load = 0
    wpb = opts['warps_per_block']
    if wpb == 'auto':
        node_warps = -(-num_nodes // 32)
        wpb = 8 if node_warps * len(parts) >= 8 * 148 * 6 else 2
        if stationary:
            wpb = 8
    wpb = int(wpb)
    mbs = opts['min_blocks_per_sm']
    if stationary:
        mbs = 1 if mbs == 'auto' else int(mbs)
        tma_load = 1
    if mbs == 'auto':
        # plain-order bodies need the whole register file of 8 warps
        light = opts['schedule'] is False or (
            opts['schedule'] == 'auto' and
            prog.stats()['varying_cost'] / max(len(parts), 1) < 1500)
        mbs = max(1, (8 if light else 16) // wpb)
    mbs = int(mbs)
    tile_bufs = int(opts['tile_bufs'])
    if tma_load != 2 and not stationary:
        # staged input must fit beside the staging buffers in 227 KB of
        # shared memory per block; otherwise the lanes read the trajectory
        # matrix directly (coalesced, read-only path)
        threads = 32 * wpb
        xseg = min(threads, 128)
        xin = (threads // xseg) * (
            -(-((prog.R + len(derived)) * (xseg + 2) * 8) // 128) * 128)
        tiles = wpb * tile_bufs * 32 * tile_cols * 8
        if xin + tiles + 128 > 227 * 1024:
            tma_load = 2

    flags = build.module_flags(fmad=opts['fmad'],
                               maxrregcount=opts['maxrregcount'])
    sched_opts = {k: opts[k] for k in codegen.SCHEDULE_DEFAULTS}
    cost = prog.stats()['varying_cost']
    emit_kwargs = dict(
        tile_cols=tile_cols, warps_per_block=wpb, min_blocks_per_sm=mbs,
        tma_load=tma_load, tma_store=tma_store, derived=derived,
        debug_nostore=opts['debug_nostore'], tile_bufs=tile_bufs, pair=pair,
        persistent=2 if stationary else int(bool(opts['persistent'])),
        num_nodes=num_nodes, blocks_per_sm=mbs if stationary else 1,
        const_rows=const_rows, const_head_pct=opts['const_head_pct'],
        store_hint=opts['store_hint'], fused_pre=opts['fused_pre'],
        item_cost=opts['item_cost'],
        strided_schedule=opts['strided_schedule'],
        const_pre_pct=opts['const_pre_pct'],
        const_kernel=opts['const_kernel'],
        tile_major=(wpb >= 8) if opts['tile_major'] == 'auto'
        else bool(opts['tile_major']),
        schedule_options=sched_opts,
        workers=(os.cpu_count() or 1) if cost >= 20000 else 1)

    # Large problems are split into several modules (contiguous ranges of
    # output groups) that nvcc compiles in parallel; the runtime launches one
    # main kernel per module.  The reference compiles its one generated C
    # function serially (opty/utils.py:866-907); at the 50-link chain that is
    # the dominant set-up cost.
    shards = opts['compile_shards']
    if stationary:
        shards = 1
    if shards == 'auto':
        big = cost >= 40000 and len(parts) >= 4
        shards = min(len(parts), os.cpu_count() or 1, 16) if big else 1
    shards = max(1, min(int(shards), len(parts)))
    if shards == 1:
        ranges = [None]
    else:
        # contiguous chunks of groups with balanced cost
        costs = [prog.range_cost(c0, c1, set(derived) or None)
                 for c0, c1 in parts]
        total = float(sum(costs)) or 1.0
        cuts, acc = [], 0.0
        for g, c in enumerate(costs[:-1]):
            acc += c
            # cut after group g once the running cost passes the next target,
            # keeping enough groups for the remaining chunks
            if len(cuts) < shards - 1 and \
                    (acc >= total * (len(cuts) + 1) / shards or
                     len(parts) - (g + 1) == shards - 1 - len(cuts)):
                cuts.append(g + 1)
        bounds = [0] + cuts + [len(parts)]
        ranges = [(a, b) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]

    logger.info('Scheduling and emitting the CUDA module%s.',
                '' if len(ranges) == 1 else 's ({})'.format(len(ranges)))
    emitted = []
    for i, rng in enumerate(ranges):
        emitted.append(codegen.emit_module(
            prog, parts, method, only_groups=rng, with_aux=(i == 0),
            **emit_kwargs))
    source, meta = emitted[0]
    logger.info('Compiling the constraint and Jacobian kernels.')

    def compile_one(src):
        return build.compile_module(src, flags, cache_dir=tmp_dir,
                                    show_compile_output=show_compile_output)

    if len(emitted) == 1:
        compiled = [compile_one(source)]
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(emitted),
                                                os.cpu_count() or 1)) as ex:
            compiled = list(ex.map(compile_one, [e[0] for e in emitted]))
    cubin, cubin_path, cache_hit = compiled[0]
    cache_hit = all(c[2] for c in compiled)
    meta['extra_modules'] = [
        {'cubin_path': c[1], 'group_range': e[1]['group_range'],
         'num_groups': e[1]['num_groups'], 'groups': e[1]['groups']}
        for e, c in zip(emitted[1:], compiled[1:])]
    return parts, derived, source, meta, cubin, cubin_path, cache_hit


def attach_extra_modules(handle, meta):
    """Loads the additional modules of a problem that was compiled in
    several pieces into ``handle``."""
    for em in meta.get('extra_modules', ()):
        with open(em['cubin_path'], 'rb') as f:
            handle.add_module(f.read())


class _PreparedModule(object):
    """Tape program, emitted CUDA-C and compiled cubin of one collocator.
    Building it needs nvcc but no GPU (``__graft_entry__.build`` uses this to
    fill the compiled-module cache ahead of time)."""

    def __init__(self, col):
        opts = col._cuda_options
        lo, hi = col._node_range
        key = self._input_key(col, hi - lo) if opts['use_index'] else None
        if key is not None and self._load(col, key):
            return
        logger.info('Lowering and differentiating the constraint function.')
        prog = col._build_program()
        self.program = prog
        (self.parts, self.derived, self.source, self.meta, self.cubin,
         self.cubin_path, self.cache_hit) = prepare_program_module(
            prog, hi - lo, col.integration_method, opts, tmp_dir=col.tmp_dir,
            show_compile_output=col.show_compile_output,
            pair=col.num_states)
        self.index_hit = False
        if key is not None and self.cubin_path:
            build.store_index(col.tmp_dir, key, {
                'meta': self.meta, 'parts': [list(p) for p in self.parts],
                'derived': list(self.derived),
                'cpu_count': os.cpu_count(),
                'cubin': os.path.basename(self.cubin_path),
                'extra': [os.path.basename(em['cubin_path'])
                          for em in self.meta.get('extra_modules', ())]})

    # -- set-up cache in front of the symbolic work -------------------------
    # The reference hashes the generated C *after* CSE, differentiation and
    # printing, so even a cache hit repeats the symbolic work (opty/utils.py:
    # 745-770: ~8 s at the 10-link pendulum, minutes at 50 links).  Here the
    # key is taken from the inputs of that work: the discrete EOM, the symbol
    # layout, the options and the emitter / skeleton versions.
    @staticmethod
    def _input_key(col, num_nodes):
        import hashlib
        hasher = hashlib.sha256()
        rows, uniform, wrt = col._program_inputs()
        for expr in col.discrete_eom:
            build.expression_digest(hasher, expr)
            hasher.update(b'|')
        hasher.update(repr([[str(a), str(b)] for a, b in rows]).encode())
        hasher.update(repr([str(u) for u in uniform]).encode())
        hasher.update(repr([str(w) for w in wrt]).encode())
        hasher.update(repr([[str(x) for x in rule]
                            for rule in col._chain_rules()]).encode())
        opts = {k: v for k, v in col._cuda_options.items()
                if k not in ('out_ring', 'prefetch_jacobian',
                             'd2h_skip_constants', 'use_index')}
        hasher.update(repr(sorted(opts.items())).encode())
        hasher.update(repr((num_nodes, col.integration_method,
                            codegen.EMITTER_VERSION)).encode())
        hasher.update(build._header_digest().encode())
        # the lowering / emitter code itself: any change invalidates the index
        here = os.path.dirname(os.path.abspath(__file__))
        for name in ('codegen.py', 'ir.py', 'lowering.py', 'program.py',
                     'schedule.py', 'direct_collocation.py', 'build.py'):
            with open(os.path.join(here, name), 'rb') as f:
                hasher.update(f.read())
        return hasher.hexdigest()[:32]

    def _load(self, col, key):
        idx = build.load_index(col.tmp_dir, key)
        if not idx:
            return False
        # the automatic number of parallel compile shards follows the host's
        # core count: a multi-module entry is only reused on a like host
        if idx['extra'] and idx.get('cpu_count') != os.cpu_count() and \
                col._cuda_options['compile_shards'] == 'auto':
            return False
        cache_dir = col.tmp_dir or build.default_cache_dir()
        paths = [os.path.join(cache_dir, name)
                 for name in [idx['cubin']] + idx['extra']]
        if not all(os.path.exists(p) and os.path.getsize(p) for p in paths):
            return False
        meta = idx['meta']
        for em, path in zip(meta.get('extra_modules', ()), paths[1:]):
            em['cubin_path'] = path
        self.meta = meta
        self.parts = [tuple(p) for p in idx['parts']]
        self.derived = idx['derived']
        self.source = None
        self.cubin_path = paths[0]
        with open(paths[0], 'rb') as f:
            self.cubin = f.read()
        self.cache_hit = True
        self.index_hit = True
        self.program = _ProgramInfo(meta)
        logger.info('Skipped lowering and compile, index %s loaded.', key)
        return True


class _ProgramInfo(object):
    """What the runtime side needs to know about a program whose module came
    out of the set-up cache (sizes only; the tape was never built)."""

    def __init__(self, meta):
        self.M, self.P, self.K, self.R = (meta['M'], meta['P'], meta['K'],
                                          meta['R'])


class _CudaEvaluator(object):
    """Owns the generated module and the runtime handle of one collocator."""

    def __init__(self, col):
        self.col = col
        opts = col._cuda_options
        prepared = _PreparedModule(col)
        prog = self.program = prepared.program
        self.parts = prepared.parts
        self.source = prepared.source
        meta = self.meta = prepared.meta
        cubin = prepared.cubin
        self.cubin_path = prepared.cubin_path
        self.cache_hit = prepared.cache_hit
        lo, hi = col._node_range
        nn = hi - lo
        M, P = prog.M, prog.P
        K = M * P

        o = col.num_instance_constraints if col._owns_instance() else 0
        if o:
            irows, _ = col._instance_constraints_jacobian_indices()
            nnz_inst = len(irows)
        else:
            nnz_inst = 0
        self.num_inst = o
        self.nnz_inst = nnz_inst

        cfg = runtime.ColloCfg()
        cfg.abi_version = runtime.ABI_VERSION
        cfg.device = col._device
        cfg.N = col.num_collocation_nodes
        cfg.node_lo, cfg.node_hi = lo, hi
        cfg.n = col.num_states
        cfg.q = col.num_unknown_input_trajectories
        cfg.k = col.num_known_input_trajectories
        cfg.r = col.num_unknown_parameters
        cfg.s = int(col._variable_duration)
        cfg.pk = col.num_known_parameters
        cfg.M, cfg.P = M, P
        cfg.method = 1 if col.integration_method == 'midpoint' else 0
        cfg.out_ring = int(opts['out_ring'])
        cfg.prefetch_jac = int(bool(opts['prefetch_jacobian']))
        cfg.con_tail = o
        cfg.jac_tail = nnz_inst
        cfg.h = 0.0 if col._variable_duration else float(
            col.node_time_interval)
        self.handle = runtime.ColloHandle(cfg, cubin)
        attach_extra_modules(self.handle, meta)
        self.nn = nn
        self.con_len = M * nn
        self.jac_len = nn * K

        self._callable_known = any(
            _is_callable_value(v) for v in col.known_trajectory_map.values())
        self._pushed_known = None
        self._push_known(None)

        if opts['d2h_skip_constants']:
            self._setup_constant_elision()

    # known values -----------------------------------------------------
    def _push_known(self, free):
        """Pushes the known parameter values and known input trajectories to
        the device if they differ from what is there.  The maps are re-read on
        every callback like the reference does (``_merge_fixed_free``,
        opty/direct_collocation.py:2891-2926, called from every
        ``constraints`` / ``jacobian``, :2973-2980): a user may change a
        parameter between two solves.  Callables see the free vector on every
        evaluation (opty/direct_collocation.py:2916-2917)."""
        col = self.col
        N = col.num_collocation_nodes
        traj = None
        if col.num_known_input_trajectories:
            traj = np.empty((col.num_known_input_trajectories, N))
            for i, sym in enumerate(col.known_input_trajectories):
                val = col.known_trajectory_map[sym]
                if _is_callable_value(val):
                    val = val(np.ones(col.num_free) if free is None else free)
                traj[i] = val
        params = None
        if col.num_known_parameters:
            params = np.array([float(col.known_parameter_map[p])
                               for p in col.known_parameters])
        last = self._pushed_known
        if last is not None and \
                (traj is None or np.array_equal(traj, last[0])) and \
                (params is None or np.array_equal(params, last[1])):
            return
        self.handle.set_known(traj, params)
        self._pushed_known = (traj, params)

    def _setup_constant_elision(self):
        """Jacobian columns whose value cannot change between calls --
        literals, and node-invariant entries when there is no free parameter
        or time interval (about half of all columns at the 10-link pendulum,
        SURVEY.md §7) -- are copied to the pinned host buffer once, with the
        first full device->host copy.  Later copies only move the column
        ranges that hold call-dependent entries."""
        col = self.col
        kinds = np.array(self.meta['entry_kind'])
        frozen_invariants = (col.num_unknown_parameters == 0 and
                             not col._variable_duration)
        changing = kinds == 2
        if not frozen_invariants:
            changing |= kinds == 1
        idx = np.nonzero(changing)[0]
        if len(idx) == 0:
            ranges = [(0, 1)]
        else:
            # merge runs of changing columns separated by short gaps
            min_gap = 32
            ranges = []
            start = prev = int(idx[0])
            for c in idx[1:]:
                c = int(c)
                if c - prev > min_gap:
                    ranges.append((start, prev + 1))
                    start = c
                prev = c
            ranges.append((start, prev + 1))
            if len(ranges) > 8:
                ranges = [(ranges[0][0], ranges[-1][1])]
        self.d2h_ranges = ranges
        self.handle.set_d2h_columns(ranges)

    # callbacks --------------------------------------------------------
    def _check_free(self, free):
        free = np.asarray(free, dtype=np.float64)
        if free.shape != (self.col.num_free,):
            raise ValueError('free must have shape ({},), got {}.'.format(
                self.col.num_free, free.shape))
        return np.ascontiguousarray(free)

    def constraints(self, free):
        free = self._check_free(free)
        self._push_known(free)
        buf = self.handle.constraints(free)
        if self.num_inst:
            buf[self.con_len:] = self.col.eval_instance_constraints(free)
        return buf.copy()

    def jacobian(self, free, refetch=False):
        """``refetch=True`` copies the whole Jacobian block from the device
        again, including the columns that cannot have changed (see
        :meth:`Problem.jacobian`)."""
        free = self._check_free(free)
        self._push_known(free)
        if refetch:
            self.handle.invalidate_host_jacobian()
        buf = self.handle.jacobian(free)
        if self.num_inst:
            buf[self.jac_len:] = \
                self.col.eval_instance_constraints_jacobian_values(free)
        return buf

    def close(self):
        self.handle.close()


class _MultiDeviceEvaluator(object):
    """Several GPUs driven by one process: one :class:`_CudaEvaluator` per
    device, each on a contiguous shard of the constraint nodes.  The shards
    need no data-path collective (node ``i`` reads trajectory columns ``i``
    and ``i + 1`` only, opty/direct_collocation.py:2145, 2153-2155): every
    device receives its window of the free vector and copies its residuals
    (``M`` strided segments of the eom-major vector, opty/direct_collocation
    .py:2446) and its Jacobian block (one contiguous slice of the node-major
    vector, opty/direct_collocation.py:2681-2684) straight into ONE pinned
    host vector each -- what a host-side IPOPT consumes."""

    def __init__(self, col, devices):
        import copy
        from .sharding import node_shard
        self.col = col
        lo, hi = col._node_range
        G = len(devices)
        if hi - lo < G:
            raise ValueError('more devices than constraint nodes')
        self.children = []
        self.shards = []
        for g, dev in enumerate(devices):
            a, b = node_shard(hi - lo + 1, g, G)
            part = copy.copy(col)
            part._node_range = (lo + a, lo + b)
            part._device = dev
            part._devices = None
            part._evaluator = None
            # the whole problem's instance constraints are appended here, not
            # by a shard
            self.shards.append(part._node_range)
            self.children.append(part)
        M = col.num_eom
        first = None
        self.evaluators = []
        for part in self.children:
            if first is not None:
                # every shard runs the same generated code as the first one
                # (the automatic geometry depends on the shard size; sums
                # accumulated in a different order would differ in the last
                # bit between neighbouring shards)
                if first.meta['persistent'] == 2:
                    # (row-stationary kernel: one equation per group anyway)
                    part._cuda_options = dict(
                        part._cuda_options, persistent='stationary',
                        tile_bufs=first.meta['tile_bufs'],
                        warps_per_block=first.meta['warps_per_block'])
                else:
                    part._cuda_options = dict(
                        part._cuda_options, groups=len(first.parts),
                        persistent=bool(first.meta['persistent']),
                        warps_per_block=first.meta['warps_per_block'],
                        min_blocks_per_sm=first.meta['min_blocks_per_sm'])
            ev = _CudaEvaluator(part)
            self.evaluators.append(ev)
            first = first or ev
        self.program = first.program
        self.meta = first.meta
        P = self.program.P
        K = M * P
        N = col.num_collocation_nodes
        owns = col._owns_instance()
        self.num_inst = col.num_instance_constraints if owns else 0
        self.nnz_inst = 0
        if self.num_inst:
            self.nnz_inst = len(col._instance_constraints_jacobian_indices()[0])
        # full-problem host vectors (the shards of a sub-range collocator
        # still index them by global node)
        self.con_full = runtime.PinnedArray(M * (N - 1) + self.num_inst)
        self.jac_full = runtime.PinnedArray((N - 1) * K + self.nnz_inst)
        for ev in self.evaluators:
            ev.handle.set_host_outputs(self.con_full, self.jac_full)
        self.nn = hi - lo
        self.con_len = M * (N - 1)
        self.jac_len = (N - 1) * K
        self._lo, self._hi, self._M, self._K, self._N = lo, hi, M, K, N

    def _check_free(self, free):
        return self.evaluators[0]._check_free(free)

    def _run(self, free, want_con, want_jac):
        for ev in self.evaluators:
            ev._push_known(free)
        for ev in self.evaluators:
            ev.handle.begin(free, want_con, want_jac)
        for ev in self.evaluators:
            ev.handle.finish()

    def constraints(self, free):
        free = self._check_free(free)
        self._run(free, True, False)
        M, N, lo, hi = self._M, self._N, self._lo, self._hi
        con = self.con_full.array
        if (lo, hi) == (0, N - 1):
            out = con.copy()
        else:
            out = np.ascontiguousarray(
                con[:M * (N - 1)].reshape(M, N - 1)[:, lo:hi]).ravel()
        if self.num_inst:
            out[self.con_len:] = self.col.eval_instance_constraints(free)
        return out

    def jacobian(self, free, refetch=False):
        free = self._check_free(free)
        if refetch:
            for ev in self.evaluators:
                ev.handle.invalidate_host_jacobian()
        self._run(free, False, True)
        jac = self.jac_full.array
        if self.num_inst:
            jac[self.jac_len:] = \
                self.col.eval_instance_constraints_jacobian_values(free)
        K, lo, hi, N = self._K, self._lo, self._hi, self._N
        if (lo, hi) == (0, N - 1):
            return jac
        return jac[lo * K:hi * K]

    def close(self):
        for ev in self.evaluators:
            ev.close()
        self.evaluators = []
        self.con_full.close()
        self.jac_full.close()


class Problem(_IpoptBase):
    """NLP facade with the reference's constructor signature
    (opty/direct_collocation.py:139-145) and cyipopt callbacks
    ``objective``, ``gradient``, ``constraints``, ``jacobianstructure``,
    ``jacobian``, ``intermediate`` (opty/direct_collocation.py:442-567).

    ``free`` is ordered ``[x_1(t_0..t_{N-1}), ..., x_n, u_1, ..., u_q,
    p_1..p_r, h]`` and the constraints ``[eom_1 @ nodes, ..., eom_M @ nodes,
    c_1..c_o]`` (opty/direct_collocation.py:116-132).
    """

    INF = 10e19

    def __init__(self, obj, obj_grad, equations_of_motion, state_symbols,
                 num_collocation_nodes, node_time_interval,
                 known_parameter_map={}, known_trajectory_map={},
                 instance_constraints=None, time_symbol=None, tmp_dir=None,
                 integration_method='backward euler', parallel=False,
                 bounds=None, show_compile_output=False, backend='cuda',
                 eom_bounds=None, device=None, cuda_options=None,
                 devices=None):

        if not equations_of_motion.has(sm.Derivative):
            raise ValueError('No time derivatives are present. The equations '
                             'of motion must be ordinary differential '
                             'equations (ODEs) or differential algebraic '
                             'equations (DAEs).')

        self.collocator = ConstraintCollocator(
            equations_of_motion, state_symbols, num_collocation_nodes,
            node_time_interval, known_parameter_map, known_trajectory_map,
            instance_constraints, time_symbol, tmp_dir, integration_method,
            parallel, show_compile_output=show_compile_output,
            backend=backend, device=device, cuda_options=cuda_options,
            devices=devices)

        self._bounds = bounds
        if eom_bounds is not None:
            bad = [k for k in eom_bounds
                   if k not in range(self.collocator.num_eom)]
            if bad:
                raise ValueError('Keys {} in eom_bounds do not correspond to '
                                 'equations of motion.'.format(bad))
        self._eom_bounds = eom_bounds

        self._obj_num_args = self._count_positional(obj)
        self._obj_grad_num_args = self._count_positional(obj_grad)
        if self._obj_num_args not in (1, 2):
            raise ValueError('The objective function can only have one or '
                             'two arguments.')
        if self._obj_grad_num_args not in (1, 2):
            raise ValueError('The gradient function can only have one or two '
                             'arguments.')
        self.obj = obj
        self.obj_grad = obj_grad

        self.con = self.collocator.generate_constraint_function()
        logger.info('Constraint function generated.')
        self.con_jac = self.collocator.generate_jacobian_function()
        logger.info('Jacobian function generated.')
        self.con_jac_rows, self.con_jac_cols = \
            self.collocator.jacobian_indices()

        self.num_free = self.collocator.num_free
        self.num_constraints = self.collocator.num_constraints

        self._generate_bound_arrays()
        self._generate_constraint_bound_arrays()

        super(Problem, self).__init__(n=self.num_free, m=self.num_constraints,
                                      lb=self.lower_bound,
                                      ub=self.upper_bound,
                                      cl=self._low_con_bounds,
                                      cu=self._upp_con_bounds)
        self.obj_value = []

    @staticmethod
    def _count_positional(func):
        code = func.__code__
        defaults = func.__defaults__
        return code.co_argcount - (len(defaults) if defaults else 0)

    # -- NLP bounds (host-side set-up, opty/direct_collocation.py:370-440) --
    def _generate_constraint_bound_arrays(self):
        low = np.zeros(self.num_constraints)
        upp = np.zeros(self.num_constraints)
        if self._eom_bounds is not None:
            per_eom = self.collocator.num_collocation_nodes - 1
            for idx, (lo, hi) in self._eom_bounds.items():
                low[idx * per_eom:(idx + 1) * per_eom] = lo
                upp[idx * per_eom:(idx + 1) * per_eom] = hi
        self._low_con_bounds = low
        self._upp_con_bounds = upp

    def _generate_bound_arrays(self):
        col = self.collocator
        N = col.num_collocation_nodes
        lb = np.full(self.num_free, -self.INF)
        ub = np.full(self.num_free, self.INF)
        n, q = col.num_states, col.num_unknown_input_trajectories
        for var, (lo, hi) in (self._bounds or {}).items():
            if var in col.state_symbols:
                sl = slice(col.state_symbols.index(var) * N,
                           (col.state_symbols.index(var) + 1) * N)
            elif var in col.unknown_input_trajectories:
                i = n + col.unknown_input_trajectories.index(var)
                sl = slice(i * N, (i + 1) * N)
            elif var in col.unknown_parameters:
                i = (n + q) * N + col.unknown_parameters.index(var)
                sl = slice(i, i + 1)
            elif col._variable_duration and var == col.time_interval_symbol:
                sl = slice(self.num_free - 1, self.num_free)
            else:
                raise ValueError('Bound variable {} not present in free '
                                 'variables.'.format(var))
            lb[sl] = lo
            ub[sl] = hi
        self.lower_bound = lb
        self.upper_bound = ub

    # -- callbacks ----------------------------------------------------------
    @property
    def bounds(self):
        """Variable bounds given at construction
        (opty/direct_collocation.py:251-255)."""
        return self._bounds

    @property
    def eom_bounds(self):
        """Equation-of-motion bounds given at construction
        (opty/direct_collocation.py:257-261)."""
        return self._eom_bounds

    def objective(self, free):
        return self.obj(free) if self._obj_num_args == 1 else \
            self.obj(self, free)

    def gradient(self, free):
        return self.obj_grad(free) if self._obj_grad_num_args == 1 else \
            self.obj_grad(self, free)

    def constraints(self, free):
        """ndarray, shape(M*(N-1) + o,)"""
        return self.con(free)

    def jacobianstructure(self):
        """(rows, cols), each int64 of shape(nnz,)"""
        return (self.con_jac_rows, self.con_jac_cols)

    def jacobian(self, free, refetch=False):
        """ndarray, shape(nnz,), aligned with ``jacobianstructure``.

        Ownership: the array is a view of a pinned host buffer owned by the
        problem, valid until the next ``jacobian`` / ``constraints`` call --
        the reference returns a view of its persistent buffer too
        (opty/direct_collocation.py:2814, 2887) and cyipopt copies it out.
        Unlike the reference, which rewrites every entry on every call,
        columns whose value cannot change between calls (literals, and
        entries that depend on known parameters only when nothing else is
        free) are copied from the device ONCE.  A consumer that modifies the
        returned array in place (e.g. scales it) must therefore ask for
        ``refetch=True`` on the next call, or construct the problem with
        ``cuda_options={'d2h_skip_constants': False}``."""
        if refetch:
            return self.con_jac(free, refetch=True)
        return self.con_jac(free)

    def intermediate(self, *args):
        self.obj_value.append(args[2])

    def solve(self, free, lagrange=[], zl=[], zu=[], respect_bounds=False):
        if respect_bounds:
            self.check_bounds_conflict(free)
        return super().solve(free, lagrange=lagrange, zl=zl, zu=zu)

    def check_bounds_conflict(self, free):
        """Raises ValueError if a bound pair is reversed or the guess violates
        its bounds (opty/direct_collocation.py:317-368)."""
        reversed_eoms = [k for k, (lo, hi) in (self._eom_bounds or {}).items()
                         if lo > hi]
        reversed_vars, outside = [], []
        for sym, (lo, hi) in (self._bounds or {}).items():
            if np.any(lo > hi):
                reversed_vars.append(sym)
            vals = self.extract_values(free, sym)
            if np.any(vals < lo) or np.any(vals > hi):
                outside.append(sym)
        if outside:
            raise ValueError('The initial guesses for {} are in conflict '
                             'with their bounds.'.format(outside))
        if reversed_eoms or reversed_vars:
            raise ValueError('The lower bound(s) for {} is (are) greater than '
                             'the upper bound(s).'.format(
                                 reversed_eoms + reversed_vars))

    # -- small conveniences used by the callbacks' consumers ----------------
    def extract_values(self, free, *variables):
        col = self.collocator
        N = col.num_collocation_nodes
        n, q = col.num_states, col.num_unknown_input_trajectories
        out = []
        for var in variables:
            if var in col.state_symbols:
                i = col.state_symbols.index(var)
                out.append(free[i * N:(i + 1) * N])
            elif var in col.unknown_input_trajectories:
                i = n + col.unknown_input_trajectories.index(var)
                out.append(free[i * N:(i + 1) * N])
            elif var in col.unknown_parameters:
                i = (n + q) * N + col.unknown_parameters.index(var)
                out.append(free[i:i + 1])
            elif col._variable_duration and var == col.time_interval_symbol:
                out.append(free[-1:])
            else:
                raise ValueError('{} is not a free variable.'.format(var))
        return np.concatenate(out)

    def _free_slices(self, variables):
        """Index ranges of ``variables`` in the free vector, layout of
        opty/direct_collocation.py:972-1002."""
        col = self.collocator
        N = col.num_collocation_nodes
        n, q = col.num_states, col.num_unknown_input_trajectories
        r = col.num_unknown_parameters
        out = []
        for var in variables:
            if var in col.state_symbols:
                i = col.state_symbols.index(var)
                out.append((i * N, (i + 1) * N))
            elif var in col.unknown_input_trajectories:
                i = n + col.unknown_input_trajectories.index(var)
                out.append((i * N, (i + 1) * N))
            elif var in col.unknown_parameters:
                i = (n + q) * N + col.unknown_parameters.index(var)
                out.append((i, i + 1))
            elif col._variable_duration and var == col.time_interval_symbol:
                out.append(((n + q) * N + r, (n + q) * N + r + 1))
            else:
                raise ValueError('{} not an unknown in this problem.'.format(
                    var))
        return out

    def fill_free(self, free, values, *variables):
        """Replaces the entries of ``free`` that belong to ``variables`` by
        ``values`` (stacked in the order of the variables), in place
        (opty/direct_collocation.py:1004-1028)."""
        idx = np.concatenate([np.arange(a, b)
                              for a, b in self._free_slices(variables)])
        free[idx] = values

    def time_vector(self, solution=None, start_time=0.0):
        """Time instances of the collocation nodes
        (opty/direct_collocation.py:1097-1132)."""
        col = self.collocator
        N = col.num_collocation_nodes
        if col._variable_duration:
            if solution is None:
                raise ValueError('Solution vector must be provided for '
                                 'variable duration.')
            h = solution[-1]
            if h <= 0.0:
                raise ValueError('Time interval must be strictly greater '
                                 'than zero.')
            if start_time >= h * (N - 1):
                raise ValueError('Start time must be less than the final '
                                 'time.')
        else:
            h = col.node_time_interval
        return np.linspace(start_time, start_time + h * (N - 1), num=N)

    def parse_free(self, free):
        col = self.collocator
        return parse_free(free, col.num_states,
                          col.num_unknown_input_trajectories,
                          col.num_collocation_nodes,
                          variable_duration=col._variable_duration)
