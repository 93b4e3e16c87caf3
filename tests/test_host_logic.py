"""CPU tests of the host side: symbol bookkeeping of the facade, the tape
(lowering + forward-mode differentiation), the emitter (run through the
test-only host harness and compared with the oracle), the compiled-module
cache and the C-ABI library's exports."""

import ctypes
import math
import os
import re
import subprocess
import sys
import tempfile
from collections import OrderedDict

import numpy as np
import pytest
import sympy as sm
import sympy.physics.mechanics as me

import cases
import workloads
from conftest import ROOT, assert_values_close, load_golden
from host_harness import host_evaluate
from opty_b200 import ConstraintCollocator, Problem, build, ir, runtime
from opty_b200 import parse_free
from opty_b200.lowering import lower_matrix
from opty_b200.program import CollocationProgram
from oracle.opty_oracle import OracleCollocator


# ---------------------------------------------------------------------------
# facade: constructor arguments, symbol sorting, discretisation
# (restates opty/tests/test_direct_collocation.py:706-790, 1073-1110)
# ---------------------------------------------------------------------------
def _msd_collocator(**kw):
    m, c, k, t = sm.symbols('m, c, k, t')
    x, v, f = [s(t) for s in sm.symbols('x, v, f', cls=sm.Function)]
    eom = sm.Matrix([x.diff() - v, m * v.diff() + c * v + k * x - f])
    par_map = OrderedDict([(m, 1.0), (c, 2.0)])
    traj_map = OrderedDict([(f, np.array([2.0, 2.5, 3.0, 3.5]))])
    col = ConstraintCollocator(eom, (x, v), 4, 0.01,
                               known_parameter_map=par_map,
                               known_trajectory_map=traj_map, time_symbol=t,
                               **kw)
    return col, (m, c, k, t, x, v, f)


def test_symbol_sorting_and_discrete_symbols():
    col, (m, c, k, t, x, v, f) = _msd_collocator()
    assert col.state_symbols == (x, v)
    assert col.state_derivative_symbols == (x.diff(t), v.diff(t))
    assert col.num_states == 2 and col.num_collocation_nodes == 4
    assert col.known_parameters == (m, c)
    assert col.unknown_parameters == (k,)
    assert col.parameters == (m, c, k)
    assert col.known_input_trajectories == (f,)
    assert col.unknown_input_trajectories == ()
    assert col.num_free == 2 * 4 + 1
    assert col.num_constraints == 2 * 3
    xi, vi, xp, vp, xn, vn, fi, fn = sm.symbols(
        'xi, vi, xp, vp, xn, vn, fi, fn', real=True)
    assert col.previous_discrete_state_symbols == (xp, vp)
    assert col.current_discrete_state_symbols == (xi, vi)
    assert col.next_discrete_state_symbols == (xn, vn)
    assert col.current_discrete_specified_symbols == (fi,)
    assert col.next_discrete_specified_symbols == (fn,)
    assert col.time_interval_symbol == sm.Symbol('h_opty', real=True)


def test_discretisation_formulas():
    col, (m, c, k, t, x, v, f) = _msd_collocator()
    xi, vi, xp, vp, xn, vn, fi, fn = sm.symbols(
        'xi, vi, xp, vp, xn, vn, fi, fn', real=True)
    h = col.time_interval_symbol
    be = sm.Matrix([(xi - xp) / h - vi,
                    m * (vi - vp) / h + c * vi + k * xi - fi])
    assert sm.simplify(col.discrete_eom - be) == sm.zeros(2, 1)
    col.integration_method = 'midpoint'
    mp = sm.Matrix([(xn - xi) / h - (vi + vn) / 2,
                    m * (vn - vi) / h + c * (vi + vn) / 2 +
                    k * (xi + xn) / 2 - (fi + fn) / 2])
    assert sm.simplify(col.discrete_eom - mp) == sm.zeros(2, 1)
    with pytest.raises(ValueError):
        col.integration_method = 'booger'


def test_known_and_unknown_order():
    """Known symbols keep the user's order, unknown ones are sorted by name
    (opty/tests/test_direct_collocation.py:2042-2088)."""
    from sympy.physics.mechanics.models import n_link_pendulum_on_cart
    me.dynamicsymbols._t = sm.Symbol('t')
    kane = n_link_pendulum_on_cart(n=3, cart_force=True, joint_torques=True)
    states = kane.q.col_join(kane.u)
    eom = kane.mass_matrix_full @ states.diff() - kane.forcing_full
    syms = {s.name: s for s in eom.free_symbols}
    funcs = {f.name: f for f in me.find_dynamicsymbols(eom)
             if hasattr(f, 'name')}
    par_map = OrderedDict([(syms['m2'], 1.0), (syms['g'], 9.81),
                           (syms['l1'], 0.5)])
    traj_map = OrderedDict([(funcs['T2'], np.ones(5)),
                            (funcs['F'], np.zeros(5))])
    col = ConstraintCollocator(eom, list(states), 5, 0.01,
                               known_parameter_map=par_map,
                               known_trajectory_map=traj_map)
    assert col.known_parameters == (syms['m2'], syms['g'], syms['l1'])
    assert [p.name for p in col.unknown_parameters] == sorted(
        n for n in syms if n not in ('m2', 'g', 'l1', 't'))
    assert col.known_input_trajectories == (funcs['T2'], funcs['F'])
    assert [f.name for f in col.unknown_input_trajectories] == ['T1', 'T3']
    rows, uniform, wrt = col._program_inputs()
    # device rows: states, unknown inputs, known inputs
    assert len(rows) == 8 + 2 + 2
    assert uniform[:3] == [syms['m2'], syms['g'], syms['l1']]
    assert uniform[-1] == col.time_interval_symbol


def test_constructor_errors():
    m, c, k, t = sm.symbols('m, c, k, t')
    x, v, f = [s(t) for s in sm.symbols('x, v, f', cls=sm.Function)]
    eom = sm.Matrix([x.diff() - v, m * v.diff() + c * v + k * x - f])
    with pytest.raises(ValueError):
        ConstraintCollocator(eom, (x, v), 4, 0.01, time_symbol=t,
                             backend='fortran')
    with pytest.raises(NotImplementedError):
        ConstraintCollocator(eom, (x, v), 4, 0.01, time_symbol=t,
                             backend='cython')
    with pytest.raises(ValueError):
        ConstraintCollocator(eom, (x, v), 4, 0.01, time_symbol=t,
                             integration_method='simpson')
    with pytest.raises(ValueError):
        ConstraintCollocator(eom, (x, x), 4, 0.01, time_symbol=t)
    with pytest.raises(ValueError):   # wrong length of a known trajectory
        ConstraintCollocator(eom, (x, v), 4, 0.01, time_symbol=t,
                             known_trajectory_map={f: np.ones(3)})
    with pytest.raises(ValueError):   # state derivative without a state
        ConstraintCollocator(eom, (x,), 4, 0.01, time_symbol=t)
    with pytest.raises(ValueError):
        ConstraintCollocator(eom, (x, v), 4, 0.01, time_symbol=t,
                             cuda_options={'no_such_option': 1})
    with pytest.raises(ValueError):   # Problem needs time derivatives
        Problem(lambda fr: 0.0, lambda fr: fr, sm.Matrix([x - v]), (x, v), 4,
                0.01, time_symbol=t)


def test_implicit_known_trajectory_symbols():
    """opty/tests/test_direct_collocation.py:83-126."""
    case = cases.implicit_known_trajectory()
    col = ConstraintCollocator(*case.collocator_args(),
                               **case.collocator_kwargs())
    keys = list(case.traj_map.keys())
    assert col._deriv_in_knw_traj
    assert col.known_input_trajectories == tuple(keys)
    assert [f.name for f in col.unknown_input_trajectories] == ['f']
    xi = sm.Symbol('xi', real=True)
    vi = sm.Symbol('vi', real=True)
    assert col.current_known_discrete_specified_symbols == (
        sm.Symbol('domegai_dvi', real=True),
        sm.Function('omegai', real=True)(vi),
        sm.Symbol('si', real=True),
        sm.Function('thetai', real=True)(xi),
        sm.Symbol('dthetai_dxi', real=True))
    rules = col._chain_rules()
    assert (sm.Function('thetai', real=True)(xi), xi,
            sm.Symbol('dthetai_dxi', real=True)) in rules
    # a function of two variables is rejected
    x, v = case.states
    th2 = sm.Function('theta', real=True)(x, v)
    eom = case.eom.subs(sm.Function('theta', real=True)(x), th2)
    with pytest.raises(ValueError):
        ConstraintCollocator(eom, case.states, 4, case.h,
                             known_parameter_map=case.par_map,
                             known_trajectory_map={th2: np.ones(4)},
                             time_symbol=case.t)


def test_instance_constraint_indexing():
    case = cases.pendulum_variable_duration()
    col = ConstraintCollocator(*case.collocator_args(),
                               **case.collocator_kwargs())
    rows, cols = col._instance_constraints_jacobian_indices()
    assert list(rows) == [6, 7, 8, 9]
    assert list(cols) == [0, 3, 4, 7]
    np.testing.assert_allclose(col.eval_instance_constraints(case.free),
                               case.expected_con[-4:])
    np.testing.assert_allclose(
        col.eval_instance_constraints_jacobian_values(case.free),
        [1.0, 3.0, 4.0, 5.0])
    th = sm.Function('theta')
    h = case.h
    with pytest.raises(ValueError):
        ConstraintCollocator(case.eom, case.states, 4, h,
                             known_parameter_map=case.par_map,
                             instance_constraints=(th(7 * h),),
                             time_symbol=case.t)


def test_parse_free():
    """opty/tests/test_utils.py parse_free cases: q = 0, 1, 2; fixed and
    variable duration."""
    n, N = 2, 3
    free = np.arange(n * N + 2 * N + 2 + 1, dtype=float)
    x, u, p, h = parse_free(free, n, 2, N, variable_duration=True)
    assert x.shape == (2, 3) and u.shape == (2, 3)
    assert list(p) == [12.0, 13.0] and h == 14.0
    x, u, p = parse_free(free[:n * N + N + 1], n, 1, N)
    assert u.shape == (3,) and list(p) == [9.0]
    x, u, p = parse_free(free[:n * N + 2], n, 0, N)
    assert u is None and list(p) == [6.0, 7.0]
    assert np.shares_memory(x, free)


# ---------------------------------------------------------------------------
# tape: lowering and differentiation against SymPy
# ---------------------------------------------------------------------------
def test_tape_lowering_and_forward_jacobian_match_sympy():
    a, b, c, d = sm.symbols('a b c d', real=True)
    exprs = [
        -a * b - c * d + a / 2,
        a * b / (c * d) + sm.sqrt(c) * d**3,
        sm.sin(a) * sm.cos(b) + sm.tan(a * b) - sm.exp(-a * a),
        sm.log(c) * sm.atan(a) + sm.asin(a / 3) + sm.acos(b / 5),
        sm.sinh(a) + sm.cosh(b) * sm.tanh(c) + c**sm.Rational(3, 2),
        sm.atan2(a, c) + sm.Abs(b) + c**d + (a + b)**-2,
        sm.Max(a, b) * sm.Min(c, d) + sm.sign(b) * a,
        sm.Piecewise((a**2, a > b), (b * c, True)) + sm.Heaviside(b) * d,
    ]
    syms = [a, b, c, d]
    rng = np.random.default_rng(0)
    for use_cse in (True, False):
        T = ir.Tape()
        leaf = {s: T.vin(i) for i, s in enumerate(syms)}
        outs = lower_matrix(T, leaf, exprs, use_sympy_cse=use_cse)
        rows = ir.forward_jacobian(T, outs, [leaf[s] for s in syms])
        for _ in range(3):
            vals = [0.3 + rng.random(), -1.2 + rng.random(),
                    1.5 + rng.random(), 0.7 + rng.random()]
            sub = dict(zip(syms, vals))
            got = T.evaluate(outs, vals, [])
            for e, g in zip(exprs, got):
                assert math.isclose(g, float(e.subs(sub)), rel_tol=1e-13,
                                    abs_tol=1e-13)
            for e, row in zip(exprs, rows):
                for k, s in enumerate(syms):
                    want = float(e.diff(s).subs(sub).evalf())
                    node = row.get(k, T.zero)
                    have = T.evaluate([node], vals, [])[0]
                    assert math.isclose(have, want, rel_tol=1e-12,
                                        abs_tol=1e-12), (e, s)


def test_tape_simplifications():
    T = ir.Tape()
    x, y = T.vin(0), T.vin(1)
    assert T.mul(x, T.zero) == T.zero
    assert T.mul(T.one, x) == x
    assert T.add(x, T.zero) == x
    assert T.sub(x, x) == T.zero
    assert T.neg(T.neg(x)) == x
    assert T.add(x, y) == T.add(y, x)          # hash-consed, commutative
    assert T.mul(x, y) == T.mul(y, x)
    assert T.powi(x, 2) == T.mul(x, x)
    assert T.op[T.div(x, T.const(4.0))] == ir.MUL   # exact power-of-two scale
    assert T.op[T.div(x, T.const(3.0))] == ir.DIV
    u = T.uin(0)
    assert not T.varying[T.mul(u, u)] and T.varying[T.mul(u, x)]


def test_program_classifies_entries_and_partitions_rows():
    w = workloads.n_link_pendulum(3, 50)
    col = ConstraintCollocator(*w.collocator_args(), **w.collocator_kwargs())
    rows, uniform, wrt = col._program_inputs()
    prog = CollocationProgram(list(col.discrete_eom), rows, uniform, wrt)
    assert (prog.M, prog.P, prog.R) == (8, 18, 9)
    kinds = prog.entry_kind()
    assert len(kinds) == 8 * 18
    assert (prog.num_literal_entries + prog.num_invariant_entries +
            prog.num_varying_entries) == 8 * 18
    # kinematic rows q' - u = 0 only hold literals and +-1/h
    assert all(k in (0, 1) for k in kinds[:4 * 18])
    for G in (1, 2, 3, 5, 8):
        parts = prog.partition_rows(G, col_align=2)
        assert parts[0][0] == 0 and parts[-1][1] == prog.M
        assert len(parts) <= G
        for (a0, a1), (b0, b1) in zip(parts, parts[1:]):
            assert a1 == b0 and a0 < a1
        assert all((r0 * prog.P) % 2 == 0 for r0, _ in parts)


# ---------------------------------------------------------------------------
# emitter, run on the CPU through the test-only host harness, vs the oracle
# ---------------------------------------------------------------------------
HARNESS_WORKLOADS = [
    ('cfg1_pendulum_swing_up_N51', lambda: workloads.pendulum_swing_up(51)),
    ('cfg3_vyasarayani2011_N101_odd',
     lambda: workloads.vyasarayani2011(101, seed=5)),
    ('cfg4_standin_pendulum4_torques_N200',
     lambda: workloads.n_link_pendulum_torques(4, 200)),
    ('cfg2_small_pendulum10_N40',
     lambda: workloads.n_link_pendulum(10, 40, seed=7)),
]


@pytest.mark.parametrize('name,make', HARNESS_WORKLOADS,
                         ids=[f[0] for f in HARNESS_WORKLOADS])
@pytest.mark.parametrize('groups', [1, 3])
def test_emitted_code_matches_reference_golden(name, make, groups):
    gold = load_golden(name)
    w = make()
    col = ConstraintCollocator(*w.collocator_args(), **w.collocator_kwargs(),
                               cuda_options={'groups': groups})
    free = w.free(col.num_free)
    con, jac = host_evaluate(col, free)
    M = col.num_eom
    nn = col.num_collocation_nodes - 1
    P = len(jac) // (nn * M)
    assert_values_close(con, gold['con'][:M * nn])
    assert_values_close(jac, gold['jac'][:nn * M * P], row_len=P)


@pytest.mark.parametrize('name,make', HARNESS_WORKLOADS[2:],
                         ids=[f[0] for f in HARNESS_WORKLOADS[2:]])
def test_narrow_staging_buffers_and_schedule_options(name, make):
    """Rows wider than the staging buffer are cut into phases that pair the
    partials with respect to a state at both nodes of the stencil; tight
    rematerialisation budgets, the tape's own association order and the
    unscheduled emission order must all still produce every entry."""
    gold = load_golden(name)
    for opts in ({'groups': 2, 'tile_cols': 12, 'live_budget': 12},
                 {'groups': 3, 'tile_cols': 20, 'reassociate': False,
                  'tile_bufs': 1},
                 {'groups': 2, 'schedule': False, 'tile_cols': 16},
                 {'groups': 1, 'live_budget': 8, 'remat_cost': 60},
                 {'groups': 2, 'load_ahead': 40, 'schedule': True}):
        w = make()      # the seeded free vector continues the workload's rng
        col = ConstraintCollocator(*w.collocator_args(),
                                   **w.collocator_kwargs(), cuda_options=opts)
        free = w.free(col.num_free)
        from host_harness import _prepare_without_nvcc
        prog, source, meta = _prepare_without_nvcc(col)
        cols = [g['cols'] for g in meta['groups']]
        assert cols[0][0] == 0 and cols[-1][1] == prog.K
        assert all(a[1] == b[0] for a, b in zip(cols, cols[1:]))
        if 'tile_cols' in opts:
            assert max(g['phases'] for g in meta['groups']) > \
                (cols[0][1] - cols[0][0]) // prog.P
        con, jac = host_evaluate(col, free)
        M = col.num_eom
        nn = col.num_collocation_nodes - 1
        assert_values_close(con, gold['con'][:M * nn])
        assert_values_close(jac, gold['jac'][:nn * M * prog.P],
                            row_len=prog.P)


def test_row_phases_cover_every_column_once():
    from opty_b200.codegen import row_phases
    for col0, ncols, tile, pair, even in (
            (0, 46, 46, 22, True), (92, 206, 52, 102, True),
            (0, 206, 52, 102, True), (86, 86, 52, 42, True),
            (0, 27, 20, 10, False), (54, 54, 20, None, True),
            (0, 45, 64, 19, False), (10, 30, 8, 15, True)):
        phases = row_phases(col0, ncols, tile, pair, even)
        seen = []
        for ph in phases:
            assert sum(w for _, w in ph) <= tile
            for c0, w in ph:
                assert w >= 1
                if even:
                    assert w % 2 == 0
                seen.extend(range(c0, c0 + w))
        assert sorted(seen) == list(range(col0, col0 + ncols)), (col0, ncols)
    # wide rows keep the two partials of a state in one phase
    phases = row_phases(0, 206, 52, 102, True)
    for ph in phases[:-1]:
        (a0, wa), (b0, wb) = ph[0], ph[1]
        assert b0 - a0 == 102 and wa == wb


def test_scheduler_reduces_live_values_and_keeps_every_operation():
    """schedule.py on the heaviest equation of the 10-link pendulum: every
    output is produced once, no value is read before it exists, and the peak
    number of live values is a fraction of the plain emission order's."""
    from opty_b200 import schedule
    w = workloads.n_link_pendulum(10, 40, seed=7)
    col = ConstraintCollocator(*w.collocator_args(), **w.collocator_kwargs())
    prog = col._build_program()
    T = prog.tape
    j = prog.M // 2 + 1
    outs = [prog.con[j]] + list(prog.jac[j])
    plain = schedule.plain_order(T, outs)
    for kw in ({}, {'reassociate': False}, {'live_budget': 16},
               {'remat_cost': 0}):
        sc = schedule.schedule_body(T, outs, **kw)
        dag = sc.dag
        assert sorted(e[1] for e in sc.events if e[0] == schedule.OUT) == \
            list(range(len(outs)))
        have = set()
        started = set()
        for e in sc.events:
            if e[0] in (schedule.OP, schedule.LOAD):
                assert all(o in have for o in dag.rops[e[1]])
                assert e[1] not in have
                have.add(e[1])
            elif e[0] == schedule.ACC:
                s_, j_ = e[1], e[2]
                assert all(o in have for o in dag.term_rops[s_][j_])
                assert (s_ not in started) == bool(e[3])
                started.add(s_)
                if all(x in have for tr in dag.term_rops[s_] for x in tr):
                    have.add(s_)
            else:
                assert all(o in have for o in dag.out_rops[e[1]])
        assert schedule.peak_live(dag, sc.events) == sc.peak_live
        assert sc.peak_live < 0.6 * plain.peak_live
    tight = schedule.schedule_body(T, outs, live_budget=16)
    loose = schedule.schedule_body(T, outs, live_budget=200)
    assert tight.peak_live < loose.peak_live
    assert tight.num_ops >= loose.num_ops


@pytest.mark.parametrize('case', cases.product_cases(),
                         ids=lambda c: c.name)
def test_emitted_code_known_answers(case):
    col = ConstraintCollocator(*case.collocator_args(),
                               **case.collocator_kwargs())
    con, jac = host_evaluate(col, case.free)
    np.testing.assert_allclose(con, case.expected_con[:len(con)], rtol=1e-12,
                               atol=1e-9)
    np.testing.assert_allclose(jac, case.expected_jac[:len(jac)], rtol=1e-12,
                               atol=1e-9)


# ---------------------------------------------------------------------------
# native pieces: module compilation + cache, C-ABI exports
# ---------------------------------------------------------------------------
def test_module_compiles_for_sm100a_and_is_cached():
    w = workloads.vyasarayani2011(101, seed=5)
    with tempfile.TemporaryDirectory() as tmp:
        col = ConstraintCollocator(*w.collocator_args(),
                                   **w.collocator_kwargs(), tmp_dir=tmp)
        first = col.prepare_module()
        assert not first.cache_hit and len(first.cubin) > 1000
        second = col.prepare_module()
        assert second.cache_hit and second.cubin == first.cubin
        sass = subprocess.run(['cuobjdump', '-sass', first.cubin_path],
                              capture_output=True, text=True).stdout
        assert 'sm_100a' in sass
        assert 'UTMALDG' in sass          # TMA tile load of the trajectory
        assert 'UTMASTG' in sass          # TMA tile store of the Jacobian
        # a module emitted without TMA has neither
        col2 = ConstraintCollocator(
            *w.collocator_args(), **w.collocator_kwargs(), tmp_dir=tmp,
            cuda_options={'tma_load': False, 'tma_store': False})
        plain = col2.prepare_module()
        sass2 = subprocess.run(['cuobjdump', '-sass', plain.cubin_path],
                               capture_output=True, text=True).stdout
        assert 'UTMASTG' not in sass2 and 'UTMALDG' not in sass2


def test_sharded_modules_compile_for_sm100a_and_carry_their_geometry():
    """Compile-only checks (no GPU): a sharded build yields one module per
    group range, only the first one holds the auxiliary kernels, and every
    module exports the ``opty_module_info`` table the runtime reads its
    kernel geometry from (nothing of it crosses the C-ABI)."""
    w = workloads.n_link_pendulum(10, 40, seed=7)
    with tempfile.TemporaryDirectory() as tmp:
        col = ConstraintCollocator(
            *w.collocator_args(), **w.collocator_kwargs(), tmp_dir=tmp,
            cuda_options={'groups': 6, 'compile_shards': 3,
                          'schedule': True, 'min_blocks_per_sm': 8})
        pm = col.prepare_module()
        extra = pm.meta['extra_modules']
        assert len(extra) == 2
        ranges = [pm.meta['group_range']] + [e['group_range'] for e in extra]
        assert ranges[0][0] == 0 and ranges[-1][1] == len(pm.parts)
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        syms = subprocess.run(['cuobjdump', '-elf', pm.cubin_path],
                              capture_output=True, text=True).stdout
        assert 'opty_module_info' in syms and 'opty_colloc_pre' in syms
        for e in extra:
            syms = subprocess.run(['cuobjdump', '-elf', e['cubin_path']],
                                  capture_output=True, text=True).stdout
            assert 'opty_colloc_eval' in syms
            assert 'opty_module_info' in syms
            assert 'opty_colloc_pre' not in syms
            assert 'opty_colloc_inv' not in syms
        # registers: the scheduled bodies fit 16 resident warps per SM
        # without spilling
        res = subprocess.run(['cuobjdump', '-res-usage', pm.cubin_path],
                             capture_output=True, text=True).stdout
        line = [ln for ln in res.splitlines() if 'REG:' in ln][0]
        regs = int(re.search(r'REG:(\d+)', line).group(1))
        assert regs <= 128 and 'LOCAL:0' in line


def test_stationary_schedule_covers_every_item_once():
    """Static work assignment of the row-stationary kernel: every (group, node
    tile) item lands in exactly one slot, in both layouts; the strided layout
    gives a group whole slots and walks its tiles round robin over them."""
    from opty_b200.codegen import stationary_schedule
    for costs, n_tiles, n_slots in (
            ([29.0, 45.5, 45.1, 44.8, 44.1, 43.1, 41.8, 40.3, 38.0, 34.3,
              31.5], 40, 148),
            ([10.0, 1.0, 1.0, 9.0], 7, 5),
            ([3.0], 1, 148),
            ([5.0, 4.0, 0.5], 313, 16)):
        for strided in (True, False):
            sched = stationary_schedule(costs, n_tiles, n_slots,
                                        strided=strided)
            assert len(sched) == n_slots
            seen = []
            for slot in sched:
                for g, t0, nt in slot:
                    seen.extend((g, t) for t in range(t0, t0 + nt))
            assert sorted(seen) == [(g, t) for g in range(len(costs))
                                    for t in range(n_tiles)]
            if strided and len(costs) <= n_slots:
                heavy = [g for g in range(len(costs))
                         if costs[g] >= 0.25 * max(costs)]
                for g in heavy:
                    slots = [i for i, sl in enumerate(sched)
                             if any(sg[0] == g for sg in sl)]
                    # contiguous block of slots owned by the group alone
                    # (light groups may be dealt into them afterwards)
                    assert slots == list(range(slots[0], slots[-1] + 1))
                    S = len(slots)
                    for i, sl in zip(slots, (sched[i] for i in slots)):
                        tiles = [t for sg in sl if sg[0] == g
                                 for t in range(sg[1], sg[1] + sg[2])]
                        assert tiles == list(range(i - slots[0], n_tiles, S))


def test_row_stationary_module_compiles_for_sm100a():
    """Compile-only (no GPU): the automatic choice at the 10-link pendulum is
    the row-stationary kernel -- one group per dynamic equation, the eleven
    ``x' = v`` rows as constant runs -- within 227 KB of shared memory and 255
    registers without spilling; its SASS holds TMA tile loads / stores and the
    bulk copies of the constant runs; a problem whose input windows do not
    fit falls back to the grid kernel."""
    w = workloads.n_link_pendulum(10, 40, seed=7)
    with tempfile.TemporaryDirectory() as tmp:
        col = ConstraintCollocator(*w.collocator_args(),
                                   **w.collocator_kwargs(), tmp_dir=tmp)
        pm = col.prepare_module()
        meta = pm.meta
        assert meta['persistent'] == 2 and meta['warps_per_block'] == 8
        assert meta['const_rows'] == list(range(11))
        assert meta['num_groups'] == 11
        assert [g['rows'] for g in meta['groups']] == \
            [[j, j + 1] for j in range(11, 22)]
        assert 0 < meta['smem_bytes'] <= 227 * 1024
        res = subprocess.run(['cuobjdump', '-res-usage', pm.cubin_path],
                             capture_output=True, text=True).stdout
        lines = res.splitlines()
        line = [lines[i + 1] for i, ln in enumerate(lines)
                if 'opty_colloc_eval' in ln][0]
        assert int(re.search(r'REG:(\d+)', line).group(1)) <= 255
        assert 'LOCAL:0' in line
        sass = subprocess.run(['cuobjdump', '-sass', pm.cubin_path],
                              capture_output=True, text=True).stdout
        assert 'UTMALDG' in sass and 'UTMASTG' in sass and 'UBLKCP' in sass
        # forcing the grid kernel still works, and so does the fall-back
        col2 = ConstraintCollocator(
            *w.collocator_args(), **w.collocator_kwargs(), tmp_dir=tmp,
            cuda_options={'persistent': False})
        assert col2.prepare_module().meta['persistent'] == 0
        col3 = ConstraintCollocator(
            *w.collocator_args(), **w.collocator_kwargs(), tmp_dir=tmp,
            cuda_options={'persistent': 'stationary', 'tile_bufs': 2})
        with pytest.raises(ValueError, match='shared memory'):
            col3.prepare_module()


def test_setup_index_skips_the_symbolic_work_and_tracks_its_inputs():
    """The set-up cache is keyed by the inputs of the symbolic work (discrete
    EOM, symbol layout, options): the same problem comes back from the index
    without lowering; a changed equation, option or node count does not."""
    with tempfile.TemporaryDirectory() as tmp:
        def prepared(w, **opts):
            col = ConstraintCollocator(*w.collocator_args(),
                                       **w.collocator_kwargs(), tmp_dir=tmp,
                                       cuda_options=opts)
            return col.prepare_module()
        w = workloads.vyasarayani2011(101, seed=5)
        first = prepared(w)
        assert not first.index_hit and first.source
        again = prepared(workloads.vyasarayani2011(101, seed=5))
        assert again.index_hit and again.source is None
        assert again.cubin == first.cubin and again.meta['K'] == first.meta['K']
        assert again.program.P == first.program.P
        assert [tuple(p) for p in again.parts] == list(first.parts)
        # different options, node count or equations: no hit
        assert not prepared(workloads.vyasarayani2011(101, seed=5),
                            tile_cols=14).index_hit
        assert not prepared(workloads.vyasarayani2011(5000)).index_hit
        w2 = workloads.vyasarayani2011(101, seed=5)
        w2.eom = w2.eom + sm.Matrix([0, w2.states[0]])
        other = prepared(w2)
        assert not other.index_hit and other.cubin != first.cubin
        # known values are not part of the module: same index entry
        w3 = workloads.pendulum_swing_up(51)
        a = prepared(w3)
        w4 = workloads.pendulum_swing_up(51)
        for k in w4.known_parameter_map:
            w4.known_parameter_map[k] *= 2.0
        b = prepared(w4)
        assert not a.index_hit and b.index_hit
        # opting out
        assert not prepared(workloads.vyasarayani2011(101, seed=5),
                            use_index=False).index_hit


def test_compile_failure_raises_import_error():
    """Build failures surface as ImportError with the compiler's stderr
    (opty/utils.py:909-916, pinned by opty/tests/test_utils.py:333-336)."""
    with tempfile.TemporaryDirectory() as tmp:
        with pytest.raises(ImportError) as err:
            build.compile_module('this is not CUDA;', build.module_flags(),
                                 cache_dir=tmp)
        assert 'STDERR' in str(err.value)


def test_c_abi_library_exports_every_declared_symbol():
    lib = runtime.load_library()
    header = open(os.path.join(ROOT, 'include', 'opty_b200.h')).read()
    declared = set(re.findall(r'\b(opty_[a-z0-9_]+)\s*\(', header))
    assert declared == set(runtime.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.opty_b200_abi_version() == runtime.ABI_VERSION


def test_c_abi_config_struct_layout_matches_header():
    src = ('#include "opty_b200.h"\n#include <stdio.h>\n#include <stddef.h>\n'
           'int main(void){printf("%zu %zu %zu", sizeof(opty_colloc_cfg), '
           'offsetof(opty_colloc_cfg, jac_tail), '
           'offsetof(opty_colloc_cfg, h)); return 0;}')
    with tempfile.TemporaryDirectory() as tmp:
        c = os.path.join(tmp, 'sz.c')
        open(c, 'w').write(src)
        exe = os.path.join(tmp, 'sz')
        subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), '-o', exe,
                        c], check=True)
        out = subprocess.run([exe], capture_output=True, text=True).stdout
    size, off_g, off_h = (int(v) for v in out.split())
    assert size == ctypes.sizeof(runtime.ColloCfg)
    assert off_g == runtime.ColloCfg.jac_tail.offset
    assert size <= 96      # the problem in the reference's notation, nothing else
    assert off_h == runtime.ColloCfg.h.offset


def test_no_cpu_fallback_without_a_gpu():
    """The product path fails loudly when no CUDA device is present."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    w = workloads.vyasarayani2011(101, seed=5)
    col = ConstraintCollocator(*w.collocator_args(), **w.collocator_kwargs())
    with pytest.raises(RuntimeError) as err:
        col.generate_constraint_function()
    assert 'CUDA' in str(err.value)
    with pytest.raises(RuntimeError):
        col.jacobian_indices()


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, 'opty_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, fn)).read()
                assert 'oracle' not in text.lower(), fn
                assert 'host_harness' not in text, fn


# ---------------------------------------------------------------------------
# objective / gradient (opty/utils.py:329-470): symbolic half on the CPU
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('name', [k for k in cases.objective_cases()
                                  if not k.startswith('tracking')])
def test_objective_lowering_matches_reference(name):
    """Integrand, partials, quadrature weights and the parameter-only terms
    against values produced by the reference's create_objective_function
    (tests/golden/make_golden_objective.py); the emitted integrand code runs
    on the CPU through the host shim."""
    from host_harness import host_objective
    gold = load_golden('objective_cases')
    c = cases.objective_cases()[name]()
    obj, grad = host_objective(
        c['objective'], c['states'], c['inputs'], c['params'], c['N'],
        c['h'], integration_method=c['method'], time_symbol=c['t'])
    np.testing.assert_allclose(obj(c['free']), gold[name + '_value'],
                               rtol=1e-13)
    np.testing.assert_allclose(grad(c['free']), gold[name + '_grad'],
                               rtol=1e-12, atol=1e-15)


def test_objective_rejects_what_the_reference_rejects():
    from opty_b200.utils import create_objective_function
    t = sm.symbols('t')
    x = sm.Function('x')(t)
    with pytest.raises(NotImplementedError):
        create_objective_function(sm.Integral(x ** 2, t), [x], [], [], 10, 1.0,
                                  integration_method='not_existing_method',
                                  time_symbol=t)
    with pytest.raises(NotImplementedError):
        create_objective_function(sm.Integral(x ** 2, (t, 0, 1)), [x], [], [],
                                  10, 1.0, time_symbol=t)
    with pytest.raises(NotImplementedError):
        create_objective_function(
            sm.Integral(sm.Integral(x, t) * x, t), [x], [], [], 10, 1.0,
            time_symbol=t)
