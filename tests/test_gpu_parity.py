"""Parity tests proper: the CUDA path -- called through the facade, which
calls through the C-ABI of include/opty_b200.h -- against the CPU oracle on
the same seeded inputs, against the golden vectors produced by the reference
itself, against the reference's hand-computed known answers, and through
size-independent properties at the full BASELINE config 2 size.

Bar: ``jacobianstructure`` bit-exact (int64), values within 1e-10 relative
(criterion spelled out in conftest.assert_values_close)."""

import hashlib

import numpy as np
import pytest

import cases
import gpu_variants
import workloads
from conftest import (assert_values_close, canonical_triplets,
                      load_golden)
from opty_b200 import ConstraintCollocator, Problem, runtime
from oracle.opty_oracle import OracleCollocator

pytestmark = pytest.mark.gpu


def _digest(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


def _collocator(w, **kw):
    return ConstraintCollocator(*w.collocator_args(),
                                **w.collocator_kwargs(), **kw)


def _eom_sizes(col):
    nn = col.num_collocation_nodes - 1
    M = col.num_eom
    return nn, M


# ---------------------------------------------------------------------------
# known answers (hand-computed in the reference's tests)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('case', cases.product_cases(),
                         ids=lambda c: c.name)
def test_known_answers(case):
    col = ConstraintCollocator(*case.collocator_args(),
                               **case.collocator_kwargs())
    con = col.generate_constraint_function()(case.free)
    jac = np.array(col.generate_jacobian_function()(case.free))
    rows, cols = col.jacobian_indices()
    assert con.shape == case.expected_con.shape
    assert jac.shape == case.expected_jac.shape
    np.testing.assert_allclose(con, case.expected_con, rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(jac, case.expected_jac, rtol=1e-12, atol=1e-9)
    assert rows.dtype == np.int64 and cols.dtype == np.int64
    assert len(rows) == len(cols) == len(jac)
    if case.expected_rows is not None:
        assert np.array_equal(rows, case.expected_rows)
        assert np.array_equal(cols, case.expected_cols)
    col.close()


# ---------------------------------------------------------------------------
# golden vectors produced by the reference (tests/golden/make_golden.py)
# ---------------------------------------------------------------------------
GOLDEN = [
    ('cfg1_pendulum_swing_up_N51', lambda: workloads.pendulum_swing_up(51)),
    ('cfg3_vyasarayani2011_N5000', lambda: workloads.vyasarayani2011(5000)),
    ('cfg3_vyasarayani2011_N101_odd',
     lambda: workloads.vyasarayani2011(101, seed=5)),
    ('cfg4_standin_pendulum4_torques_N200',
     lambda: workloads.n_link_pendulum_torques(4, 200)),
    ('cfg2_small_pendulum10_N40',
     lambda: workloads.n_link_pendulum(10, 40, seed=7)),
]


@pytest.mark.parametrize('name,make', GOLDEN, ids=[g[0] for g in GOLDEN])
def test_matches_reference_golden(name, make):
    gold = load_golden(name)
    w = make()
    col = _collocator(w)
    free = w.free(col.num_free)
    assert np.array_equal(free, gold['free'])
    con = col.generate_constraint_function()(free)
    jac = np.array(col.generate_jacobian_function()(free))
    rows, cols = col.jacobian_indices()
    # jacobianstructure: bit-exact
    assert np.array_equal(rows, gold['rows'])
    assert np.array_equal(cols, gold['cols'])
    nn, M = _eom_sizes(col)
    P = col._evaluator.program.P
    assert con.shape == gold['con'].shape and jac.shape == gold['jac'].shape
    assert_values_close(con[:M * nn], gold['con'][:M * nn])
    assert_values_close(jac[:nn * M * P], gold['jac'][:nn * M * P], row_len=P)
    # instance-constraint parts (host lambdify in both implementations)
    np.testing.assert_allclose(con[M * nn:], gold['con'][M * nn:],
                               rtol=1e-13, atol=0)
    np.testing.assert_allclose(jac[nn * M * P:], gold['jac'][nn * M * P:],
                               rtol=1e-13, atol=0)
    col.close()


# ---------------------------------------------------------------------------
# BASELINE config 2 at full size
# ---------------------------------------------------------------------------
@pytest.fixture(scope='module')
def config2():
    w = workloads.n_link_pendulum(10, 10000)
    col = _collocator(w)
    free = w.free(col.num_free)
    con_f = col.generate_constraint_function()
    jac_f = col.generate_jacobian_function()
    con = con_f(free)
    jac = np.array(jac_f(free))
    yield w, col, free, con, jac
    col.close()


def test_config2_against_reference_sample_and_structure_digest(config2):
    w, col, free, con, jac = config2
    gold = load_golden('cfg2_pendulum10_N10000')
    assert _digest(free) == str(gold['free_sha256'])
    nn, M = _eom_sizes(col)
    K = M * col._evaluator.program.P
    assert len(jac) == int(gold['nnz']) == 10118988
    nodes = gold['nodes']
    assert_values_close(con.reshape(M, nn)[:, nodes], gold['con'])
    assert_values_close(jac.reshape(nn, K)[nodes].ravel(),
                        gold['jac'].ravel(), row_len=K // M)
    rows, cols = col.jacobian_indices()
    assert rows.dtype == np.int64 and cols.dtype == np.int64
    # all 10 118 988 COO entries bit-equal to the reference's
    assert _digest(rows) == str(gold['rows_sha256'])
    assert _digest(cols) == str(gold['cols_sha256'])


def _relative_error_report(name, got, want, row_len=None):
    """Max pure relative error over ALL non-zero entries, and max error in
    units of the row scale; printed so that the margin to the 1e-10 bar is
    visible in the test log."""
    nz = want != 0.0
    rel = np.abs(got[nz] - want[nz]) / np.abs(want[nz])
    if row_len is None:
        scale = np.full(want.shape, np.max(np.abs(want)))
    else:
        scale = np.repeat(np.max(np.abs(want.reshape(-1, row_len)), axis=1),
                          row_len)
    srel = np.abs(got - want)[scale > 0] / scale[scale > 0]
    worst = int(np.argmax(rel))
    print('\n[parity] {}: {} non-zero entries, max relative error {:.3e} '
          '(entry magnitude {:.3e}, row scale {:.3e}), 99.99th percentile '
          '{:.3e}, max error / row scale {:.3e}, bit-identical {:.1f} %'
          .format(name, int(nz.sum()), rel.max(),
                  np.abs(want[nz][worst]), scale[nz][worst],
                  np.quantile(rel, 0.9999), srel.max(),
                  100.0 * np.mean(got == want)))
    return rel.max(), srel.max()


def test_config2_against_oracle_full_size(config2):
    """Default build (FMA contraction on, scheduled bodies, re-associated
    sums) against the oracle at all 10 118 988 entries.  The criterion is the
    one of conftest.assert_values_close; on top of it the maximum pure
    relative error over all non-zero entries is printed and bounded: entries
    that are not formed by cancellation (|entry| >= 1e-6 of their row's
    scale) must be within 1e-10 relative, every entry within 1e-14 of its
    row's scale."""
    w, col, free, con, jac = config2
    orc = OracleCollocator(*w.collocator_args(), **w.collocator_kwargs())
    ocon = orc.constraints(free)
    ojac = orc.jacobian(free)
    P = col._evaluator.program.P
    assert_values_close(con, ocon)
    assert_values_close(jac, ojac, row_len=P)
    rel_j, srel_j = _relative_error_report('config 2 Jacobian', jac, ojac, P)
    rel_c, srel_c = _relative_error_report('config 2 residuals', con, ocon)
    assert srel_j < 1e-14 and srel_c < 1e-14
    scale = np.repeat(np.max(np.abs(ojac.reshape(-1, P)), axis=1), P)
    solid = np.abs(ojac) >= 1e-6 * scale
    solid &= ojac != 0.0
    assert np.max(np.abs(jac[solid] - ojac[solid]) /
                  np.abs(ojac[solid])) < 1e-10
    solid = np.abs(ocon) >= 1e-6 * np.max(np.abs(ocon))
    assert np.max(np.abs(con[solid] - ocon[solid]) /
                  np.abs(ocon[solid])) < 1e-10


def test_config2_properties(config2):
    w, col, free, con, jac = config2
    nn, M = _eom_sizes(col)
    P = col._evaluator.program.P
    K = M * P
    N = col.num_collocation_nodes
    con_f = col.generate_constraint_function()
    jac_f = col.generate_jacobian_function()

    # (1) deterministic: a second evaluation gives identical bits
    f2 = free.copy()
    f2[0] += 1.0
    con_f(f2)                      # move away ...
    assert np.array_equal(con_f(free), con)      # ... and back
    assert np.array_equal(np.array(jac_f(free)), jac)

    # (2) locality: node i only depends on trajectory columns i and i+1.
    # Perturbing column c changes exactly the node blocks c-1 and c.
    c = 4321
    f3 = free.copy()
    f3[c::N][:col.num_states] += 0.25
    j3 = np.array(jac_f(f3)).reshape(nn, K)
    c3 = con_f(f3).reshape(M, nn)
    changed = np.nonzero(np.any(j3 != jac.reshape(nn, K), axis=1))[0]
    assert set(changed) <= {c - 1, c} and len(changed) > 0
    changed = np.nonzero(np.any(c3 != con.reshape(M, nn), axis=0))[0]
    assert set(changed) == {c - 1, c}

    # (3) translation: shifting the trajectories by one node shifts the node
    # blocks by one (the kernel has no dependence on the absolute node index)
    rows_total = col.num_states + col.num_unknown_input_trajectories
    traj = free[:rows_total * N].reshape(rows_total, N)
    f4 = free.copy()
    f4[:rows_total * N] = np.roll(traj, -1, axis=1).ravel()
    j4 = np.array(jac_f(f4)).reshape(nn, K)
    c4 = con_f(f4).reshape(M, nn)
    assert np.array_equal(j4[:nn - 1], jac.reshape(nn, K)[1:])
    assert np.array_equal(c4[:, :nn - 1], con.reshape(M, nn)[:, 1:])

    # (4) the Jacobian is the derivative of the residuals: directional
    # finite difference at full size (coo_matvec with the bit-exact structure)
    rows, cols = col.jacobian_indices()
    rng = np.random.default_rng(5)
    d = rng.standard_normal(free.size)
    eps = 1e-6
    fd = (con_f(free + eps * d) - con_f(free - eps * d)) / (2 * eps)
    jv = np.bincount(rows, weights=jac * d[cols], minlength=len(con))
    assert np.max(np.abs(fd - jv)) <= 1e-5 * np.max(np.abs(jv))


def test_config2_kernel_variants_agree_bitwise(config2):
    """Group count, staging tile width, TMA vs warp-per-node stores, block
    size, input staging, the shared pre-pass and the scheduler's
    rematerialisation budget change the order of the operations, not the
    operations: with the association order of the tape (``reassociate=
    False``) and without FMA contraction (which nvcc applies differently to
    differently shaped code) all variants must be bit-identical -- including
    the unscheduled emission order."""
    w, col, free, con, jac = config2
    variants = gpu_variants.CONFIG2_VARIANTS
    ref_con = ref_jac = None
    for opts in variants:
        opts = dict(opts, **gpu_variants.BITWISE)
        other = _collocator(w, cuda_options=opts)
        c2 = other.generate_constraint_function()(free)
        j2 = np.array(other.generate_jacobian_function()(free))
        if ref_con is None:
            ref_con, ref_jac = c2, j2
        assert np.array_equal(c2, ref_con), opts
        assert np.array_equal(j2, ref_jac), opts
        other.close()
    # the default build (FMA contraction on, sums accumulated in arrival
    # order) stays within the parity bar of the uncontracted one
    P = col._evaluator.program.P
    assert_values_close(con, ref_con)
    assert_values_close(jac, ref_jac, row_len=P)


def test_config2_unfused_build_matches_reference_residuals_mostly_bitwise(
        config2):
    """With ``fmad=False, reassociate=False`` the residual code has the
    reference's operations in the reference's association order: most residuals are bit-identical to the oracle's (the rest differ
    by sin/cos rounding)."""
    w, col, free, con, jac = config2
    other = _collocator(w, cuda_options={'fmad': False,
                                         'reassociate': False})
    c2 = other.generate_constraint_function()(free)
    orc = OracleCollocator(*w.collocator_args(), **w.collocator_kwargs())
    ocon = orc.constraints(free)
    assert np.mean(c2 == ocon) > 0.8
    assert_values_close(c2, ocon)
    other.close()


def test_node_range_shards_reproduce_the_whole(config2):
    """Sharding the constraint nodes (multi-GPU decomposition, here on one
    device) reproduces the unsharded result bit for bit."""
    w, col, free, con, jac = config2
    nn, M = _eom_sizes(col)
    K = M * col._evaluator.program.P
    rows, cols = col.jacobian_indices()
    bounds = gpu_variants.CONFIG2_SHARD_BOUNDS
    assert bounds[-1] == nn
    # same group count => same generated module => same bits (the automatic
    # choice depends on the shard size, and FMA contraction on code shape)
    # (the row-stationary kernel always has one equation per group)
    opts = {} if col._evaluator.meta['persistent'] == 2 else \
        {'groups': col._evaluator.meta['num_groups']}
    for lo, hi in zip(bounds, bounds[1:]):
        part = _collocator(w, node_range=(lo, hi), cuda_options=opts)
        c = part.generate_constraint_function()(free)
        j = np.array(part.generate_jacobian_function()(free))
        r, cc = part.jacobian_indices()
        assert np.array_equal(c.reshape(M, hi - lo),
                              con.reshape(M, nn)[:, lo:hi])
        assert np.array_equal(j, jac[lo * K:hi * K])
        assert np.array_equal(r, rows[lo * K:hi * K])
        assert np.array_equal(cc, cols[lo * K:hi * K])
        part.close()


# ---------------------------------------------------------------------------
# stand-in for BASELINE config 4 at a larger size, against the oracle
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('opts', gpu_variants.CONFIG4_VARIANTS,
                         ids=gpu_variants.CONFIG4_IDS)
def test_config4_standin_against_oracle(opts):
    w = workloads.n_link_pendulum_torques(4, 2000)
    col = _collocator(w, cuda_options=opts)
    free = w.free(col.num_free)
    orc = OracleCollocator(*w.collocator_args(), **w.collocator_kwargs())
    con = col.generate_constraint_function()(free)
    jac = np.array(col.generate_jacobian_function()(free))
    ocon, ojac = orc.constraints(free), orc.jacobian(free)
    nn, M = _eom_sizes(col)
    P = col._evaluator.program.P
    assert_values_close(con[:M * nn], ocon[:M * nn])
    assert_values_close(jac[:nn * M * P], ojac[:nn * M * P], row_len=P)
    np.testing.assert_allclose(con[M * nn:], ocon[M * nn:], rtol=1e-13)
    np.testing.assert_allclose(jac[nn * M * P:], ojac[nn * M * P:],
                               rtol=1e-13)
    rows, cols = col.jacobian_indices()
    orows, ocols = orc.jacobian_indices()
    assert np.array_equal(rows, orows) and np.array_equal(cols, ocols)
    # a new value of the free time interval / parameters is picked up
    f2 = free.copy()
    f2[-1] *= 1.5
    f2[-2] += 0.1
    assert_values_close(col.generate_constraint_function()(f2)[:M * nn],
                        orc.constraints(f2)[:M * nn])
    assert_values_close(
        np.array(col.generate_jacobian_function()(f2))[:nn * M * P],
        orc.jacobian(f2)[:nn * M * P], row_len=P)
    col.close()


def test_config4_standin_full_size_against_oracle():
    """The stand-in of BASELINE config 4 at the size of its scaling runs
    (8 links, 20 000 backward-Euler nodes: 18 states, 8 unknown inputs, a
    known trajectory, unknown parameters, a free time interval, instance
    constraints; 16.9 M Jacobian entries) against the oracle, whole vectors;
    and its node shards (the strong-scaling decomposition over 8 GPUs, here
    the first, a middle and the last one on this device) bit for bit against
    the unsharded evaluation."""
    from opty_b200.sharding import node_shard
    w = workloads.n_link_pendulum_torques(8, 20000)
    col = _collocator(w)
    free = w.free(col.num_free)
    orc = OracleCollocator(*w.collocator_args(), **w.collocator_kwargs())
    con = col.generate_constraint_function()(free)
    jac = np.array(col.generate_jacobian_function()(free))
    ocon, ojac = orc.constraints(free), orc.jacobian(free)
    nn, M = _eom_sizes(col)
    P = col._evaluator.program.P
    K = M * P
    assert len(jac) >= 16_000_000
    assert_values_close(con[:M * nn], ocon[:M * nn])
    assert_values_close(jac[:nn * K], ojac[:nn * K], row_len=P)
    np.testing.assert_allclose(con[M * nn:], ocon[M * nn:], rtol=1e-13)
    np.testing.assert_allclose(jac[nn * K:], ojac[nn * K:], rtol=1e-13)
    rows, cols = col.jacobian_indices()
    orows, ocols = orc.jacobian_indices()
    assert np.array_equal(rows, orows) and np.array_equal(cols, ocols)
    opts = {'groups': col._evaluator.meta['num_groups'],
            'warps_per_block': col._evaluator.meta['warps_per_block'],
            'min_blocks_per_sm': col._evaluator.meta['min_blocks_per_sm']}
    for rank in (0, 3, 7):
        lo, hi = node_shard(20000, rank, 8)
        part = _collocator(w, node_range=(lo, hi), cuda_options=opts)
        c = part.generate_constraint_function()(free)
        j = np.array(part.generate_jacobian_function()(free))
        assert np.array_equal(c.reshape(M, hi - lo),
                              con[:M * nn].reshape(M, nn)[:, lo:hi])
        assert np.array_equal(j, jac[lo * K:hi * K])
        part.close()
    col.close()


@pytest.mark.parametrize('links,nodes', [(3, 200), (5, 700)])
def test_row_stationary_kernel_with_free_parameters_and_interval(links, nodes):
    """Even P: the automatic choice is the row-stationary kernel, here on
    backward-Euler problems with several unknown inputs, a known trajectory,
    unknown parameters and a free time interval (the node-invariant table is
    re-evaluated with every new free vector); three node tiles at 700 nodes,
    the last one ragged."""
    w = workloads.n_link_pendulum_torques(links, nodes)
    col = _collocator(w)
    assert col.prepare_module().meta['persistent'] == 2
    free = w.free(col.num_free)
    orc = OracleCollocator(*w.collocator_args(), **w.collocator_kwargs())
    nn, M = _eom_sizes(col)
    con_f = col.generate_constraint_function()
    jac_f = col.generate_jacobian_function()
    P = col._evaluator.program.P
    f2 = free.copy()
    f2[-1] *= 1.5
    f2[-2] += 0.1
    for point in (free, f2, free):
        con, jac = con_f(point), np.array(jac_f(point))
        ocon, ojac = orc.constraints(point), orc.jacobian(point)
        assert_values_close(con[:M * nn], ocon[:M * nn])
        assert_values_close(jac[:nn * M * P], ojac[:nn * M * P], row_len=P)
        np.testing.assert_allclose(con[M * nn:], ocon[M * nn:], rtol=1e-13)
        np.testing.assert_allclose(jac[nn * M * P:], ojac[nn * M * P:],
                                   rtol=1e-13)
    rows, cols = col.jacobian_indices()
    orows, ocols = orc.jacobian_indices()
    assert np.array_equal(rows, orows) and np.array_equal(cols, ocols)
    col.close()


# ---------------------------------------------------------------------------
# boundary behaviour
# ---------------------------------------------------------------------------
def test_problem_callback_surface():
    """Shapes, dtypes and ownership of the cyipopt callbacks
    (opty/direct_collocation.py:498-562)."""
    w = workloads.pendulum_swing_up(51)
    prob = Problem(lambda fr: float(np.sum(fr**2)), lambda fr: 2.0 * fr,
                   *w.collocator_args(), **w.collocator_kwargs(),
                   bounds={w.states[0]: (-10.0, 10.0)},
                   eom_bounds={0: (-20.0, 20.0)})
    free = w.free(prob.num_free)
    assert prob.num_free == 153 and prob.num_constraints == 104
    g = prob.constraints(free)
    assert g.shape == (104,) and g.dtype == np.float64
    rows, cols = prob.jacobianstructure()
    vals = prob.jacobian(free)
    assert vals.shape == rows.shape == cols.shape == (504,)
    assert rows.dtype == np.int64
    assert prob.objective(free) == float(np.sum(free**2))
    assert prob.gradient(free).shape == (153,)
    assert np.all(prob.lower_bound[:51] == -10.0)
    assert np.all(prob._low_con_bounds[:50] == -20.0)
    prob.intermediate(0, 0, 1.5)
    assert prob.obj_value == [1.5]
    # constraints() returns a fresh array; jacobian() a persistent buffer that
    # the next call overwrites (opty/direct_collocation.py:2814, 2887)
    v1 = np.array(vals)
    g2 = prob.constraints(free + 1.0)
    assert not np.shares_memory(g, g2) and not np.array_equal(g, g2)
    # ... and a constraints() call (which already starts moving the next
    # Jacobian) leaves the array returned by the last jacobian() untouched
    assert np.array_equal(v1, vals)
    v2 = prob.jacobian(free + 1.0)
    assert not np.array_equal(v1, np.array(v2))
    v3 = prob.jacobian(free + 2.0)
    assert np.shares_memory(vals, v3) or np.shares_memory(v2, v3)
    with pytest.raises(ValueError):
        prob.constraints(free[:-1])
    with pytest.raises(ValueError):
        prob.jacobian(np.zeros((153, 1)))
    prob.collocator.close()


def test_problem_helpers_fill_free_and_time_vector():
    """Host-side conveniences of the facade (opty/direct_collocation.py:
    1004-1028, 1097-1132) (the facade creates its device handle at construction)."""
    from opty_b200 import Problem
    import sympy as sm
    w = workloads.n_link_pendulum_torques(4, 50)
    prob = Problem(lambda f: 0.0, lambda f: np.zeros_like(f),
                   *w.collocator_args(), **w.collocator_kwargs(),
                   bounds={w.states[0]: (-1.0, 1.0)})
    col = prob.collocator
    N = col.num_collocation_nodes
    free = np.zeros(col.num_free)
    x1 = w.states[1]
    prob.fill_free(free, np.arange(N, dtype=float), x1)
    assert np.array_equal(prob.extract_values(free, x1), np.arange(N))
    assert np.count_nonzero(free) == N - 1
    h = col.time_interval_symbol
    prob.fill_free(free, 0.25, h)
    assert free[-1] == 0.25
    p0 = col.unknown_parameters[0]
    prob.fill_free(free, np.hstack((np.ones(N), 3.0)), w.states[0], p0)
    assert prob.extract_values(free, p0)[0] == 3.0
    with pytest.raises(ValueError):
        prob.fill_free(free, 1.0, sm.Symbol('nope'))
    np.testing.assert_allclose(prob.time_vector(solution=free, start_time=1.0),
                               1.0 + 0.25 * np.arange(N))
    with pytest.raises(ValueError):
        prob.time_vector()
    assert prob.bounds == {w.states[0]: (-1.0, 1.0)}
    assert prob.eom_bounds is None


def test_line_search_call_pattern_with_output_ring():
    """IPOPT's line search evaluates g at trial points without asking for
    jac_g there; with ``out_ring=2`` those evaluations do not wait for the
    speculative Jacobian copy of the previous point.  Whatever the call
    order, every result must belong to the point it was asked at."""
    w = workloads.n_link_pendulum(10, 40, seed=7)
    col = _collocator(w, cuda_options={'out_ring': 2})
    orc = OracleCollocator(*w.collocator_args(), **w.collocator_kwargs())
    base = w.free(col.num_free)
    pts = [base * (1.0 + 0.01 * i) for i in range(6)]
    con_f = col.generate_constraint_function()
    jac_f = col.generate_jacobian_function()
    P = col._evaluator.program.P
    # c = constraints, j = jacobian, digit = point
    pattern = ['c0', 'j0', 'c1', 'c2', 'c3', 'j3', 'j4', 'c4', 'c5', 'c0',
               'j0', 'j0', 'c1', 'j2', 'c2', 'c3', 'c4', 'c5', 'j5']
    for step in pattern:
        x = pts[int(step[1])]
        if step[0] == 'c':
            assert_values_close(con_f(x), orc.constraints(x))
        else:
            assert_values_close(np.array(jac_f(x)), orc.jacobian(x),
                                row_len=P)
    col.close()


def test_callable_known_trajectory_sees_free():
    """Known trajectories given as callables are re-evaluated with ``free``
    on every call (opty/direct_collocation.py:2916-2917)."""
    case = cases.msd_unknown_trajectory('backward euler')
    f_sym = list(case.traj_map.keys())[0]
    fs = case.traj_map[f_sym]
    traj_map = {f_sym: (lambda free: fs + free[0])}
    col = ConstraintCollocator(case.eom, case.states, case.N, case.h,
                               known_parameter_map=case.par_map,
                               known_trajectory_map=traj_map,
                               time_symbol=case.t)
    con = col.generate_constraint_function()(case.free)
    expected = case.expected_con.copy()
    expected[3:] -= case.free[0]
    np.testing.assert_allclose(con, expected, rtol=1e-12)
    col.close()


def test_two_atom_instance_constraints_against_reference_and_oracle():
    """Periodicity instance constraints with two function atoms each
    (examples-gallery/advanced/plot_human_gait.py:163-184), free node time
    interval, unknown parameters, a known trajectory: against the reference's
    fixture (instance part as a set of triplets: its order is Python's set
    iteration order in the generating process, opty/direct_collocation.py:
    2244, 2264) and, in this process, entry for entry against the oracle."""
    gold = load_golden('cfg4_periodic_pendulum4_N200')
    w = workloads.n_link_pendulum_periodic(4, 200)
    col = _collocator(w)
    free = w.free(col.num_free)
    assert np.array_equal(free, gold['free'])
    con = col.generate_constraint_function()(free)
    jac = np.array(col.generate_jacobian_function()(free))
    rows, cols = col.jacobian_indices()
    nn, M = _eom_sizes(col)
    P = col._evaluator.program.P
    nnz = nn * M * P
    assert len(con) == len(gold['con']) and len(jac) == len(gold['jac'])
    assert_values_close(con[:M * nn], gold['con'][:M * nn])
    np.testing.assert_allclose(con[M * nn:], gold['con'][M * nn:],
                               rtol=1e-13, atol=1e-15)
    assert_values_close(jac[:nnz], gold['jac'][:nnz], row_len=P)
    assert np.array_equal(rows[:nnz], gold['rows'][:nnz])
    assert np.array_equal(cols[:nnz], gold['cols'][:nnz])
    got = canonical_triplets(rows, cols, jac, nnz)
    want = canonical_triplets(gold['rows'], gold['cols'], gold['jac'], nnz)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    np.testing.assert_allclose(got[2], want[2], rtol=1e-13, atol=1e-15)
    orc = OracleCollocator(*w.collocator_args(), **w.collocator_kwargs())
    orows, ocols = orc.jacobian_indices()
    assert np.array_equal(rows, orows) and np.array_equal(cols, ocols)
    np.testing.assert_allclose(jac[nnz:], orc.jacobian(free)[nnz:],
                               rtol=1e-13, atol=1e-15)
    col.close()


def test_known_parameter_map_is_read_on_every_call():
    """The reference merges ``known_parameter_map`` / ``known_trajectory_map``
    into the arguments on every call (opty/direct_collocation.py:2973-2980):
    changing a value between two evaluations must change the results, for the
    collocation part and for the host-side instance constraints alike."""
    w = workloads.n_link_pendulum_torques(4, 200)
    col = _collocator(w)
    free = w.free(col.num_free)
    con_f = col.generate_constraint_function()
    jac_f = col.generate_jacobian_function()
    con0 = con_f(free)
    jac0 = np.array(jac_f(free))
    key = [p for p in w.known_parameter_map if p.name == 'l1'][0]
    w.known_parameter_map[key] *= 1.5
    traj_key = list(w.known_trajectory_map)[0]
    w.known_trajectory_map[traj_key] = w.known_trajectory_map[traj_key] + 0.5
    orc = OracleCollocator(*w.collocator_args(), **w.collocator_kwargs())
    con1 = con_f(free)
    jac1 = np.array(jac_f(free))
    assert not np.array_equal(con0, con1)
    assert not np.array_equal(jac0, jac1)
    nn, M = _eom_sizes(col)
    P = col._evaluator.program.P
    assert_values_close(con1[:M * nn], orc.constraints(free)[:M * nn])
    assert_values_close(jac1[:nn * M * P], orc.jacobian(free)[:nn * M * P],
                        row_len=P)
    col.close()


def test_jacobian_refetch_restores_columns_a_consumer_overwrote():
    w = workloads.n_link_pendulum(10, 40, seed=7)
    col = _collocator(w)
    free = w.free(col.num_free)
    jac_f = col.generate_jacobian_function()
    ref = np.array(jac_f(free))
    view = jac_f(free * 1.01)
    view *= 2.0                       # a consumer scales the buffer in place
    view = jac_f(free * 1.01)
    view[:] = 7.0
    again = np.array(jac_f(free, refetch=True))
    assert np.array_equal(again, ref)
    col.close()


@pytest.mark.parametrize('make', [
    lambda: workloads.n_link_pendulum_periodic(4, 200),
    lambda: workloads.n_link_pendulum(10, 40, seed=7),
    lambda: workloads.vyasarayani2011(101, seed=5),
], ids=['periodic4', 'pendulum10', 'vyasarayani'])
def test_sharded_handles_fill_one_host_vector(make):
    """The ``devices=`` path (one process, one handle per shard, every shard
    copying straight into its slice of one pinned host vector: contiguous
    Jacobian blocks, M strided residual segments) -- here with three shards
    on the one GPU of the test box -- reproduces the unsharded vectors bit
    for bit, instance-constraint tails included."""
    w = make()
    one = _collocator(w)
    many = ConstraintCollocator(*w.collocator_args(), **w.collocator_kwargs(),
                                devices=[0, 0, 0])
    free = w.free(one.num_free)
    con_f, jac_f = (many.generate_constraint_function(),
                    many.generate_jacobian_function())
    con1_f, jac1_f = (one.generate_constraint_function(),
                      one.generate_jacobian_function())
    for point in (free, free * 1.01, free * 1.01, free):
        assert np.array_equal(con_f(point), con1_f(point))
        assert np.array_equal(np.array(jac_f(point)), np.array(jac1_f(point)))
    # jacobian before constraints at a new point, and a refetch
    p2 = free * 0.99
    assert np.array_equal(np.array(jac_f(p2)), np.array(jac1_f(p2)))
    assert np.array_equal(con_f(p2), con1_f(p2))
    view = jac_f(p2)
    view[:] = -1.0
    assert np.array_equal(np.array(jac_f(p2, refetch=True)),
                          np.array(jac1_f(p2)))
    r1, c1 = one.jacobian_indices()
    r2, c2 = many.jacobian_indices()
    assert np.array_equal(r1, r2) and np.array_equal(c1, c2)
    one.close()
    many.close()


def test_c_abi_rejects_bad_configurations():
    w = workloads.vyasarayani2011(101, seed=5)
    col = _collocator(w)
    pm = col.prepare_module()
    cfg = runtime.ColloCfg()
    cfg.abi_version = 99
    with pytest.raises(ValueError):
        runtime.ColloHandle(cfg, pm.cubin)
    col.generate_constraint_function()
    good = col._evaluator.handle.cfg
    bad = runtime.ColloCfg.from_buffer_copy(good)
    bad.P = good.P + 1
    with pytest.raises(ValueError):
        runtime.ColloHandle(bad, pm.cubin)
    bad = runtime.ColloCfg.from_buffer_copy(good)
    with pytest.raises(RuntimeError):     # not a cubin
        runtime.ColloHandle(bad, b'not a cubin' * 100)
    # evaluating before the known values were supplied is a state error
    fresh = runtime.ColloHandle(runtime.ColloCfg.from_buffer_copy(good),
                                pm.cubin)
    with pytest.raises(RuntimeError):
        fresh.constraints(np.zeros(fresh.free_len))
    fresh.close()
    col.close()


# ---------------------------------------------------------------------------
# a larger model: 20-link pendulum (n = M = 42, P = 86, 43 k ops per node)
# ---------------------------------------------------------------------------
def _check_sampled_entries(gold, con, jac, M, P, nn, rtol=1e-10):
    """Residuals and Jacobian entries against the second oracle's sampled
    values (tests/golden/make_sampled_jacobian.py: ``sm.diff`` + 40-digit
    mpmath, correctly rounded).  Returns the max relative errors."""
    ent = gold['entries']
    got = np.array([jac.reshape(nn, M, P)[n_, r_, c_] for n_, r_, c_ in ent])
    exact = gold['jac_exact']
    scale = np.array([np.abs(jac.reshape(nn, M, P)[n_, r_]).max()
                      for n_, r_, c_ in ent])
    err = np.abs(got - exact)
    rel = err / np.abs(exact)
    ref_rel = np.abs(gold['jac_f64'] - exact) / np.abs(exact)
    print('\n[parity] {} sampled Jacobian entries vs exact values: max '
          'relative error {:.3e} (float64 lambdify of the reference\'s '
          'numpy backend: {:.3e}), max error / row scale {:.3e}'.format(
              len(ent), rel.max(), ref_rel.max(), (err / scale).max()))
    assert np.all(err <= rtol * np.abs(exact) + 1e-14 * scale)
    res = gold['res_entries']
    rgot = np.array([con.reshape(M, nn)[r_, n_] for n_, r_ in res])
    rex = gold['res_exact']
    rerr = np.abs(rgot - rex)
    print('[parity] {} sampled residuals vs exact values: max relative '
          'error {:.3e}'.format(len(res), (rerr / np.abs(rex)).max()))
    assert np.all(rerr <= rtol * np.abs(rex) + 1e-14 * np.abs(con).max())
    return rel.max()


def test_20_link_pendulum_residuals_and_jacobian():
    """The set-up pipeline scales (the reference needs ~3 min for this
    model's Jacobian).  Residuals are checked against the
    oracle's compiled C at every node; the Jacobian against 320 sampled
    entries evaluated exactly by the second oracle (rtol 1e-10) and against
    directional finite differences of the residuals."""
    w = workloads.n_link_pendulum(20, 2000, seed=9)
    col = _collocator(w)
    free = w.free(col.num_free)
    con_f = col.generate_constraint_function()
    jac_f = col.generate_jacobian_function()
    con = con_f(free)
    jac = np.array(jac_f(free))
    orc = OracleCollocator(*w.collocator_args(), **w.collocator_kwargs())
    assert_values_close(con, orc.constraints(free))
    rows, cols = col.jacobian_indices()
    assert len(rows) == len(jac) == 1999 * 42 * 86
    gold = load_golden('pendulum20_N2000_sampled_entries')
    assert np.array_equal(gold['free_head'], free[:8])
    _check_sampled_entries(gold, con, jac, 42, 86, 1999)
    rng = np.random.default_rng(2)
    d = rng.standard_normal(free.size)
    eps = 1e-6
    fd = (con_f(free + eps * d) - con_f(free - eps * d)) / (2 * eps)
    jv = np.bincount(rows, weights=jac * d[cols], minlength=len(con))
    assert np.max(np.abs(fd - jv)) <= 1e-5 * np.max(np.abs(jv))
    col.close()


# ---------------------------------------------------------------------------
# BASELINE config 5: 50-link chain (n = M = 102, P = 206, 544 k ops per node)
# ---------------------------------------------------------------------------
def test_config5_50_link_chain_against_exact_sampled_entries():
    """The module is prepared in the build container (tools/config5.py
    prepare: 3 min of SymPy derivation + scheduling + parallel nvcc) and
    travels as cubins plus a SymPy-free problem dump; here it is evaluated at
    2 000 nodes of the 50 000-node problem's free vector and checked against
    320 Jacobian entries and their residuals evaluated exactly by the second
    oracle (rtol 1e-10), against high-precision SymPy ``evalf`` of eight
    residual rows at node 0, and against directional finite differences."""
    import json
    import os
    from conftest import ROOT
    import sys
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import config5
    if not os.path.exists(config5.DUMP):
        pytest.fail('config 5 module dump {} is missing: run '
                    '`python tools/config5.py prepare` (done by '
                    '__graft_entry__.build)'.format(config5.DUMP))
    with open(config5.DUMP) as f:
        dump = json.load(f)
    n, q, M, P = dump['n'], dump['q'], dump['M'], dump['P']
    N_full, N = dump['num_nodes_full'], 2000
    free_full = config5.full_free_vector(dump)
    free = np.concatenate([free_full[j * N_full:j * N_full + N]
                           for j in range(n + q)])
    h = config5.make_handle(dump, N)
    nn = N - 1
    con = h.constraints(free).copy()
    jac = np.array(h.jacobian(free))
    gold = load_golden('cfg5_pendulum50_sampled_entries')
    assert np.array_equal(gold['free_head'], free[:8])
    _check_sampled_entries(gold, con, jac, M, P, nn)
    rows0 = load_golden('cfg5_pendulum50_node0_rows')
    got = con.reshape(M, nn)[rows0['rows'], 0]
    assert np.max(np.abs(got - rows0['values']) /
                  np.abs(rows0['values'])) < 1e-10
    rows, cols = runtime.jacobian_indices(0, N, 0, nn, n, q, 0, 0, M, 1)
    d = np.random.default_rng(2).standard_normal(free.size)
    eps = 1e-6
    fd = (h.constraints(free + eps * d).copy() -
          h.constraints(free - eps * d).copy()) / (2 * eps)
    jv = np.bincount(rows, weights=jac * d[cols], minlength=len(con))
    assert np.max(np.abs(fd - jv)) <= 1e-5 * np.max(np.abs(jv))
    h.close()
