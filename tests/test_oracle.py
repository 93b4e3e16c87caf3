"""Pins the CPU oracle (oracle/opty_oracle.py) to the reference: golden
vectors produced by running csu-hmc/opty itself (tests/golden/make_golden.py)
and the hand-computed known-answer cases of the reference's own tests."""

import hashlib

import numpy as np
import pytest

import cases
import workloads
from conftest import (assert_values_close, canonical_triplets,
                      load_golden)
from oracle.opty_oracle import OracleCollocator, forward_jacobian


def _digest(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


FULL_FIXTURES = [
    ('cfg1_pendulum_swing_up_N51', lambda: workloads.pendulum_swing_up(51)),
    ('cfg3_vyasarayani2011_N5000', lambda: workloads.vyasarayani2011(5000)),
    ('cfg3_vyasarayani2011_N101_odd',
     lambda: workloads.vyasarayani2011(101, seed=5)),
    ('cfg4_standin_pendulum4_torques_N200',
     lambda: workloads.n_link_pendulum_torques(4, 200)),
    ('cfg2_small_pendulum10_N40',
     lambda: workloads.n_link_pendulum(10, 40, seed=7)),
]


@pytest.mark.parametrize('name,make', FULL_FIXTURES,
                         ids=[f[0] for f in FULL_FIXTURES])
def test_oracle_matches_reference_golden(name, make):
    gold = load_golden(name)
    w = make()
    orc = OracleCollocator(*w.collocator_args(), **w.collocator_kwargs())
    free = w.free(orc.num_free)
    assert np.array_equal(free, gold['free'])
    rows, cols = orc.jacobian_indices()
    assert rows.dtype == np.int64 and cols.dtype == np.int64
    assert np.array_equal(rows, gold['rows'])
    assert np.array_equal(cols, gold['cols'])
    # same generated C, same compiler, same libm => identical bits
    assert np.array_equal(orc.constraints(free), gold['con'])
    assert np.array_equal(orc.jacobian(free), gold['jac'])


def test_oracle_matches_reference_on_config2():
    """BASELINE config 2 (10-link pendulum, 10 000 midpoint nodes): the
    fixture holds a sample of nodes and SHA-256 digests of the complete
    reference arrays."""
    gold = load_golden('cfg2_pendulum10_N10000')
    w = workloads.n_link_pendulum(10, 10000)
    orc = OracleCollocator(*w.collocator_args(), **w.collocator_kwargs())
    free = w.free(orc.num_free)
    assert _digest(free) == str(gold['free_sha256'])
    con = orc.constraints(free)
    jac = orc.jacobian(free)
    rows, cols = orc.jacobian_indices()
    nn, M = orc.N - 1, orc.M
    K = M * orc.P
    nodes = gold['nodes']
    assert len(jac) == int(gold['nnz']) == 10118988
    assert np.array_equal(con.reshape(M, nn)[:, nodes], gold['con'])
    assert np.array_equal(jac.reshape(nn, K)[nodes], gold['jac'])
    assert np.array_equal(rows.reshape(nn, K)[nodes], gold['rows'])
    assert np.array_equal(cols.reshape(nn, K)[nodes], gold['cols'])
    assert _digest(rows.astype(np.int64)) == str(gold['rows_sha256'])
    assert _digest(cols.astype(np.int64)) == str(gold['cols_sha256'])
    assert _digest(con) == str(gold['con_sha256'])
    assert _digest(jac) == str(gold['jac_sha256'])


@pytest.mark.parametrize('case', cases.all_cases(), ids=lambda c: c.name)
def test_oracle_known_answers(case):
    orc = OracleCollocator(*case.collocator_args(),
                           **{k: v for k, v in
                              case.collocator_kwargs().items()})
    con = orc.constraints(case.free)
    jac = orc.jacobian(case.free)
    np.testing.assert_allclose(con, case.expected_con, rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(jac, case.expected_jac, rtol=1e-12, atol=1e-9)
    rows, cols = orc.jacobian_indices()
    assert len(rows) == len(cols) == len(jac)
    if case.expected_rows is not None:
        assert np.array_equal(rows, case.expected_rows)
        assert np.array_equal(cols, case.expected_cols)


def test_oracle_parallel_is_bit_identical():
    """OpenMP and serial loops give identical bits."""
    w = workloads.n_link_pendulum_torques(4, 200)
    a = OracleCollocator(*w.collocator_args(), **w.collocator_kwargs())
    b = OracleCollocator(*w.collocator_args(), **w.collocator_kwargs(),
                         parallel=True)
    free = w.free(a.num_free)
    assert np.array_equal(a.constraints(free), b.constraints(free))
    assert np.array_equal(a.jacobian(free), b.jacobian(free))


def test_forward_jacobian_matches_sympy_diff():
    import sympy as sm
    a, b, c = sm.symbols('a b c', real=True)
    expr = sm.ImmutableDenseMatrix([a * sm.sin(b) + c / a,
                                    sm.exp(a * b) * c**2 - b])
    wrt = sm.ImmutableDenseMatrix([a, b, c])
    repl, (jac,) = forward_jacobian(expr, wrt)
    full = jac
    for sym, sub in reversed(repl):
        full = full.xreplace({sym: sub})
    assert sm.simplify(full - expr.jacobian(wrt)) == sm.zeros(2, 3)


# ---------------------------------------------------------------------------
# second oracle: the reference's backend='numpy' path
# ---------------------------------------------------------------------------
NUMPY_FIXTURES = [
    ('cfg1_pendulum_swing_up_N51_numpy_backend',
     lambda: workloads.pendulum_swing_up(51)),
    ('cfg3_vyasarayani2011_N101_odd_numpy_backend',
     lambda: workloads.vyasarayani2011(101, seed=5)),
    ('cfg4_standin_pendulum4_torques_N30_numpy_backend',
     lambda: workloads.n_link_pendulum_torques(4, 30)),
    ('cfg2_small_pendulum10_N12_numpy_backend',
     lambda: workloads.n_link_pendulum(10, 12, seed=7)),
]


@pytest.mark.parametrize('name,make', NUMPY_FIXTURES,
                         ids=[f[0] for f in NUMPY_FIXTURES])
def test_lambdify_oracle_matches_reference_numpy_backend(name, make):
    """The second oracle reproduces the outputs of the reference run with
    backend='numpy' (tests/golden/make_golden_numpy.py) bit for bit, and the
    two oracles -- different differentiation, different evaluator -- agree
    to 1e-10 between themselves."""
    from oracle.lambdify_oracle import LambdifyOracle
    gold = load_golden(name)
    w = make()
    lam = LambdifyOracle(*w.collocator_args(), **w.collocator_kwargs())
    free = w.free(lam.num_free)
    assert np.array_equal(free, gold['free'])
    con = lam.constraints(free)
    jac = lam.jacobian(free)
    assert np.array_equal(con, gold['con'])
    assert np.array_equal(jac, gold['jac'])
    orc = OracleCollocator(*w.collocator_args(), **w.collocator_kwargs())
    nnz_colloc = (lam.N - 1) * lam.M * lam.P
    assert_values_close(orc.constraints(free), con)
    ojac = orc.jacobian(free)
    assert_values_close(ojac[:nnz_colloc], jac[:nnz_colloc], row_len=lam.P)
    assert_values_close(ojac[nnz_colloc:], jac[nnz_colloc:])


def test_lambdify_oracle_sampled_entries_agree_with_full_evaluation():
    """``jacobian_entries`` / ``residual_entries`` (the sampled form used for
    the 20- and 50-link chains) give the same numbers as the full evaluation,
    in float64 and -- to rounding -- in 40-digit arithmetic."""
    from oracle.lambdify_oracle import LambdifyOracle
    w = workloads.n_link_pendulum_torques(4, 30)
    lam = LambdifyOracle(*w.collocator_args(), **w.collocator_kwargs())
    free = w.free(lam.num_free)
    M, P, nn = lam.M, lam.P, lam.N - 1
    jac = lam.jacobian(free)[:nn * M * P].reshape(nn, M, P)
    con = lam.constraints(free)[:M * nn].reshape(M, nn)
    rng = np.random.default_rng(0)
    entries = [(int(rng.integers(nn)), int(rng.integers(M)),
                int(rng.integers(P))) for _ in range(24)]
    got = lam.jacobian_entries(free, entries)
    exact = lam.jacobian_entries(free, entries, dps=40)
    ref = np.array([jac[e] for e in entries])
    scale = np.array([np.abs(jac[e[0], e[1]]).max() for e in entries])
    assert np.all(np.abs(got - ref) <= 1e-12 * np.abs(ref) + 1e-14 * scale)
    assert np.all(np.abs(exact - ref) <= 1e-10 * np.abs(ref) + 1e-14 * scale)
    for (node, row, col), v in zip(entries, ref):
        assert (v != 0.0) == lam.structural_nonzero(row, col) or v == 0.0
    res = lam.residual_entries(free, [(e[0], e[1]) for e in entries], dps=40)
    rref = np.array([con[e[1], e[0]] for e in entries])
    assert np.all(np.abs(res - rref) <= 1e-10 * np.abs(rref) +
                  1e-14 * np.abs(con).max())


def test_oracle_two_atom_instance_constraints_match_reference():
    """Periodicity constraints ``x(0) - y(T)`` (examples-gallery/advanced/
    plot_human_gait.py:163-184): two function atoms per constraint.  EOM part
    bit for bit in order; the instance part as (row, col, value) triplets."""
    gold = load_golden('cfg4_periodic_pendulum4_N200')
    w = workloads.n_link_pendulum_periodic(4, 200)
    orc = OracleCollocator(*w.collocator_args(), **w.collocator_kwargs())
    free = w.free(orc.num_free)
    assert np.array_equal(free, gold['free'])
    assert np.array_equal(orc.constraints(free), gold['con'])
    jac = orc.jacobian(free)
    rows, cols = orc.jacobian_indices()
    nnz = (orc.N - 1) * orc.M * orc.P
    assert len(jac) - nnz == 18 and orc.o == 10
    assert np.array_equal(jac[:nnz], gold['jac'][:nnz])
    assert np.array_equal(rows[:nnz], gold['rows'][:nnz])
    assert np.array_equal(cols[:nnz], gold['cols'][:nnz])
    got = canonical_triplets(rows, cols, jac, nnz)
    want = canonical_triplets(gold['rows'], gold['cols'], gold['jac'], nnz)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
