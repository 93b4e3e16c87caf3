"""GPU tests of the objective / gradient path:
``opty_b200.utils.create_objective_function`` -- integrand lowered through the
tape -> CUDA-C emitter, quadrature weights and reductions in
``opty_colloc_quadrature`` -- against values produced by the reference's
``opty.utils.create_objective_function`` (tests/golden/objective_cases.npz,
made by tests/golden/make_golden_objective.py) and against the hand-computed
answers of the reference's own tests (opty/tests/test_utils.py:67-220)."""

import numpy as np
import pytest
import sympy as sm

import cases
from conftest import load_golden
from opty_b200.utils import create_objective_function

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', list(cases.objective_cases()))
def test_objective_and_gradient_match_reference(name):
    gold = load_golden('objective_cases')
    c = cases.objective_cases()[name]()
    obj, grad = create_objective_function(
        c['objective'], c['states'], c['inputs'], c['params'], c['N'],
        c['h'], integration_method=c['method'], time_symbol=c['t'])
    free = c['free']
    value = obj(free)
    g = grad(free)
    assert isinstance(value, float)
    assert g.shape == (len(free),)
    np.testing.assert_allclose(value, gold[name + '_value'], rtol=1e-10)
    if name + '_grad_index' in gold.files:
        g = g[gold[name + '_grad_index']]
    np.testing.assert_allclose(g, gold[name + '_grad'], rtol=1e-10,
                               atol=1e-15)
    # a second point, then the first again: nothing stale, bit-reproducible
    other = free * 1.25
    assert obj(other) != value
    assert obj(free) == value
    g2 = grad(free)
    if name + '_grad_index' in gold.files:
        g2 = g2[gold[name + '_grad_index']]
    assert np.array_equal(g2, g)


def test_objective_known_answers_of_the_reference_tests():
    """opty/tests/test_utils.py:97-106 and :182-205."""
    t = sm.symbols('t')
    x, v, f1, f2 = [f(t) for f in sm.symbols('x, v, f1, f2', cls=sm.Function)]
    m, c, k = sm.symbols('m, c, k')
    N = 20
    rng = np.random.default_rng(3)
    xv, vv, f1v, f2v = (rng.random(N) for _ in range(4))
    mv, cv, kv = rng.random(3)
    free = np.hstack((xv, vv, f1v, f2v, cv, kv, mv))
    obj, grad = create_objective_function(
        sm.Integral(x ** 2, t), [x, v], [f2, f1], [m, c, k], N, 0.5,
        time_symbol=t)
    np.testing.assert_allclose(obj(free), 0.5 * (xv[1:] ** 2).sum())
    np.testing.assert_allclose(grad(free), np.hstack((
        0, 0.5 * 2 * xv[1:], np.zeros(N * 3 + 3))))
    expr = (sm.Integral(x ** 2 + m ** 2, t) +
            sm.Integral(c ** 2 * f2 ** 2, t) + sm.sin(k) ** 2)
    obj, grad = create_objective_function(
        expr, [x, v], [f2, f1], [m, c, k], N, 0.3,
        integration_method='midpoint', time_symbol=t)
    x_mid = (xv[1:] + xv[:-1]) / 2
    f2_mid = (f2v[1:] + f2v[:-1]) / 2
    np.testing.assert_allclose(
        obj(free), 0.3 * ((x_mid ** 2).sum() + (N - 1) * mv ** 2 +
                          (cv ** 2 * f2_mid ** 2).sum()) + np.sin(kv) ** 2)
    np.testing.assert_allclose(grad(free), np.hstack((
        0.3 * xv[0], 0.3 * 2 * xv[1:-1], 0.3 * xv[-1], np.zeros(N * 2),
        0.3 * cv ** 2 * f2v[0], 0.3 * 2 * cv ** 2 * f2v[1:-1],
        0.3 * cv ** 2 * f2v[-1], 0.3 * 2 * cv * (f2_mid ** 2).sum(),
        2 * np.sin(kv) * np.cos(kv), 0.3 * (N - 1) * 2 * mv)))
