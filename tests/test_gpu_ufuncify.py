"""The generic operator seam ``ufuncify_matrix(args, expr, const=...)``
(opty/utils.py:639-670) on the GPU; mirrors the reference's own test
(opty/tests/test_utils.py:244-336) and cross-checks against the oracle's
restatement of it."""

import numpy as np
import pytest
import sympy as sm

from opty_b200.utils import ufuncify_matrix
from oracle.opty_oracle import compile_matrix_function

pytestmark = pytest.mark.gpu


def _setup():
    a, b, c, d, I, i = sm.symbols('a, b, if, d_{badsym}, I, i')
    mat = sm.Matrix([
        [a**2 * sm.cos(sm.pi * b)**c, sm.tan(b) / sm.sin(a + b) + c**4],
        [a**2 + b**2 - sm.sqrt(c), ((a + b + c) * (a + b)) / a * sm.sin(b)]])
    return (a, b, c, d, I, i), mat


def _numpy_eval(n, a_vals, b_vals, c_vals):
    result = np.empty((n, 2, 2))
    result[:, 0, 0] = a_vals**2 * np.cos(np.pi * b_vals)**c_vals
    result[:, 0, 1] = np.tan(b_vals) / np.sin(a_vals + b_vals) + c_vals**4
    result[:, 1, 0] = a_vals**2 + b_vals**2 - np.sqrt(c_vals)
    result[:, 1, 1] = (((a_vals + b_vals + c_vals) * (a_vals + b_vals)) /
                       a_vals * np.sin(b_vals))
    return result


def test_ufuncify_matrix_like_the_reference_test():
    (a, b, c, d, I, i), mat = _setup()
    n = 10000
    rng = np.random.default_rng(0)
    a_vals = rng.random(n)
    b_vals = rng.random(n)
    c_vals = rng.random(n) + 10.0
    c_val = rng.random() + 10.0

    f = ufuncify_matrix((a, b, c), mat)
    result = np.empty((n, 4))
    out = f(result, a_vals, b_vals, c_vals)
    assert out.shape == (n, 2, 2) and np.shares_memory(out, result)
    np.testing.assert_allclose(out, _numpy_eval(n, a_vals, b_vals, c_vals),
                               rtol=1e-12)

    for parallel in (False, True):
        f = ufuncify_matrix((a, b, c), mat, const=(c,), parallel=parallel)
        result = np.empty((n, 4))
        np.testing.assert_allclose(f(result, a_vals, b_vals, c_val),
                                   _numpy_eval(n, a_vals, b_vals, c_val),
                                   rtol=1e-12)
        # a new value of the const argument is picked up
        np.testing.assert_allclose(f(result, a_vals, b_vals, c_val + 1.0),
                                   _numpy_eval(n, a_vals, b_vals, c_val + 1.0),
                                   rtol=1e-12)

    # awkward symbol names never reach a compiler here
    for sym in (I, i, d):
        f = ufuncify_matrix((a, b, sym), mat.xreplace({c: sym}))
        result = np.empty((n, 4))
        np.testing.assert_allclose(f(result, a_vals, b_vals, c_vals),
                                   _numpy_eval(n, a_vals, b_vals, c_vals),
                                   rtol=1e-12)


def test_ufuncify_matrix_against_oracle_and_argument_checks():
    (a, b, c, d, I, i), mat = _setup()
    rng = np.random.default_rng(1)
    for n in (1, 33, 1000):
        a_vals, b_vals = rng.random(n), rng.random(n)
        c_vals = rng.random(n) + 10.0
        f = ufuncify_matrix((a, b, c), mat)
        g = compile_matrix_function((a, b, c), mat)
        got = f(np.empty((n, 4)), a_vals, b_vals, c_vals)
        want = g(np.empty((n, 4)), a_vals, b_vals, c_vals)
        np.testing.assert_allclose(got, want, rtol=1e-12)
    # a cse() pair is accepted like a matrix (opty/utils.py:745-749)
    pair = sm.cse(mat, sm.numbered_symbols('z_'), order='none')
    f = ufuncify_matrix((a, b, c), pair)
    np.testing.assert_allclose(f(np.empty((n, 4)), a_vals, b_vals, c_vals),
                               want, rtol=1e-12)
    # odd number of outputs: the non-TMA store path
    col = sm.Matrix([[a * b, sm.sin(a), c]])
    f = ufuncify_matrix((a, b, c), col)
    got = f(np.empty((n, 3)), a_vals, b_vals, c_vals)
    np.testing.assert_allclose(got[:, 0, :], np.stack(
        [a_vals * b_vals, np.sin(a_vals), c_vals], axis=1), rtol=1e-13)
    with pytest.raises(ValueError):      # non-contiguous argument
        f(np.empty((n, 3)), a_vals[::2], b_vals, c_vals)
    with pytest.raises(ValueError):      # wrong dtype
        f(np.empty((n, 3)), a_vals.astype(np.float32), b_vals, c_vals)
    with pytest.raises(ValueError):      # wrong matrix shape
        f(np.empty((n, 4)), a_vals, b_vals, c_vals)
    with pytest.raises(ValueError):      # unknown symbol in the expressions
        ufuncify_matrix((a, b), col)
