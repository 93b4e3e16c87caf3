import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line(
        'markers', 'gpu: needs a CUDA device (run with -m gpu on a B200)')


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


def assert_values_close(actual, expected, row_len=None, rtol=1e-10,
                        scale_tol=1e-14):
    """The parity criterion for float64 results (north star: within 1e-10
    relative of the reference):

        |actual - expected| <= rtol * |expected| + scale_tol * scale

    ``scale`` is the largest magnitude in the entry's row -- the ``row_len``
    consecutive partials of one equation at one node -- or in the whole array
    if ``row_len`` is None.  The second term only matters for entries that are
    formed by cancellation of much larger terms (|entry| << row scale), where
    *any* re-association of the float64 sum moves the result by a few ulps of
    the large terms; it allows 1e-14 of the row scale, i.e. ~45 ulps.
    Structural zeros must be exactly zero on both sides.
    """
    actual = np.asarray(actual, dtype=float)
    expected = np.asarray(expected, dtype=float)
    assert actual.shape == expected.shape
    assert np.all(np.isfinite(actual)) and np.all(np.isfinite(expected))
    if row_len is None:
        scale = np.max(np.abs(expected)) if expected.size else 0.0
    else:
        scale = np.repeat(np.max(np.abs(expected.reshape(-1, row_len)),
                                 axis=1), row_len)
    err = np.abs(actual - expected)
    bound = rtol * np.abs(expected) + scale_tol * scale
    bad = err > bound
    if np.any(bad):
        i = int(np.argmax(err - bound))
        raise AssertionError(
            '{} of {} entries outside tolerance; worst at {}: actual {!r}, '
            'expected {!r}'.format(int(bad.sum()), bad.size, i, actual[i],
                                   expected[i]))
    zero = expected == 0.0
    assert np.all(actual[zero] == 0.0), 'structural zeros differ'


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


def canonical_triplets(rows, cols, vals, start):
    """(row, col, value) triplets of the entries from ``start`` on, sorted by
    (row, col): the instance-constraint part of a Jacobian, whose order inside
    one constraint follows Python's set iteration in the reference
    (opty/direct_collocation.py:2244, 2264) and may differ between
    processes."""
    r = np.asarray(rows[start:])
    c = np.asarray(cols[start:])
    v = np.asarray(vals[start:], dtype=float)
    order = np.lexsort((c, r))
    return r[order], c[order], v[order]
