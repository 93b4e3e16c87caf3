"""Second, compiler-free CPU oracle: the reference's ``backend='numpy'`` path.

TEST INFRASTRUCTURE ONLY (same rule as ``oracle/opty_oracle.py``: imported by
``tests/`` and the fixture generators under ``tests/golden/`` only, never by
``opty_b200/``).

Restates, from the reference's sources,

    lambdify_matrix            opty/utils.py:598-636
    symbolic Jacobian          opty/direct_collocation.py:2757-2758
                               (``discrete_eom_matrix.jacobian(wrt_matrix.T)``:
                               plain SymPy differentiation of every entry, no
                               CSE-aware forward mode, no C compiler)
    constraint function        opty/direct_collocation.py:2373-2380, 2382-2446
    Jacobian function          opty/direct_collocation.py:2798-2805, 2816-2887

It differs from the first oracle (generated C through gcc) in differentiation
algorithm, operation order and evaluator, which is what makes it a *second*
opinion.  Two uses:

* small problems: the full ``constraints(free)`` / ``jacobian(free)`` with the
  reference's per-node Python loop.  Parity is PINNED by
  ``tests/golden/*_numpy_backend.npz``, produced by running the reference
  itself with ``backend='numpy'`` (``tests/golden/make_golden_numpy.py``).
* large problems (20- and 50-link chains, where neither the C oracle nor a
  full symbolic Jacobian can be built in reasonable time):
  ``residual_entries`` / ``jacobian_entries`` evaluate sampled (node, row[,
  column]) entries -- ``sm.diff`` of one discrete EOM row with respect to one
  ``wrt`` symbol, then ``lambdify`` with NumPy (float64, what the reference's
  numpy backend computes) or mpmath (arbitrary precision: the exact value to
  1 ulp, against which 1e-10 is meaningful even for ill-conditioned entries).
"""

from functools import partial

import numpy as np
import sympy as sm

from .opty_oracle import OracleCollocator


def lambdify_matrix(args, expr):
    """Restatement of ``opty.utils.lambdify_matrix`` (opty/utils.py:598-636):
    ``f(store, *num_args) -> store.reshape(n, rows, cols)``; array arguments
    have shape ``(n,)``, constants are floats; one ``lambdify`` call per
    node."""
    single = sm.lambdify(args, expr, modules='numpy',
                         cse=partial(sm.cse, order='none', list=False),
                         docstring_limit=0)
    rows, cols = expr.shape

    def loop(result, *num_args):
        n = result.shape[0]
        for i in range(n):
            vals = [a if isinstance(a, float) else a[i] for a in num_args]
            result[i] = single(*vals).flatten().squeeze()
        return result.reshape(n, rows, cols)

    return loop


class LambdifyOracle(OracleCollocator):
    """``OracleCollocator`` bookkeeping (symbol order, discretisation,
    argument and ``wrt`` order: all cited there) with the NumPy evaluators of
    the reference's ``backend='numpy'``."""

    def _constraint_loop(self):
        # opty/direct_collocation.py:2378-2380
        if self._con_loop is None:
            self._con_loop = lambdify_matrix(
                self._args(), sm.ImmutableDenseMatrix(self.discrete_eom))
        return self._con_loop

    def _jacobian_loop(self):
        # opty/direct_collocation.py:2757-2758, 2803-2805
        if self._jac_loop is None:
            eom = sm.ImmutableDenseMatrix(self.discrete_eom)
            wrt = sm.ImmutableDenseMatrix([list(self._wrt())])
            self._jac_loop = lambdify_matrix(self._args(),
                                             eom.jacobian(wrt.T))
            self._jac_buffer = np.empty((self.N - 1, self.M * self.P))
        return self._jac_loop

    # -- sampled entries (large models) ---------------------------------
    def _node_values(self, free, node):
        """Scalar argument values of constraint node ``node`` in ``_args()``
        order (the slicing of opty/direct_collocation.py:2411-2437 at one
        node)."""
        vals = self._numeric_args(np.asarray(free, dtype=float))
        return [v if isinstance(v, float) else float(v[node]) for v in vals]

    def _evaluate(self, expr, values, dps):
        args = self._args()
        if dps is None:
            f = sm.lambdify(args, expr, modules='numpy', cse=True,
                            docstring_limit=0)
            return [float(f(*v)) for v in values]
        import mpmath
        f = sm.lambdify(args, expr, modules='mpmath', cse=True,
                        docstring_limit=0)
        out = []
        with mpmath.workdps(dps):
            for v in values:
                out.append(float(f(*[mpmath.mpf(x) for x in v])))
        return out

    def residual_entries(self, free, entries, dps=None):
        """Values of discrete EOM ``row`` at constraint node ``node`` for
        ``entries = [(node, row), ...]``; ``dps``: mpmath decimal digits or
        None for NumPy float64."""
        out = np.empty(len(entries))
        by_row = {}
        for idx, (node, row) in enumerate(entries):
            by_row.setdefault(row, []).append((idx, node))
        for row, items in by_row.items():
            vals = self._evaluate(self.discrete_eom[row],
                                  [self._node_values(free, nd)
                                   for _, nd in items], dps)
            for (idx, _), v in zip(items, vals):
                out[idx] = v
        return out

    def jacobian_entries(self, free, entries, dps=None):
        """``d discrete_eom[row] / d wrt[col]`` at constraint node ``node``
        for ``entries = [(node, row, col), ...]`` (the entry the reference
        stores at ``jac[node*M*P + row*P + col]``,
        opty/direct_collocation.py:2681-2684)."""
        wrt = self._wrt()
        out = np.empty(len(entries))
        by_rc = {}
        for idx, (node, row, col) in enumerate(entries):
            by_rc.setdefault((row, col), []).append((idx, node))
        for (row, col), items in by_rc.items():
            partial_ = sm.diff(self.discrete_eom[row], wrt[col])
            vals = self._evaluate(partial_,
                                  [self._node_values(free, nd)
                                   for _, nd in items], dps)
            for (idx, _), v in zip(items, vals):
                out[idx] = v
        return out

    def structural_nonzero(self, row, col):
        return self.discrete_eom[row].has(self._wrt()[col])
