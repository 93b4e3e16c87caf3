"""Stand-in for the third-party ``cyipopt`` package (absent from this image and
not installable offline), used ONLY so that the unmodified reference installed
under ``baseline/_ref`` can be imported by ``bench.py --impl reference`` and by
the fixture generators under ``tests/golden/``.

The reference imports ``cyipopt`` at module level (opty/direct_collocation.py:
10) but its constraint / Jacobian path never calls IPOPT: ``Problem.__init__``
only forwards the problem sizes and bounds to ``cyipopt.Problem.__init__``
(opty/direct_collocation.py:242-247)."""


class Problem(object):

    def __init__(self, n=None, m=None, lb=None, ub=None, cl=None, cu=None,
                 **kwargs):
        self._n, self._m = n, m
        self._lb, self._ub, self._cl, self._cu = lb, ub, cl, cu
        self._options = {}

    def add_option(self, key, value):
        self._options[key] = value

    def solve(self, *args, **kwargs):
        raise RuntimeError('IPOPT is not available: cyipopt stand-in in use.')
