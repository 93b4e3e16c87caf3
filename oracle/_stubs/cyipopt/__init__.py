"""Minimal stand-in for the third-party ``cyipopt`` package (absent in this image).

TEST INFRASTRUCTURE ONLY.  It lets ``/root/reference/opty`` be imported so that
its constraint / Jacobian code path can be executed as the parity oracle.  The
reference never touches IPOPT on that path (opty/direct_collocation.py:242-247
only forwards the problem sizes and bounds to ``cyipopt.Problem.__init__``).
"""


class Problem(object):

    def __init__(self, n=None, m=None, lb=None, ub=None, cl=None, cu=None,
                 **kwargs):
        self._ipopt_n, self._ipopt_m = n, m
        self._ipopt_lb, self._ipopt_ub = lb, ub
        self._ipopt_cl, self._ipopt_cu = cl, cu
        self._ipopt_options = {}

    def add_option(self, key, value):
        self._ipopt_options[key] = value

    def solve(self, *args, **kwargs):
        raise RuntimeError('IPOPT is not available: cyipopt stub in use.')
