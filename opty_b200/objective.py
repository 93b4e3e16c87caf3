"""Objective and gradient of a running-cost integral on the device.

Counterpart of ``opty.utils.create_objective_function`` (opty/utils.py:
329-470; pinned by opty/tests/test_utils.py:67-220): the objective is a
symbolic expression that may contain indefinite integrals of time,
``Integral(f(x(t), u(t), p), t)``, discretised with the quadrature that
belongs to the integration method

    backward Euler   J = h * sum_{i>=1} f(x_i, u_i, p)            (UT:418-423)
    midpoint         J = h * sum_{i<N-1} f(x_mid_i, u_mid_i, p)   (UT:424-434)

and the gradient with respect to the free vector ``[x_1(t_0..), ..., u_q(..),
p_1..p_r]`` -- for the midpoint rule with the reference's weights: trajectory
partials evaluated at the NODE values with weights ``(1/2, 1, ..., 1, 1/2)``,
parameter partials at the midpoints (opty/utils.py:456-466).

The reference evaluates these with NumPy ``lambdify`` on the host.  Here the
integrand and its partials are lowered through the same tape -> CUDA-C
emitter as the constraints (one straight-line body per point,
:class:`opty_b200.program.CollocationProgram.from_matrix`), the quadrature
weights are applied and the sums reduced on the device
(``opty_colloc_quadrature``, csrc/runtime.cu), and the free vector is only
uploaded when it differs from the resident copy, so that IPOPT's ``f`` /
``grad_f`` pair at one point costs one upload.

Supported objectives: sums of integrals with numeric coefficients plus terms
that depend on the unknown parameters only -- every form the reference's
tests and examples use.  (Products of integrals do not evaluate correctly in
the reference either: its gradient path replaces every integral by the
un-summed array, opty/utils.py:369-374.)
"""

import numpy as np
import sympy as sm
import sympy.physics.mechanics as me

from . import runtime
from .program import CollocationProgram
from .utils import sort_sympy

__all__ = ['create_objective_function']

_ELEMENTWISE = 2
_RULES = {'backward euler': 0, 'midpoint': 1}


def _split_objective(objective, time_symbol):
    """``objective = sum_k c_k Integral(f_k, t) + rest`` ->
    ``(sum_k c_k f_k, rest)``."""
    integrand = sm.S.Zero
    rest = sm.S.Zero
    for term in sm.Add.make_args(sm.expand(objective, deep=False)):
        integrals = term.atoms(sm.Integral)
        if not integrals:
            rest += term
            continue
        coeff, factors = term.as_coeff_mul()
        if len(factors) != 1 or not isinstance(factors[0], sm.Integral):
            raise NotImplementedError(
                'Only sums of integrals with numeric coefficients (plus terms '
                'that depend on the unknown parameters alone) are supported '
                'as objectives, got the term {}.'.format(term))
        integral = factors[0]
        if integral.function.atoms(sm.Integral):
            raise NotImplementedError('Nested integrals are not supported.')
        if integral.limits != ((time_symbol,),):
            raise NotImplementedError(
                'Only indefinite integrals of time are supported.')
        integrand += coeff * integral.function
    return integrand, rest


def create_objective_function(objective, state_symbols,
                              unknown_input_trajectories, unknown_parameters,
                              num_collocation_nodes, node_time_interval,
                              integration_method='backward euler',
                              time_symbol=None, device=0, tmp_dir=None,
                              cuda_options=None):
    """Returns ``(obj, obj_grad)`` with the call signatures of the
    reference: ``obj(free) -> float``, ``obj_grad(free) -> ndarray`` of shape
    ``((n + q)*N + r,)``.  Arguments as in opty/utils.py:329-364; ``device``,
    ``tmp_dir`` and ``cuda_options`` select the GPU, the compiled-module cache
    and kernel options."""
    from .direct_collocation import (DEFAULT_CUDA_OPTIONS,
                                     attach_extra_modules,
                                     prepare_program_module)
    if time_symbol is None:
        time_symbol = me.dynamicsymbols._t
    if integration_method not in _RULES:
        raise NotImplementedError(
            "Integration method '{}' is not implemented.".format(
                integration_method))
    rule = _RULES[integration_method]
    states = list(state_symbols)
    inputs = list(sort_sympy(unknown_input_trajectories))
    params = list(sort_sympy(unknown_parameters))
    n, q, r = len(states), len(inputs), len(params)
    N = int(num_collocation_nodes)
    h = float(node_time_interval)
    arrays = states + inputs
    na = n + q

    integrand, rest = _split_objective(sm.sympify(objective), time_symbol)
    if rest.atoms(sm.Function) - set():
        bad = [f for f in rest.atoms(sm.Function)
               if f in arrays or getattr(f, 'args', ()) == (time_symbol,)]
        if bad:
            raise NotImplementedError(
                'Terms outside an integral may only depend on the unknown '
                'parameters, found {}.'.format(bad))
    # parameter-only part: r scalars, evaluated on the host
    rest_f = sm.lambdify([params], [rest] + [rest.diff(p) for p in params],
                         modules='numpy')

    handle = None
    if integrand != 0:
        cur = [sm.Symbol('opty_obj_a{}i'.format(k), real=True)
               for k in range(na)]
        nxt = [sm.Symbol('opty_obj_a{}n'.format(k), real=True)
               for k in range(na)]
        at_node = dict(zip(arrays, cur))
        f_node = me.msubs(integrand, at_node)
        grad_x = [me.msubs(integrand.diff(a), at_node) for a in arrays]
        if rule == 1:
            at_mid = {a: (c + m) / 2 for a, c, m in zip(arrays, cur, nxt)}
            f_sum = me.msubs(integrand, at_mid)
            grad_p = [me.msubs(integrand.diff(p), at_mid) for p in params]
            next_args = dict(zip(cur, nxt))
        else:
            f_sum = f_node
            grad_p = [me.msubs(integrand.diff(p), at_node) for p in params]
            next_args = None
        matrix = sm.Matrix([[f_sum] + grad_x + grad_p])
        opts = dict(DEFAULT_CUDA_OPTIONS)
        opts['d2h_skip_constants'] = False
        opts['prefetch_jacobian'] = False
        opts['out_ring'] = 1
        if cuda_options:
            opts.update(cuda_options)
        prog = CollocationProgram.from_matrix(
            cur + params, matrix, const=params,
            use_sympy_cse=opts['use_sympy_cse'], next_args=next_args)
        (_, _, _, meta, cubin, _, _) = prepare_program_module(
            prog, N, 'elementwise', opts, tmp_dir=tmp_dir)
        cfg = runtime.ColloCfg()
        cfg.abi_version = runtime.ABI_VERSION
        cfg.device = int(device)
        cfg.N = N
        cfg.node_lo, cfg.node_hi = 0, N
        cfg.n = na
        cfg.q = cfg.k = cfg.s = cfg.pk = 0
        cfg.r = r
        cfg.M, cfg.P = 1, 1 + na + r
        cfg.method = _ELEMENTWISE
        cfg.out_ring = 1
        cfg.prefetch_jac = 0
        cfg.con_tail = cfg.jac_tail = 0
        cfg.h = 0.0
        handle = runtime.ColloHandle(cfg, cubin)
        attach_extra_modules(handle, meta)
        handle.set_known(None, None)

    num_free = na * N + r

    def _evaluate(free):
        free = np.ascontiguousarray(free, dtype=np.float64)
        if free.shape != (num_free,):
            raise ValueError('free must have shape ({},), got {}.'.format(
                num_free, free.shape))
        tail = np.asarray(rest_f(free[na * N:]), dtype=float)
        if handle is None:
            grad = np.zeros(num_free)
            value = 0.0
        else:
            value, grad = handle.quadrature(free, h, rule)
        grad[na * N:] += tail[1:]
        return value + float(tail[0]), grad

    def obj(free):
        return _evaluate(free)[0]

    def obj_grad(free):
        return _evaluate(free)[1]

    obj.handle = handle
    return obj, obj_grad
