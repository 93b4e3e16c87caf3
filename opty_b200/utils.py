"""Host-side helpers of the collocation hot path.

Same names and call signatures as the corresponding functions of
``opty.utils`` so that code written against the reference keeps working.
"""

import numpy as np

__all__ = ['parse_free', 'sort_sympy', 'ufuncify_matrix']


def parse_free(free, n, q, N, variable_duration=False):
    """Splits the free vector into its parts (views, no copies).

    Follows ``opty.utils.parse_free`` (opty/utils.py:277-326): ``free`` is
    ordered ``[x_1(0..N-1), ..., x_n(...), u_1(...), ..., u_q(...), p_1..p_r,
    h]`` (opty/direct_collocation.py:116-125).

    Parameters
    ----------
    free : ndarray, shape(n*N + q*N + r + s)
    n : integer, number of states
    q : integer, number of unknown input trajectories
    N : integer, number of collocation nodes
    variable_duration : boolean, optional
        True if the last entry of ``free`` is the node time interval.

    Returns
    -------
    states : ndarray, shape(n, N)
    specified_values : ndarray, shape(q, N), shape(N,) if q == 1, or None if
        q == 0
    constant_values : ndarray, shape(r,)
    time_interval : float, only if ``variable_duration``

    """
    split_x = n * N
    split_u = split_x + q * N
    states = free[:split_x].reshape((n, N))
    if q == 0:
        specified = None
    elif q == 1:
        specified = free[split_x:split_u]
    else:
        specified = free[split_x:split_u].reshape((q, N))
    if variable_duration:
        return states, specified, free[split_u:-1], free[-1]
    return states, specified, free[split_u:]


def sort_sympy(seq):
    """Returns the symbols sorted by name, or the functions of time sorted by
    function name (opty/utils.py:473-480)."""
    items = list(seq)
    try:
        return sorted(items, key=lambda s: s.name)
    except AttributeError:
        return sorted(items, key=lambda f: f.__class__.__name__)


def _coo_matrix(jac_vals, row_idxs, col_idxs):
    """Dense array from triplets; later duplicates overwrite earlier ones
    (the behaviour the reference's tests rely on, opty/utils.py:38-44)."""
    dense = np.zeros((int(np.max(row_idxs)) + 1, int(np.max(col_idxs)) + 1),
                     dtype=np.asarray(jac_vals).dtype)
    # np's fancy assignment keeps the last value for repeated indices
    dense[np.asarray(row_idxs), np.asarray(col_idxs)] = jac_vals
    return dense


def ufuncify_matrix(args, expr, const=None, tmp_dir=None, parallel=False,
                    show_compile_output=False, **kwargs):
    """CUDA version of ``opty.utils.ufuncify_matrix`` (opty/utils.py:639-670);
    see :mod:`opty_b200.ufuncify`."""
    from .ufuncify import ufuncify_matrix as _impl
    return _impl(args, expr, const=const, tmp_dir=tmp_dir, parallel=parallel,
                 show_compile_output=show_compile_output, **kwargs)


def create_objective_function(objective, state_symbols,
                              unknown_input_trajectories, unknown_parameters,
                              num_collocation_nodes, node_time_interval,
                              integration_method='backward euler',
                              time_symbol=None, **kwargs):
    """CUDA version of ``opty.utils.create_objective_function``
    (opty/utils.py:329-470); see :mod:`opty_b200.objective`."""
    from .objective import create_objective_function as _impl
    return _impl(objective, state_symbols, unknown_input_trajectories,
                 unknown_parameters, num_collocation_nodes,
                 node_time_interval, integration_method=integration_method,
                 time_symbol=time_symbol, **kwargs)
