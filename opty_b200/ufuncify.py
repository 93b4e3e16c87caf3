"""CUDA version of the reference's generic operator ``ufuncify_matrix``
(opty/utils.py:639-928): "a function that evaluates a matrix of expressions in
a tight loop".

Same call signature and the same returned-function convention::

    f = ufuncify_matrix(args, expr, const=None, tmp_dir=None, parallel=False,
                        show_compile_output=False)
    result = f(matrix, *values)      # matrix: C-contiguous float64 (n, rows*cols)
    # result is ``matrix`` (filled) reshaped to (n, rows, cols)

Non-``const`` values are C-contiguous float64 arrays of length ``n``,
``const`` values floats (opty/utils.py:781-793).  ``parallel`` is accepted and
ignored: the CUDA grid is the loop.  Symbol names never reach a compiler here,
so names that are C keywords or invalid C identifiers (which make the
reference raise ImportError, opty/tests/test_utils.py:301-336) are fine.
"""

import numpy as np

from . import runtime
from .direct_collocation import (DEFAULT_CUDA_OPTIONS,
                                 prepare_program_module,
                                 attach_extra_modules)
from .program import CollocationProgram

ELEMENTWISE = 2


def ufuncify_matrix(args, expr, const=None, tmp_dir=None, parallel=False,
                    show_compile_output=False, device=0, cuda_options=None):
    args = list(args)
    const = tuple(const) if const is not None else ()
    opts = dict(DEFAULT_CUDA_OPTIONS)
    opts['d2h_skip_constants'] = False
    opts['prefetch_jacobian'] = False
    if cuda_options:
        opts.update(cuda_options)
    prog = CollocationProgram.from_matrix(args, expr, const=const,
                                          use_sympy_cse=opts['use_sympy_cse'])
    rows, cols = prog.M, prog.P
    is_const = [a in const for a in args]
    num_arrays = sum(1 for c in is_const if not c)
    handles = {}

    def handle_for(n):
        h = handles.get(n)
        if h is None:
            (parts, derived, source, meta, cubin, path, hit) = \
                prepare_program_module(
                    prog, n, 'elementwise', opts, tmp_dir=tmp_dir,
                    show_compile_output=show_compile_output)
            cfg = runtime.ColloCfg()
            cfg.abi_version = runtime.ABI_VERSION
            cfg.out_ring = int(opts['out_ring'])
            cfg.prefetch_jac = 0
            cfg.con_tail = cfg.jac_tail = 0
            cfg.device = int(device)
            cfg.N = n
            cfg.node_lo, cfg.node_hi = 0, n
            cfg.n = num_arrays
            cfg.q = cfg.k = cfg.s = cfg.pk = 0
            cfg.r = len(const)
            cfg.M, cfg.P = rows, cols
            cfg.method = ELEMENTWISE
            cfg.h = 0.0
            h = runtime.ColloHandle(cfg, cubin)
            attach_extra_modules(h, meta)
            h.set_known(None, None)
            handles[n] = h
        return h

    def eval_matrix_loop(matrix, *values):
        if len(values) != len(args):
            raise TypeError('expected {} arguments after matrix, got {}'
                            .format(len(args), len(values)))
        if (not isinstance(matrix, np.ndarray) or matrix.ndim != 2 or
                matrix.dtype != np.float64 or
                not matrix.flags['C_CONTIGUOUS'] or
                matrix.shape[1] != rows * cols):
            raise ValueError('matrix must be a C-contiguous float64 array of '
                             'shape (n, {})'.format(rows * cols))
        n = matrix.shape[0]
        h = handle_for(n)
        free = h.free_pinned
        pos = 0
        tail = num_arrays * n
        for flag, v in zip(is_const, values):
            if flag:
                free[tail] = float(v)
                tail += 1
            else:
                v = np.asarray(v)
                if (v.dtype != np.float64 or v.ndim != 1 or
                        not v.flags['C_CONTIGUOUS']):
                    raise ValueError('Buffer dtype mismatch or not '
                                     'C-contiguous: array arguments must be '
                                     '1-D C-contiguous float64')
                if v.shape[0] != n:
                    raise ValueError('array arguments must have length '
                                     '{}'.format(n))
                free[pos:pos + n] = v
                pos += n
        out = h.jacobian(free)
        matrix[...] = out[:n * rows * cols].reshape(n, rows * cols)
        return matrix.reshape(n, rows, cols)

    eval_matrix_loop.program = prog
    return eval_matrix_loop
