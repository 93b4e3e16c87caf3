"""TEST INFRASTRUCTURE ONLY: runs an emitted CUDA-C module on the CPU by
compiling it with g++ against ``tests/host_shim/colloc_kernel.cuh``.

This exists so that the ``-m "not gpu"`` suite can check the emitter's
arithmetic (lowering, forward-mode differentiation, group partition and tile
bookkeeping) against the oracle.  Nothing under ``opty_b200/`` imports it.
"""

import ctypes
import hashlib
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(_HERE, 'host_shim')
_BUILD = os.path.join(tempfile.gettempdir(), 'opty_b200_host_harness')


def compile_for_host(source):
    os.makedirs(_BUILD, exist_ok=True)
    key = hashlib.sha256(source.encode()).hexdigest()[:24]
    so = os.path.join(_BUILD, 'mod_{}.so'.format(key))
    if not os.path.exists(so):
        src = os.path.join(_BUILD, 'mod_{}.cpp'.format(key))
        with open(src, 'w') as f:
            f.write(source)
        # -ffp-contract=off: no FMA contraction, like nvcc --fmad=false
        cmd = ['g++', '-O1', '-ffp-contract=off', '-shared', '-fPIC', '-w',
               '-I', SHIM, '-o', so, src]
        subprocess.run(cmd, check=True, capture_output=True, text=True)
    return ctypes.CDLL(so)


def _prepare_without_nvcc(collocator):
    """The product's own lowering / grouping / emission
    (``_PreparedModule``), minus the nvcc step."""
    from opty_b200 import build
    from opty_b200.direct_collocation import _PreparedModule
    real = build.compile_module
    build.compile_module = lambda *a, **k: (b'', '', False)
    try:
        collocator._cuda_options['use_index'] = False   # the source is needed
        if collocator._cuda_options['groups'] == 'auto':
            collocator._cuda_options['groups'] = 3
        pm = _PreparedModule(collocator)
    finally:
        build.compile_module = real
    return pm.program, pm.source, pm.meta


def host_evaluate(collocator, free, known_traj=None):
    """Evaluates constraints and Jacobian of ``collocator`` (an
    ``opty_b200.ConstraintCollocator``) at ``free`` with the emitted code
    compiled for the host.  Returns ``(con, jac)`` in the reference layouts
    (eom-major residuals, node-major partials), EOM part only."""
    prepared = _prepare_without_nvcc(collocator)
    prog, source, meta = prepared
    lib = compile_for_host(source)
    N = collocator.num_collocation_nodes
    n = collocator.num_states
    q = collocator.num_unknown_input_trajectories
    k = collocator.num_known_input_trajectories
    r = collocator.num_unknown_parameters
    free = np.ascontiguousarray(free, dtype=float)
    traj = np.zeros((n + q + k + meta['D'], N))
    traj[:n + q] = free[:(n + q) * N].reshape(n + q, N)
    for i, sym in enumerate(collocator.known_input_trajectories):
        val = collocator.known_trajectory_map[sym]
        traj[n + q + i] = val(free) if callable(val) else val
    uni = [float(collocator.known_parameter_map[p])
           for p in collocator.known_parameters]
    uni += list(free[(n + q) * N:(n + q) * N + r])
    if collocator._variable_duration:
        uni.append(free[-1])
    else:
        uni.append(float(collocator.node_time_interval))
    uni = np.array(uni, dtype=float)
    nn = N - 1
    con = np.empty(prog.M * nn)
    jac = np.empty(nn * prog.K)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.host_eval.argtypes = [dp, dp, ctypes.c_longlong, ctypes.c_int, dp, dp]
    lib.host_eval(uni.ctypes.data_as(dp), traj.ctypes.data_as(dp), N, nn,
                  con.ctypes.data_as(dp), jac.ctypes.data_as(dp))
    return con, jac


class HostQuadratureHandle(object):
    """Stand-in for ``runtime.ColloHandle`` of an objective module: runs the
    emitted integrand code on the CPU (host shim) and applies the quadrature
    weights of ``opty_colloc_quadrature`` (csrc/runtime.cu) in NumPy."""

    source = None     # set by ``host_objective`` before the handle is built

    def __init__(self, cfg, cubin):
        self.cfg = cfg
        self.lib = compile_for_host(HostQuadratureHandle.source)

    def set_known(self, traj, params):
        pass

    def quadrature(self, free, h, rule):
        c = self.cfg
        N, na, r, P = c.N, c.n, c.r, c.P
        ldt = N + 16
        traj = np.zeros((max(na, 1), ldt))
        traj[:na, :N] = free[:na * N].reshape(na, N)
        uni = np.concatenate((free[na * N:], [0.0]))
        con = np.zeros(1)
        jac = np.zeros(N * P)
        dp = ctypes.POINTER(ctypes.c_double)
        self.lib.host_eval.argtypes = [dp, dp, ctypes.c_longlong, ctypes.c_int,
                                       dp, dp]
        self.lib.host_eval(uni.ctypes.data_as(dp), traj.ctypes.data_as(dp),
                           ldt, N, con.ctypes.data_as(dp),
                           jac.ctypes.data_as(dp))
        vals = jac.reshape(N, P)
        i = np.arange(N)
        if rule == 1:
            ws = (i < N - 1).astype(float)
            wt = np.where((i == 0) | (i == N - 1), 0.5, 1.0)
        else:
            ws = (i > 0).astype(float)
            wt = ws
        keep = ws != 0
        value = h * np.sum(vals[keep, 0])
        grad = np.zeros(na * N + r)
        for a in range(na):
            grad[a * N:(a + 1) * N] = h * wt * vals[:, 1 + a]
        for s in range(r):
            grad[na * N + s] = h * np.sum(vals[keep, 1 + na + s])
        return value, grad


def host_objective(*args, **kwargs):
    """``opty_b200.objective.create_objective_function`` with the device
    handle replaced by :class:`HostQuadratureHandle` and nvcc skipped."""
    from opty_b200 import build, objective
    real_compile, real_handle = build.compile_module, \
        objective.runtime.ColloHandle

    def capture(src, flags, **k):
        HostQuadratureHandle.source = src
        return (b'', '', False)
    build.compile_module = capture
    objective.runtime.ColloHandle = HostQuadratureHandle
    try:
        kwargs.setdefault('cuda_options', {})['use_index'] = False
        return objective.create_objective_function(*args, **kwargs)
    finally:
        build.compile_module = real_compile
        objective.runtime.ColloHandle = real_handle
