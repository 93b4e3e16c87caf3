"""ctypes binding of the C-ABI in ``include/opty_b200.h``.

The reference binds its generated C through a generated Cython wrapper
(opty/utils.py:500-529); here one fixed shared library is bound once with
``ctypes`` and the generated code travels as a cubin blob.
"""

import ctypes
import os

import numpy as np

from . import build

OPTY_MAX_GROUPS = 1024
ABI_VERSION = 7

EXPORTS = (
    'opty_b200_abi_version', 'opty_colloc_create', 'opty_colloc_destroy',
    'opty_colloc_add_module', 'opty_colloc_set_known',
    'opty_colloc_upload_free', 'opty_colloc_eval_device',
    'opty_colloc_constraints', 'opty_colloc_jacobian', 'opty_colloc_begin',
    'opty_colloc_finish', 'opty_colloc_set_host_outputs', 'opty_host_alloc',
    'opty_host_free', 'opty_colloc_host_buffers',
    'opty_colloc_device_buffers', 'opty_colloc_set_d2h_columns',
    'opty_colloc_invalidate_host_jacobian', 'opty_colloc_quadrature',
    'opty_colloc_last_kernel_ms', 'opty_colloc_time_device_evals',
    'opty_colloc_launch_count', 'opty_colloc_jacobian_indices',
    'opty_colloc_last_error',
)


class ColloCfg(ctypes.Structure):
    """``opty_colloc_cfg`` of include/opty_b200.h: the problem in the
    reference's notation; no kernel geometry."""
    _fields_ = [
        ('abi_version', ctypes.c_int32),
        ('device', ctypes.c_int32),
        ('N', ctypes.c_int32),
        ('node_lo', ctypes.c_int32),
        ('node_hi', ctypes.c_int32),
        ('n', ctypes.c_int32),
        ('q', ctypes.c_int32),
        ('k', ctypes.c_int32),
        ('r', ctypes.c_int32),
        ('s', ctypes.c_int32),
        ('pk', ctypes.c_int32),
        ('M', ctypes.c_int32),
        ('P', ctypes.c_int32),
        ('method', ctypes.c_int32),
        ('out_ring', ctypes.c_int32),
        ('prefetch_jac', ctypes.c_int32),
        ('con_tail', ctypes.c_int32),
        ('jac_tail', ctypes.c_int32),
        ('h', ctypes.c_double),
    ]


_lib = None


def load_library(path=None):
    """Loads ``libopty_b200.so`` (building it first if needed) and declares
    the prototypes.  Raises ImportError if it cannot be built or loaded: there
    is no CPU fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    lib_path = path or build.build_runtime()
    try:
        lib = ctypes.CDLL(lib_path)
    except OSError as err:
        raise ImportError('Unable to load the opty_b200 runtime library {}: '
                          '{}'.format(lib_path, err)) from err
    c_dp = ctypes.POINTER(ctypes.c_double)
    c_vp = ctypes.c_void_p
    lib.opty_b200_abi_version.restype = ctypes.c_int
    lib.opty_colloc_last_error.restype = ctypes.c_char_p
    lib.opty_colloc_create.argtypes = [ctypes.POINTER(ColloCfg), c_vp,
                                       ctypes.c_size_t, ctypes.POINTER(c_vp)]
    lib.opty_colloc_destroy.argtypes = [c_vp]
    lib.opty_colloc_set_known.argtypes = [c_vp, c_vp, c_vp]
    lib.opty_colloc_upload_free.argtypes = [c_vp, c_vp]
    lib.opty_colloc_eval_device.argtypes = [c_vp, ctypes.c_int]
    lib.opty_colloc_constraints.argtypes = [c_vp, c_vp, c_vp]
    lib.opty_colloc_jacobian.argtypes = [c_vp, c_vp, c_vp]
    lib.opty_colloc_host_buffers.argtypes = [c_vp, ctypes.POINTER(c_dp),
                                             ctypes.POINTER(c_dp),
                                             ctypes.POINTER(c_dp)]
    lib.opty_colloc_device_buffers.argtypes = [
        c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(ctypes.c_int64),
        ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), ctypes.POINTER(c_vp)]
    lib.opty_colloc_set_d2h_columns.argtypes = [c_vp, ctypes.c_int, c_vp,
                                                c_vp]
    lib.opty_colloc_invalidate_host_jacobian.argtypes = [c_vp]
    lib.opty_colloc_add_module.argtypes = [c_vp, c_vp, ctypes.c_size_t]
    lib.opty_colloc_begin.argtypes = [c_vp, c_vp, ctypes.c_int, ctypes.c_int]
    lib.opty_colloc_finish.argtypes = [c_vp]
    lib.opty_colloc_set_host_outputs.argtypes = [c_vp, c_vp, c_vp]
    lib.opty_host_alloc.argtypes = [ctypes.c_size_t, ctypes.POINTER(c_vp)]
    lib.opty_host_free.argtypes = [c_vp]
    lib.opty_colloc_quadrature.argtypes = [
        c_vp, c_vp, ctypes.c_double, ctypes.c_int,
        ctypes.POINTER(ctypes.c_double), c_vp]
    lib.opty_colloc_last_kernel_ms.argtypes = [c_vp,
                                               ctypes.POINTER(ctypes.c_float)]
    lib.opty_colloc_time_device_evals.argtypes = [
        c_vp, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
    lib.opty_colloc_launch_count.argtypes = [c_vp,
                                             ctypes.POINTER(ctypes.c_int64)]
    lib.opty_colloc_jacobian_indices.argtypes = [ctypes.c_int] * 10 + [c_vp,
                                                                       c_vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ('opty_colloc_last_error',):
            fn.restype = ctypes.c_int
    if lib.opty_b200_abi_version() != ABI_VERSION:
        raise ImportError('libopty_b200.so ABI version mismatch; rebuild it.')
    if path is None:
        _lib = lib
    return lib


def _check(lib, rc):
    if rc != 0:
        msg = lib.opty_colloc_last_error().decode(errors='replace')
        if rc == -1:
            raise ValueError('opty_b200: ' + msg)
        raise RuntimeError('opty_b200 (code {}): {}'.format(rc, msg))


def _as_f64(arr, length, name):
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    if arr.ndim != 1 or arr.shape[0] != length:
        raise ValueError('{} must have shape ({},), got {}'.format(
            name, length, arr.shape))
    return arr


def _view(ptr, count):
    if count == 0:
        return np.empty(0)
    buf = (ctypes.c_double * count).from_address(
        ctypes.addressof(ptr.contents))
    return np.frombuffer(buf, dtype=np.float64, count=count)


class ColloHandle(object):
    """Owns one ``opty_colloc_t``: one device, one node range."""

    def __init__(self, cfg, cubin):
        self.lib = load_library()
        self.cfg = cfg
        self._cubin = ctypes.create_string_buffer(cubin, len(cubin))
        handle = ctypes.c_void_p()
        _check(self.lib, self.lib.opty_colloc_create(
            ctypes.byref(cfg), ctypes.cast(self._cubin, ctypes.c_void_p),
            len(cubin), ctypes.byref(handle)))
        self._h = handle
        self.nn = cfg.node_hi - cfg.node_lo
        self.K = cfg.M * cfg.P
        self.free_len = (cfg.n + cfg.q) * cfg.N + cfg.r + cfg.s
        self.con_len = cfg.M * self.nn
        self.jac_len = self.nn * self.K
        pf = ctypes.POINTER(ctypes.c_double)()
        pc = ctypes.POINTER(ctypes.c_double)()
        _check(self.lib, self.lib.opty_colloc_host_buffers(
            self._h, ctypes.byref(pf), ctypes.byref(pc), None))
        self.free_pinned = _view(pf, self.free_len)
        self.con_pinned = _view(pc, self.con_len + cfg.con_tail)
        self._jac_views = {}
        self.jac_pinned = None    # pinned Jacobian buffers are lazy

    def _current_jac_view(self):
        """NumPy view of the pinned buffer that holds the most recently
        fetched Jacobian (with speculative copies there are two buffers that
        take turns)."""
        pj = ctypes.POINTER(ctypes.c_double)()
        _check(self.lib, self.lib.opty_colloc_host_buffers(
            self._h, None, None, ctypes.byref(pj)))
        addr = ctypes.addressof(pj.contents)
        view = self._jac_views.get(addr)
        if view is None:
            view = _view(pj, self.jac_len + self.cfg.jac_tail)
            self._jac_views[addr] = view
        return view

    def close(self):
        if getattr(self, '_h', None) is not None and self._h:
            self.free_pinned = self.con_pinned = self.jac_pinned = None
            self._jac_views = {}
            self.lib.opty_colloc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_known(self, traj, params):
        cfg = self.cfg
        tp = pp = None
        if cfg.k > 0:
            traj = np.ascontiguousarray(traj, dtype=np.float64)
            if traj.shape != (cfg.k, cfg.N):
                raise ValueError('known trajectories must have shape '
                                 '({}, {})'.format(cfg.k, cfg.N))
            tp = traj.ctypes.data
        if cfg.pk > 0:
            params = _as_f64(params, cfg.pk, 'known parameters')
            pp = params.ctypes.data
        _check(self.lib, self.lib.opty_colloc_set_known(self._h, tp, pp))

    def upload_free(self, free):
        free = _as_f64(free, self.free_len, 'free')
        _check(self.lib, self.lib.opty_colloc_upload_free(
            self._h, free.ctypes.data))

    def eval_device(self, sync=True):
        _check(self.lib, self.lib.opty_colloc_eval_device(self._h,
                                                          1 if sync else 0))

    def constraints(self, free):
        """Returns a view of the pinned residual buffer (valid until the next
        call)."""
        free = _as_f64(free, self.free_len, 'free')
        _check(self.lib, self.lib.opty_colloc_constraints(
            self._h, free.ctypes.data, None))
        return self.con_pinned

    def jacobian(self, free):
        """Returns a view of the pinned Jacobian buffer (valid until the next
        call)."""
        free = _as_f64(free, self.free_len, 'free')
        _check(self.lib, self.lib.opty_colloc_jacobian(
            self._h, free.ctypes.data, None))
        self.jac_pinned = self._current_jac_view()
        return self.jac_pinned

    def device_buffers(self):
        traj = ctypes.c_void_p()
        con = ctypes.c_void_p()
        jac = ctypes.c_void_p()
        uni = ctypes.c_void_p()
        ldt = ctypes.c_int64()
        _check(self.lib, self.lib.opty_colloc_device_buffers(
            self._h, ctypes.byref(traj), ctypes.byref(ldt), ctypes.byref(con),
            ctypes.byref(jac), ctypes.byref(uni)))
        return {'traj': traj.value, 'ldt': ldt.value, 'con': con.value,
                'jac': jac.value, 'uni': uni.value}

    def set_d2h_columns(self, ranges):
        n = len(ranges)
        b = (ctypes.c_int32 * max(n, 1))(*[r[0] for r in ranges])
        e = (ctypes.c_int32 * max(n, 1))(*[r[1] for r in ranges])
        _check(self.lib, self.lib.opty_colloc_set_d2h_columns(
            self._h, n, ctypes.cast(b, ctypes.c_void_p),
            ctypes.cast(e, ctypes.c_void_p)))

    def invalidate_host_jacobian(self):
        _check(self.lib, self.lib.opty_colloc_invalidate_host_jacobian(
            self._h))

    def add_module(self, cubin):
        buf = ctypes.create_string_buffer(cubin, len(cubin))
        self._extra_cubins = getattr(self, '_extra_cubins', []) + [buf]
        _check(self.lib, self.lib.opty_colloc_add_module(
            self._h, ctypes.cast(buf, ctypes.c_void_p), len(cubin)))

    def set_host_outputs(self, con_full, jac_full):
        """Redirects this handle's device->host copies into its slices of
        full-problem host vectors (``PinnedArray`` or None, None)."""
        cp = None if con_full is None else con_full.ctypes.data
        jp = None if jac_full is None else jac_full.ctypes.data
        _check(self.lib, self.lib.opty_colloc_set_host_outputs(self._h, cp,
                                                               jp))

    def begin(self, free, want_con, want_jac):
        """``free`` must stay alive until :meth:`finish`."""
        _check(self.lib, self.lib.opty_colloc_begin(
            self._h, free.ctypes.data, int(want_con), int(want_jac)))

    def finish(self):
        _check(self.lib, self.lib.opty_colloc_finish(self._h))

    def quadrature(self, free, scale, rule):
        """``(value, grad)`` of the running-cost integral; ``grad`` has one
        entry per array argument and node followed by one per scalar
        argument."""
        free = _as_f64(free, self.free_len, 'free')
        value = ctypes.c_double()
        grad = np.empty(self.cfg.n * self.cfg.N + self.cfg.r)
        _check(self.lib, self.lib.opty_colloc_quadrature(
            self._h, free.ctypes.data, float(scale), int(rule),
            ctypes.byref(value), grad.ctypes.data))
        return value.value, grad

    def last_kernel_ms(self):
        ms = ctypes.c_float()
        _check(self.lib, self.lib.opty_colloc_last_kernel_ms(
            self._h, ctypes.byref(ms)))
        return ms.value

    def time_device_evals(self, steps):
        """Device time (ms, CUDA events on the launching stream) of ``steps``
        back-to-back device-resident evaluations."""
        ms = ctypes.c_float()
        _check(self.lib, self.lib.opty_colloc_time_device_evals(
            self._h, int(steps), ctypes.byref(ms)))
        return ms.value

    def launch_count(self):
        cnt = ctypes.c_int64()
        _check(self.lib, self.lib.opty_colloc_launch_count(
            self._h, ctypes.byref(cnt)))
        return cnt.value


class PinnedArray(object):
    """Page-locked, device-portable float64 host vector (``opty_host_alloc``)
    with a NumPy view ``.array``."""

    def __init__(self, count):
        self.lib = load_library()
        ptr = ctypes.c_void_p()
        _check(self.lib, self.lib.opty_host_alloc(max(int(count), 1) * 8,
                                                  ctypes.byref(ptr)))
        self._ptr = ptr
        buf = (ctypes.c_double * int(count)).from_address(ptr.value)
        self.array = np.frombuffer(buf, dtype=np.float64, count=int(count))
        self.ctypes = self.array.ctypes

    def close(self):
        if getattr(self, '_ptr', None):
            self.array = None
            self.ctypes = None
            self.lib.opty_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def jacobian_indices(device, N, node_lo, node_hi, n, q, r, s, M, method):
    """COO rows/cols (int64) of the EOM part of the constraint Jacobian for
    constraint nodes ``[node_lo, node_hi)``, generated on the device."""
    lib = load_library()
    P = (2 * n + 2 * q if method == 1 else 2 * n + q) + r + s
    count = (node_hi - node_lo) * M * P
    rows = np.empty(count, dtype=np.int64)
    cols = np.empty(count, dtype=np.int64)
    _check(lib, lib.opty_colloc_jacobian_indices(
        device, N, node_lo, node_hi, n, q, r, s, M, method,
        rows.ctypes.data, cols.ctypes.data))
    return rows, cols
