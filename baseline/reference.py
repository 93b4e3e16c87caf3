"""Loader for the UNMODIFIED reference (csu-hmc/opty) installed under
``baseline/_ref`` by

    python -m pip install --no-index --no-build-isolation --no-deps \
        --find-links /opt/wheelhouse --target baseline/_ref <copy of /root/reference>

(``/root/reference`` is read-only, so the wheel is built from a copy under
/tmp; ``--no-deps`` because ``cyipopt`` is not installable offline).  Used by
``bench.py --impl reference`` / ``cpu_baseline`` and by the fixture generators;
nothing under ``opty_b200/`` imports this.

The reference generates C + Cython at run time and compiles it with the
interpreter's ``sysconfig`` compiler; this image's default ``$CC`` wrapper
cannot link ``-fopenmp``, hence ``CC=/usr/bin/gcc``.  Generated
modules are cached in ``baseline/_ref/_codegen`` through the reference's own
``tmp_dir`` mechanism (opty/utils.py:824-864): ``__graft_entry__.build()``
fills it in the build container and it travels to the GPU box with
``baseline/_ref`` (same image, same Python ABI).
"""

import ctypes
import ctypes.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')
CODEGEN_DIR = os.path.join(REF_DIR, '_codegen')


def available():
    return os.path.isdir(os.path.join(REF_DIR, 'opty'))


def host_threads():
    """Hardware threads this process may run on."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def prepare_environment(threads=None):
    """Must run before the first generated module (and with it libgomp) is
    loaded.  ``torchrun`` exports ``OMP_NUM_THREADS=1`` for nproc > 1: the
    CPU arm overrides it explicitly."""
    threads = int(threads or host_threads())
    os.environ['OMP_NUM_THREADS'] = str(threads)
    os.environ.pop('OMP_THREAD_LIMIT', None)
    if os.path.exists('/usr/bin/gcc'):
        os.environ['CC'] = '/usr/bin/gcc'
        os.environ['LDSHARED'] = '/usr/bin/gcc -shared'
    for p in (os.path.join(HERE, 'stubs'), REF_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    return threads


def omp_max_threads():
    """``omp_get_max_threads()`` of the libgomp the generated modules use
    (None if libgomp is not loadable)."""
    for name in ('libgomp.so.1', ctypes.util.find_library('gomp')):
        if not name:
            continue
        try:
            lib = ctypes.CDLL(name)
            lib.omp_get_max_threads.restype = ctypes.c_int
            return int(lib.omp_get_max_threads())
        except OSError:
            continue
    return None


def collocator(workload, parallel, tmp_dir=CODEGEN_DIR):
    """The reference's ``ConstraintCollocator(backend='cython')`` for a
    ``workloads.Workload``."""
    from opty.direct_collocation import ConstraintCollocator
    os.makedirs(tmp_dir, exist_ok=True)
    return ConstraintCollocator(
        *workload.collocator_args(), **workload.collocator_kwargs(),
        backend='cython', parallel=parallel, tmp_dir=tmp_dir)
